"""Building blocks with the reference's names (basics, unet, backbone) running on the pmoe_b200 tape / eval executors."""
