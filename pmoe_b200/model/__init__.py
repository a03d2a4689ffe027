"""Reference-named model modules of pmoe_b200 (get_model, PredictiveUnet, UNet ...): the nn.Module API over the sm_100a kernels."""
