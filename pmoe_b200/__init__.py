"""pmoe_b200 — B200-native (sm_100a) hot path of mhnazeri/PMoE behind the reference's nn.Module API."""
__all__ = ["ops"]
