"""neuma_b200 -- B200-native (sm_100a) drop-in for NeuMA's MPM + Particle-GS hot path."""
from ._lib import Context, NeumaError, load  # noqa: F401

__all__ = ["Context", "NeumaError", "load"]
