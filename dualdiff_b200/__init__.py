"""dualdiff_b200 — B200-native (sm_100a) implementation of DualDiff's denoising-step hot path."""
__version__ = "0.1.0"
