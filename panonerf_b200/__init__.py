"""panonerf_b200: B200-native (sm_100a) implementation of Pano-NeRF's mip-NeRF volumetric-rendering hot path."""
from .datasets.base_datasets import Rays  # noqa: F401

__all__ = ["Rays"]
