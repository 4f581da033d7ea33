"""mlagg-unet_b200: B200-native (sm_100a) MLAgg + MSMM hot path behind the reference's module API.

Import as `mlagg_unet_b200` (alias package at the repo root)."""
__version__ = "0.1.0"
