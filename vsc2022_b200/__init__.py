"""vsc2022_b200 -- Blackwell-native engine for the vsc2022 video-copy-detection hot path.

Host side mirrors the reference's operator interface (vsc.index, vsc.candidates,
vsc.baseline.score_normalization, vsc.baseline.localization, vcsl.vta); the arithmetic
runs in hand-written sm_100a CUDA behind the C ABI in include/vsc_b200.h.
"""
__version__ = "0.1.0"
