"""B200-native MC-CNN stereo-matching hot path (drop-in for WHDY/SceneDepthEstimation's
match.py / match_single.py / process_functional.py / mc_cnn_brunch.py / error_calculate.py).

The CUDA kernels live in csrc/ behind the C ABI of include/mccnn_b200.h; Python is host glue.
"""
__all__ = ["process_functional", "mc_cnn_brunch", "match", "match_single", "error_calculate", "engine", "synthetic"]
