"""suo_slam_b200 — B200-native per-frame hot path of SUO-SLAM (rpng/suo_slam).

Host side is Python/PyTorch plumbing only; the hot path is libsuo_b200.so
(hand-written CUDA for sm_100a behind the C ABI in include/suo_b200.h).
There is no CPU fallback: every op raises if the library or a B200 is missing.
"""
__all__ = ["arch", "synth"]
