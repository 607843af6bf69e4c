"""xotorch_b200 — B200-native (sm_100a) shard-decode engine behind xotorch's Shard / InferenceEngine / PeerHandle API.

Only the hot path is here: csrc/ (CUDA kernels + the C ABI of include/xotb200.h), the ctypes
binding (_lib.py) and the host-side mirror of the reference interfaces (inference/, networking/, topology/).
"""
__version__ = "0.1.0"
