"""equihgnn_b200 -- B200-native (sm_100a) implementation of EquiHGNN's message-passing hot path.

Importing the package does not touch the GPU; the CUDA library (`csrc/libeqh_b200.so`) is
loaded on first use and every operator fails loudly if it is missing (no CPU fallback).
"""
__version__ = "0.1.0"
