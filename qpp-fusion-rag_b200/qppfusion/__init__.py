"""
qppfusion -- B200-native (sm_100a) retrieval-scoring -> QPP -> rank-fusion hot path of QPP-Fusion-RAG,
behind the reference's own Python API (src/fusion.py, src/qpp.py, src/retrievers/bge.py).

Importing the package is cheap and GPU-free; the CUDA library (lib/libqppfusion.so) is loaded on first
use and every compute call fails loudly without it or without an sm_100 device (no CPU fallback).
"""
__version__ = "0.1.0"
