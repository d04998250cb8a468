"""naf_b200 -- B200-native cross-scale neighbourhood-attention forward of valeoai/NAF.

Public surface (mirrors the reference's): `NAF`, `CrossAttention`, `RoPE`, `encoder`,
`naf(pretrained, device)`, `ModelWrapper`; plus `HostPipeline` (stream-overlapped host-to-host serving loop).  All heavy lifting is in libnaf_b200.so (hand-written
sm_100a CUDA behind the C ABI of include/naf_b200.h); there is no CPU or PyTorch fallback.
"""
from . import _lib, ops, taps
from .hub import ModelWrapper, naf
from .layers import CrossAttention, RoPE, encoder
from .model import NAF
from .pipeline import GraphedNAF, HostPipeline

__all__ = ["NAF", "CrossAttention", "RoPE", "encoder", "naf", "ModelWrapper", "HostPipeline", "GraphedNAF", "ops", "taps", "_lib"]
__version__ = "0.1.0"
