from .attentions import CrossAttention
from .convolutions import encoder
from .rope import RoPE

__all__ = ["CrossAttention", "encoder", "RoPE"]
