"""Stand-in for xformers: memory_efficient_attention == softmax(q k^T * scale) v on [B*H, S, d]."""
from . import ops  # noqa: F401
