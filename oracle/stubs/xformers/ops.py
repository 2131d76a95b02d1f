import torch.nn.functional as F


def memory_efficient_attention(query, key, value, attn_bias=None, op=None, scale=None):
    assert attn_bias is None and op is None
    return F.scaled_dot_product_attention(query, key, value, scale=scale)
