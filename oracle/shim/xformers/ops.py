import torch


def memory_efficient_attention(query, key, value, attn_bias=None, op=None, scale=None, p=0.0):
    """softmax(Q K^T * scale + bias) V on [B, M, H, K] (or [B, M, K]) tensors — the published contract of
    xformers.ops.memory_efficient_attention, evaluated densely in the input dtype."""
    three_d = query.dim() == 3
    if three_d:
        query, key, value = query[:, :, None], key[:, :, None], value[:, :, None]
    q = query.transpose(1, 2)
    k = key.transpose(1, 2)
    v = value.transpose(1, 2)
    if scale is None:
        scale = q.shape[-1] ** -0.5
    s = torch.matmul(q, k.transpose(-1, -2)) * scale
    if attn_bias is not None:
        b = attn_bias
        if b.dim() == 3:  # [B*H or B, Lq, Lk]
            b = b.reshape(q.shape[0], -1, b.shape[-2], b.shape[-1])
        s = s + b
    o = torch.matmul(s.softmax(dim=-1), v).transpose(1, 2)
    return o[:, :, 0] if three_d else o
