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
    b = attn_bias
    if b is not None and b.dim() == 3:  # [B*H or B, Lq, Lk]
        b = b.reshape(q.shape[0], -1, b.shape[-2], b.shape[-1])
    # batch-chunked (dense scores below ~1 GiB at HD latents); identical to the one-shot evaluation per batch item
    step = max(1, (1 << 28) // max(1, q.shape[1] * q.shape[2] * k.shape[2]))
    outs = []
    for i in range(0, q.shape[0], step):
        s = torch.matmul(q[i:i + step], k[i:i + step].transpose(-1, -2)) * scale
        if b is not None:
            s = s + b[i:i + step]
        outs.append(torch.matmul(s.softmax(dim=-1), v[i:i + step]))
    o = (outs[0] if len(outs) == 1 else torch.cat(outs)).transpose(1, 2)
    return o[:, :, 0] if three_d else o
