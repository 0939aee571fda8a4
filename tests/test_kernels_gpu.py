"""Parity of the non-GEMM CUDA kernels against torch fp32 references on the same bf16-rounded inputs (GPU).
Tolerances (bf16 outputs): max-abs <= 2e-2 * max|ref| unless stated."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(out, ref):
    return ((out.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-6)).item()


def _mk(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).cuda()


def _heads_pad(x, heads, d, dp):
    """[rows, heads*d] -> [rows, heads*dp] zero padded per head"""
    r = x.shape[0]
    out = x.new_zeros((r, heads, dp))
    out[:, :, :d] = x.reshape(r, heads, d)
    return out.reshape(r, heads * dp)


def _ref_attn(q, k, v, n, lq, lk, heads, d):
    qh = q.float().reshape(n, lq, heads, d).transpose(1, 2)
    kh = k.float().reshape(-1, lk, heads, d).transpose(1, 2)
    vh = v.float().reshape(-1, lk, heads, d).transpose(1, 2)
    return qh, kh, vh


@pytest.mark.parametrize("d,L,n", [(40, 1400, 2), (40, 200, 3), (80, 350, 2), (160, 91, 3), (160, 28, 2), (80, 1400, 1)])
def test_self_attention_fused_qkv(d, L, n):
    from dualdiff_b200 import ops
    heads, C = 8, 8 * d
    q = _mk((n * L, C), 1); k = _mk((n * L, C), 2); v = _mk((n * L, C), 3)
    dp = 48 if d == 40 else d
    fused = torch.cat([_heads_pad(q, heads, d, dp), _heads_pad(k, heads, d, dp), v], dim=1).contiguous()
    out = ops.attention(fused, fused, fused, n_img=n, lq=L, lk=L, heads=heads, head_dim=d,
                        q_col0=0, k_col0=heads * dp, v_col0=2 * heads * dp)
    qh, kh, vh = _ref_attn(q, k, v, n, L, L, heads, d)
    ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(n * L, C)
    assert _rel(out, ref) < 2e-2, _rel(out, ref)


@pytest.mark.parametrize("d,L,peak", [(40, 1400, 40.0), (40, 1400, 400.0), (40, 1400, 4000.0), (80, 350, 900.0), (160, 300, 2500.0)])
def test_self_attention_rising_logits(d, L, peak):
    """keys ordered so that the logits of every query rise steadily along the sequence: every KV tile raises the running
    maximum (lazy reference update + deferred O rescale of the pipelined softmax); at the larger peaks a single 32-key
    chunk jumps by more than 2^64, which exercises the in-place range-guard rescale of l, P and O."""
    from dualdiff_b200 import ops
    n, heads, C = 2, 8, 8 * d
    g = torch.Generator().manual_seed(7)
    q = torch.randn(n * L, C, generator=g) * 0.1 + 1.0
    ramp = (torch.arange(L, dtype=torch.float32) / L).repeat(n)[:, None]
    k = (torch.randn(n * L, C, generator=g) * 0.05 + ramp) * (peak / (d ** 0.5))
    v = torch.randn(n * L, C, generator=g)
    q, k, v = (t.to(torch.bfloat16).cuda() for t in (q, k, v))
    dp = 48 if d == 40 else d
    fused = torch.cat([_heads_pad(q, heads, d, dp), _heads_pad(k, heads, d, dp), v], dim=1).contiguous()
    out = ops.attention(fused, fused, fused, n_img=n, lq=L, lk=L, heads=heads, head_dim=d,
                        q_col0=0, k_col0=heads * dp, v_col0=2 * heads * dp)
    qh, kh, vh = _ref_attn(q, k, v, n, L, L, heads, d)
    ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(n * L, C)
    assert torch.isfinite(out.float()).all()
    assert _rel(out, ref) < 2e-2, _rel(out, ref)


@pytest.mark.parametrize("d,L,lk", [(40, 1400, 106), (80, 350, 110), (160, 91, 78), (40, 1400, 77)])
def test_cross_attention_text(d, L, lk):
    from dualdiff_b200 import ops
    n, heads, C = 2, 8, 8 * d
    q = _mk((n * L, C), 1); k = _mk((n * lk, C), 2); v = _mk((n * lk, C), 3)
    dp = 48 if d == 40 else d
    kv = torch.cat([_heads_pad(k, heads, d, dp), v], dim=1).contiguous()
    out = ops.attention(_heads_pad(q, heads, d, dp), kv, kv, n_img=n, lq=L, lk=lk, heads=heads, head_dim=d,
                        k_col0=0, v_col0=heads * dp)
    qh, kh, vh = _ref_attn(q, k, v, n, L, lk, heads, d)
    ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(n * L, C)
    assert _rel(out, ref) < 2e-2, _rel(out, ref)


@pytest.mark.parametrize("d,L", [(40, 1400), (80, 350), (160, 91)])
def test_cross_view_attention_two_neighbours(d, L):
    """networks/blocks.py:190-217 with neighboring_attn_type='add': out_v = A(q_v, kv_left) + A(q_v, kv_right)"""
    from dualdiff_b200 import ops
    B, n_cam, heads, C = 2, 6, 8, 8 * d
    n = B * n_cam
    nb = {0: [5, 1], 1: [0, 2], 2: [1, 3], 3: [2, 4], 4: [3, 5], 5: [4, 0]}
    q = _mk((n * L, C), 1); k = _mk((n * L, C), 2); v = _mk((n * L, C), 3)
    dp = 48 if d == 40 else d
    fused = torch.cat([_heads_pad(q, heads, d, dp), _heads_pad(k, heads, d, dp), v], dim=1).contiguous()
    kv_map = torch.tensor([[b * n_cam + j for j in nb[c]] for b in range(B) for c in range(n_cam)],
                          dtype=torch.int32).cuda()
    out = ops.attention(fused, fused, fused, n_img=n, lq=L, lk=L, heads=heads, head_dim=d, q_col0=0,
                        k_col0=heads * dp, v_col0=2 * heads * dp, kv_map=kv_map, n_src=2)
    qh, kh, vh = _ref_attn(q, k, v, n, L, L, heads, d)
    ref = torch.zeros_like(qh)
    for s in range(2):
        idx = kv_map[:, s].long()
        ref += F.scaled_dot_product_attention(qh, kh[idx], vh[idx])
    ref = ref.transpose(1, 2).reshape(n * L, C)
    assert _rel(out, ref) < 2e-2, _rel(out, ref)


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("lq,lk,n_src", [(128, 48, 1), (129, 49, 1), (256, 47, 2), (257, 96, 1), (1400, 1400, 2), (640, 1, 1),
                                         (5600, 200, 1), (100, 145, 2)])
def test_head_dim_40_kernel_variants(lq, lk, n_src, variant):
    """the level-0 kernel (two query tiles per CTA, 48-key tiles; variant 0 = with 25 % of the exponentials on the FMA pipe,
    2 = all on MUFU) and the one-tile kernel (1) on ragged shapes: odd / even query-tile counts, key counts around the
    48-key tile, a single key, one and two K/V sources"""
    from dualdiff_b200 import ops
    d, heads, n = 40, 8, 3
    C = heads * d
    q = _mk((n * lq, C), 11); k = _mk((n * lk, C), 12, 2.0); v = _mk((n * lk, C), 13)
    kv = torch.cat([_heads_pad(k, heads, d, 48), v], dim=1).contiguous()
    kv_map = torch.tensor([[(i + 1) % n, (i + 2) % n] for i in range(n)], dtype=torch.int32).cuda() if n_src == 2 else None
    out = ops.attention(_heads_pad(q, heads, d, 48), kv, kv, n_img=n, lq=lq, lk=lk, heads=heads, head_dim=d,
                        k_col0=0, v_col0=heads * 48, kv_map=kv_map, n_src=n_src, variant=variant)
    qh, kh, vh = _ref_attn(q, k, v, n, lq, lk, heads, d)
    if n_src == 1:
        ref = F.scaled_dot_product_attention(qh, kh, vh)
    else:
        ref = sum(F.scaled_dot_product_attention(qh, kh[kv_map[:, s].long()], vh[kv_map[:, s].long()]) for s in range(2))
    ref = ref.transpose(1, 2).reshape(n * lq, C)
    assert torch.isfinite(out.float()).all()
    assert _rel(out, ref) < 2e-2, _rel(out, ref)


@pytest.mark.parametrize("lq,lk,n_src", [(128, 48, 1), (257, 96, 1), (1400, 1400, 2), (640, 1, 1), (1400, 106, 1), (100, 145, 2)])
def test_head_dim_40_denominator_from_ones_column(lq, lk, n_src):
    """`v_ones` layout: V heads on a 48-column stride with 1.0 in column 40, so the softmax denominator is column 40 of the
    P V accumulator (the kernel keeps no row sum).  The rising logits force O (and with it the denominator) to be rescaled."""
    from dualdiff_b200 import ops
    d, heads, n = 40, 8, 3
    C = heads * d
    g = torch.Generator().manual_seed(21)
    ramp = torch.linspace(0, 1, lk).repeat(n)[:, None]
    q = _mk((n * lq, C), 11)
    k = (_mk((n * lk, C), 12, 2.0).cpu().float() * (0.3 + 3.0 * ramp)).to(torch.bfloat16).cuda()
    v = _mk((n * lk, C), 13)
    vp = torch.zeros(n * lk, heads, 48, dtype=torch.bfloat16, device="cuda")
    vp[:, :, :d] = v.reshape(n * lk, heads, d)
    vp[:, :, d] = 1.0
    kv = torch.cat([_heads_pad(k, heads, d, 48), vp.reshape(n * lk, heads * 48)], dim=1).contiguous()
    kv_map = torch.tensor([[(i + 1) % n, (i + 2) % n] for i in range(n)], dtype=torch.int32).cuda() if n_src == 2 else None
    out = ops.attention(_heads_pad(q, heads, d, 48), kv, kv, n_img=n, lq=lq, lk=lk, heads=heads, head_dim=d,
                        k_col0=0, v_col0=heads * 48, kv_map=kv_map, n_src=n_src, v_ones=True)
    qh, kh, vh = _ref_attn(q, k, v, n, lq, lk, heads, d)
    if n_src == 1:
        ref = F.scaled_dot_product_attention(qh, kh, vh)
    else:
        ref = sum(F.scaled_dot_product_attention(qh, kh[kv_map[:, s].long()], vh[kv_map[:, s].long()]) for s in range(2))
    ref = ref.transpose(1, 2).reshape(n * lq, C)
    assert torch.isfinite(out.float()).all()
    assert _rel(out, ref) < 2e-2, _rel(out, ref)
    # the same inputs through the kernel's own row sum agree to bf16 rounding
    base = ops.attention(_heads_pad(q, heads, d, 48), kv, kv, n_img=n, lq=lq, lk=lk, heads=heads, head_dim=d,
                         k_col0=0, v_col0=heads * 48, v_hs=48, kv_map=kv_map, n_src=n_src)
    assert _rel(out, base) < 1e-2


@pytest.mark.parametrize("two_pass", [False, True])
@pytest.mark.parametrize("n,H,W,c1,c2,padded,silu,eps", [
    (3, 28, 50, 320, 0, True, True, 1e-5), (2, 14, 25, 640, 0, False, False, 1e-6), (2, 7, 13, 1280, 1280, True, True, 1e-5),
    (2, 14, 25, 1280, 640, True, True, 1e-5), (2, 28, 50, 640, 320, True, True, 1e-5), (2, 4, 7, 1280, 0, True, True, 1e-5),
    (5, 28, 50, 320, 0, False, False, 1e-6), (1, 56, 100, 320, 0, True, True, 1e-5), (3, 1, 1, 1280, 1280, True, True, 1e-5),
    (2, 5, 7, 256, 0, True, True, 1e-6), (2, 9, 11, 128, 0, True, True, 1e-6),
    (1, 27, 50, 320, 0, True, True, 1e-5), (2, 27, 50, 320, 0, False, True, 1e-5), (1, 15, 25, 1280, 640, True, False, 1e-6)])
def test_groupnorm_silu(n, H, W, c1, c2, padded, silu, eps, two_pass):
    """both forms of the operator: the single-pass kernel (a [rows x few groups] slab in shared memory; 28x50 / 27x50 / 56x100
    maps split their rows over a cluster of 2 / 2 / 8 CTAs, odd row counts unevenly) and the two-kernel fallback (forced here; taken on its own for images too large for a cluster, e.g. 640 channels at 28x50, or
    fewer than 8 channels per group, e.g. the VAE's 128-channel layers)"""
    from dualdiff_b200 import ops, packing
    x1 = _mk((n * H * W, c1), 1) * 2 + 0.5
    x2 = _mk((n * H * W, c2), 2) if c2 else None
    C = c1 + c2
    gamma = (1 + 0.1 * torch.randn(C, generator=torch.Generator().manual_seed(3))).cuda()
    beta = (0.1 * torch.randn(C, generator=torch.Generator().manual_seed(4))).cuda()
    out = ops.groupnorm(x1, gamma, beta, n_img=n, hw=(H, W), x2=x2, eps=eps, silu=silu, padded_out=padded, two_pass=two_pass)
    x = torch.cat([x1, x2], 1) if c2 else x1
    ref = F.group_norm(x.float().reshape(n, H, W, C).permute(0, 3, 1, 2), 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 3, 1)
    ref = packing.to_padded(ref) if padded else ref.reshape(n * H * W, C)
    assert _rel(out, ref) < 1e-2, _rel(out, ref)
    if padded:  # halo must be exactly zero
        o = out.reshape(n, H + 1, W + 1, C)
        assert o[:, H].abs().max() == 0 and o[:, :, W].abs().max() == 0
    again = ops.groupnorm(x1, gamma, beta, n_img=n, hw=(H, W), x2=x2, eps=eps, silu=silu, padded_out=padded, two_pass=two_pass)
    assert torch.equal(out, again)        # fixed reduction orders: bit-reproducible


def test_groupnorm_large_offset_activations():
    """activations with a mean far above their spread (mean 200, std 0.5: E[x^2] - mean^2 would lose every significant bit
    in fp32): the single-pass kernel forms the variance from squared deviations around the group mean"""
    from dualdiff_b200 import ops
    n, H, W, C = 2, 14, 25, 640
    g = torch.Generator().manual_seed(11)
    x = (200.0 + 0.5 * torch.randn(n * H * W, C, generator=g)).to(torch.bfloat16).cuda()
    gamma = torch.ones(C).cuda(); beta = torch.zeros(C).cuda()
    out = ops.groupnorm(x, gamma, beta, n_img=n, hw=(H, W), eps=1e-5, silu=False)
    ref = F.group_norm(x.double().reshape(n, H, W, C).permute(0, 3, 1, 2), 32, None, None, 1e-5).permute(0, 2, 3, 1).reshape(n * H * W, C)
    assert _rel(out, ref.float()) < 2e-2, _rel(out, ref.float())


@pytest.mark.parametrize("rows,C", [(1400, 320), (701, 640), (91, 1280), (5, 1280), (1403, 320), (3, 320), (77, 768), (130, 512),
                                    (9, 2048)])
def test_layernorm(rows, C):
    """the packed kernel (320 / 640 / 1280 / 768 channels: 4 / 2 / 1 / 1 rows per warp, row counts that leave the last warp
    partly empty) and the generic one (512, 2048)"""
    from dualdiff_b200 import ops
    x = _mk((rows, C), 1) * 3 + 1
    gamma = (1 + 0.1 * torch.randn(C, generator=torch.Generator().manual_seed(3))).cuda()
    beta = (0.1 * torch.randn(C, generator=torch.Generator().manual_seed(4))).cuda()
    out = ops.layernorm(x, gamma, beta)
    ref = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    assert _rel(out, ref) < 1e-2


def test_downsample_conv_via_im2col():
    from dualdiff_b200 import ops, packing
    n, H, W, ci, co = 2, 28, 50, 320, 320
    x = _mk((n, H, W, ci), 1); w = _mk((co, ci, 3, 3), 2, (9 * ci) ** -0.5)
    cols, (ho, wo) = ops.im2col_s2(x.reshape(n * H * W, ci), n_img=n, hw=(H, W))
    out = ops.gemm(cols, packing.pack_conv3x3(w))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), None, stride=2, padding=1)
    assert ref.shape[-2:] == (ho, wo)
    assert _rel(out, ref.permute(0, 2, 3, 1).reshape(n * ho * wo, co)) < 1e-2
    # odd sizes: 7x13 -> 4x7
    x = _mk((n, 7, 13, 64), 3); w = _mk((64, 64, 3, 3), 4, 0.05)
    cols, (ho, wo) = ops.im2col_s2(x.reshape(n * 91, 64), n_img=n, hw=(7, 13))
    out = ops.gemm(cols, packing.pack_conv3x3(w))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), None, stride=2, padding=1)
    assert (ho, wo) == (4, 7) and _rel(out, ref.permute(0, 2, 3, 1).reshape(n * 28, 64)) < 1e-2


@pytest.mark.parametrize("hw,hw2", [((4, 7), (7, 13)), ((7, 13), (14, 25)), ((14, 25), (28, 50)), ((8, 8), (16, 16))])
def test_upsample_nearest_to_size(hw, hw2):
    from dualdiff_b200 import ops, packing
    n, C = 2, 64
    x = _mk((n, hw[0], hw[1], C), 1)
    out = ops.upsample_pad(x.reshape(-1, C), n_img=n, hw=hw, hw2=hw2)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), size=hw2, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(out.float(), packing.to_padded(ref))


def _patch_rows(x, cp):
    """torch restatement of dd_nchw_patches: [n, C, H, W] -> [n*H*W, cp], columns tap-major / channel-minor, zero tail"""
    n, c, h, w = x.shape
    cols = F.unfold(x, 3, padding=1)                                         # [n, C*9, H*W], channel-major / tap-minor
    cols = cols.reshape(n, c, 9, h * w).permute(0, 3, 2, 1).reshape(n * h * w, 9 * c)
    out = x.new_zeros((n * h * w, cp))
    out[:, :9 * c] = cols
    return out


@pytest.mark.parametrize("n_outer,n_view,h,w,shared", [(2, 6, 28, 50, True), (1, 3, 5, 7, False), (2, 2, 8, 12, False)])
def test_latent_conv_in(n_outer, n_view, h, w, shared):
    """conv_in 4->320 on fp32 NCHW latents (CFG duplication = stride_outer 0) as ONE K = 40 GEMM over the 3x3 patch matrix:
    the patch kernel is bit-equal to the torch restatement, the GEMM equals F.conv2d (+ the ControlNet's fused condition
    residual, unet_addon_rawbox.py:965,990) and the nine-tap implicit-GEMM convolution on the same weights"""
    from dualdiff_b200 import ops
    from dualdiff_b200.packing import pack_conv3x3, pack_conv3x3_patch
    g = torch.Generator().manual_seed(1)
    n_src = n_view if shared else n_outer * n_view
    lat = torch.randn(n_src, 4, h, w, generator=g)
    wt, b = torch.randn(320, 4, 3, 3, generator=g) * 0.2, torch.randn(320, generator=g)
    res = torch.randn(n_outer * n_view * h * w, 320, generator=g).to(torch.bfloat16)
    kw = dict(n_outer=n_outer, n_view=n_view, c=4, h=h, w=w, stride_outer=0 if shared else n_view * 4 * h * w,
              stride_view=4 * h * w, stride_c=h * w, stride_h=w)
    cols = ops.nchw_patches(lat.cuda(), cp=40, **kw)
    full = torch.cat([lat] * n_outer) if shared else lat
    assert torch.equal(cols.float().cpu(), _patch_rows(full, 40).to(torch.bfloat16).float())
    out = ops.gemm(cols, pack_conv3x3_patch(wt).cuda(), bias=b.cuda(), res1=res.cuda()).float().cpu()
    w8 = torch.zeros(320, 8, 3, 3); w8[:, :4] = wt
    pad = ops.nchw_to_padded(lat.cuda(), cp=8, **kw)
    taps = ops.gemm(pad, pack_conv3x3(w8).cuda(), bias=b.cuda(), taps=9, conv_hw=(h, w), n_img=n_outer * n_view,
                    res1=res.cuda()).float().cpu()
    ref = F.conv2d(full.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float(), b, padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, 320) + res.float()
    assert (out - ref).abs().max() <= 2e-2 * ref.abs().max()
    assert (out - taps).abs().max() <= 2e-2 * ref.abs().max()


def test_small_fp32_pieces():
    from dualdiff_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(12, 189, generator=g).cuda(); w = torch.randn(768, 189, generator=g).cuda() * 0.05
    b = torch.randn(768, generator=g).cuda()
    assert _rel(ops.linear_f32(x, w, b, act=1), F.silu(x @ w.t() + b)) < 1e-5
    t = torch.tensor([801.0, 0.0, 999.0, 40.0]).cuda()
    half = 160
    f = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half).cuda()
    ref = torch.cat([torch.cos(t[:, None] * f), torch.sin(t[:, None] * f)], -1)
    assert (ops.timestep_embedding(t, 320) - ref).abs().max() < 2e-4
    cam = torch.tensor([[1266.4, 0.0, 816.3], [0.3, -2.0, 491.0], [1e-3, 7.5, 1.0]]).cuda()
    ref = torch.cat([cam] + [fn(cam * 2.0 ** k) for k in range(4) for fn in (torch.sin, torch.cos)], -1)
    assert (ops.fourier_embed(cam.contiguous()) - ref).abs().max() < 1e-4
    v = torch.randn(1000, generator=g).cuda() * 3
    assert _rel(ops.silu_to_bf16(v), F.silu(v)) < 1e-2
    a_ = _mk((64, 320), 1); b_ = _mk((64, 320), 2); c_ = _mk((64, 320), 3)
    assert _rel(ops.add_bf16(a_, b_, c_), a_.float() + b_.float() + c_.float()) < 1e-2


def test_box_features():
    from dualdiff_b200 import ops
    g = torch.Generator().manual_seed(0)
    n, P = 37, 8
    boxes = (torch.rand(n, P, 3, generator=g) * 100 - 50).cuda()
    classes = torch.randint(0, 10, (n,), generator=g).cuda()
    masks = (torch.rand(n, generator=g) < 0.6).cuda()
    tokens = torch.randn(10, 768, generator=g).cuda()
    null_pos = torch.randn(27 * P, generator=g).cuda(); null_cls = torch.randn(768, generator=g).cuda()
    pos = torch.empty(n, 27 * P, device="cuda"); cls = torch.empty(n, 768, device="cuda")
    ops.box_features(boxes, classes, masks, tokens, null_pos, null_cls, pos, cls)
    m = masks.float()[:, None]
    emb = torch.cat([boxes] + [fn(boxes * 2.0 ** k) for k in range(4) for fn in (torch.sin, torch.cos)], -1).reshape(n, -1)
    assert (pos - (emb * m + null_pos * (1 - m))).abs().max() < 1e-4
    assert (cls - (tokens[classes] * m + null_cls * (1 - m))).abs().max() < 1e-6


def test_box_features_padded_class_ids_follow_the_reference():
    """the reference's collate pads `classes` with -1 where masks == 0 (dataset/utils.py:243,283) and indexes
    class_tokens[-1] (a Python wrap-around) before multiplying by the mask: the kernel must neither read before the
    token table nor let whatever lies there leak in, an unmasked -k counts from the end, and ids outside
    [-n_classes, n_classes) raise like the reference's IndexError.  Checked against the oracle (torch indexing)."""
    from dualdiff_b200 import ops
    from oracle import dualdiff_oracle as O
    g = torch.Generator().manual_seed(5)
    R, L, P, n_cls = 3, 9, 8, 10
    boxes = torch.rand(R, L, P, 3, generator=g) * 100 - 50
    masks = torch.rand(R, L, generator=g) < 0.5
    classes = torch.randint(0, n_cls, (R, L), generator=g)
    classes[~masks] = -1                      # collate padding
    boxes[~masks] = 0
    classes[0, 0], masks[0, 0] = -3, True     # unmasked negative id: Python semantics = n_cls - 3
    sd = {"bbox_embedder._class_tokens": torch.randn(n_cls, 768, generator=g),
          "bbox_embedder.null_pos_feature": torch.randn(27 * P, generator=g),
          "bbox_embedder.null_class_feature": torch.randn(768, generator=g)}
    m = masks.reshape(-1, 1).float()
    ref_cls = sd["bbox_embedder._class_tokens"][classes.reshape(-1)] * m + sd["bbox_embedder.null_class_feature"][None] * (1 - m)
    ref_pos = O.fourier_embed(boxes.reshape(R * L, P, 3)).reshape(R * L, -1) * m + sd["bbox_embedder.null_pos_feature"][None] * (1 - m)
    # poison the memory in front of the token table: a stray read of class_tokens[-1] would pick up NaNs
    arena = torch.full((n_cls + 4, 768), float("nan"), device="cuda")
    arena[4:] = sd["bbox_embedder._class_tokens"].cuda()
    tokens = arena[4:]
    pos = torch.empty(R * L, 27 * P, device="cuda"); cls = torch.empty(R * L, 768, device="cuda")
    ops.box_features(boxes.reshape(R * L, P, 3).cuda(), classes.reshape(-1).cuda(), masks.reshape(-1).cuda(), tokens,
                     sd["bbox_embedder.null_pos_feature"].cuda(), sd["bbox_embedder.null_class_feature"].cuda(), pos, cls)
    assert torch.isfinite(cls).all() and torch.isfinite(pos).all()
    assert (cls.cpu() - ref_cls).abs().max() < 1e-6
    assert (pos.cpu() - ref_pos).abs().max() < 1e-4
    bad = classes.reshape(-1).clone(); bad[1] = n_cls
    with pytest.raises(IndexError):
        ops.box_features(boxes.reshape(R * L, P, 3).cuda(), bad.cuda(), masks.reshape(-1).cuda(), tokens,
                         sd["bbox_embedder.null_pos_feature"].cuda(), sd["bbox_embedder.null_class_feature"].cuda(), pos, cls)
    with pytest.raises(TypeError):
        ops.box_features(boxes.reshape(R * L, P, 3).cuda(), classes.reshape(-1).int().cuda(), masks.reshape(-1).cuda(), tokens,
                         sd["bbox_embedder.null_pos_feature"].cuda(), sd["bbox_embedder.null_class_feature"].cuda(), pos, cls)


def test_layout_roundtrip_and_cfg_sched():
    from dualdiff_b200 import ops
    g = torch.Generator().manual_seed(0)
    n, c, H, W = 6, 4, 28, 50
    x = torch.randn(n, c, H, W, generator=g).cuda()
    rows = ops.nchw_to_rows(x)
    assert torch.equal(rows.float().reshape(n, H, W, c).permute(0, 3, 1, 2), x.to(torch.bfloat16).float())
    back = ops.rows_to_nchw(rows, n, (H, W))
    assert torch.equal(back, x.to(torch.bfloat16).float())
    # cfg + linear-combination scheduler update
    eps = torch.randn(2 * n * H * W, c, generator=g).cuda()
    lat = torch.randn(n, c, H, W, generator=g).cuda(); last = torch.randn(n, c, H, W, generator=g).cuda()
    m0 = torch.randn(n, c, H, W, generator=g).cuda(); m1 = torch.randn(n, c, H, W, generator=g).cuda()
    coef = torch.tensor([2.0, 0.7, 1.3, 0.11, -0.2, 0.05, 0.3, 0.6, 0.9, -0.4, 0.25, 0, 0, 0, 0, 0]).cuda()
    e = eps.reshape(2, n, H, W, c).permute(0, 1, 4, 2, 3)
    eg = e[0] + 2.0 * (e[1] - e[0])
    x0 = (lat - 0.7 * eg) * 1.3
    xc = 0.6 * lat + 0.11 * last - 0.2 * m0 + 0.05 * m1 + 0.3 * x0
    xn = 0.9 * xc - 0.4 * x0 + 0.25 * m0
    m0_old = m0.clone()
    ops.cfg_sched_step(eps, lat, last, m0, m1, coef, n_img=n, c=c, hw=H * W, cfg=True)
    assert (lat - xn).abs().max() < 1e-5 and (last - xc).abs().max() < 1e-5
    assert (m0 - x0).abs().max() < 1e-5 and torch.equal(m1, m0_old)
