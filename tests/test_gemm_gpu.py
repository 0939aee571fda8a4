"""tcgen05 GEMM / implicit-GEMM conv parity (GPU). Reference = torch fp32 on the same bf16-rounded inputs.
Tolerance: bf16 output rounding -> max-abs <= 2e-2 * max|ref|."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(out, ref):
    return ((out.float() - ref).abs().max() / ref.abs().max().clamp_min(1e-6)).item()


def _mk(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).cuda()


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 128, 64, 128), (300, 320, 320, 0), (1400, 320, 320, 160), (1000, 640, 1280, 0),
    (257, 1280, 768, 256), (129, 960, 320, 192), (77, 64, 768, 0), (50, 32, 128, 0),
    (4096, 2560, 640, 0), (133, 200, 72, 0),
])
def test_linear(M, N, K, bn):
    from dualdiff_b200 import ops
    a = _mk((M, K), 1); w = _mk((N, K), 2, K ** -0.5)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(3)).cuda()
    out = ops.gemm(a, w, bias=bias, force_bn=bn)
    ref = a.float() @ w.float().t() + bias
    assert _rel(out, ref) < 1e-2, _rel(out, ref)


@pytest.mark.parametrize("M,N,K,res,tma", [(40000, 1152, 320, False, True), (40001, 320, 320, True, True), (30000, 320, 192, True, True),
                                           (40000, 640, 128, True, False), (25000, 704, 256, False, True)])
def test_short_k_contiguous_tile_ranges(M, N, K, res, tma):
    """K <= 320 with several n-tiles per m-tile: every worker walks a contiguous range of the tile order and keeps the A
    blocks of an m-tile in the ring for all of its n-tiles (GemmDev::areuse); several tiles per worker, ragged M / N tails,
    both epilogues.  Reference = torch fp32 matmul on the GPU over ALL rows."""
    from dualdiff_b200 import ops
    a = _mk((M, K), 1); w = _mk((N, K), 2, K ** -0.5)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(3)).cuda()
    r1 = _mk((M, N), 4) if res else None
    out = ops.gemm(a, w, bias=bias, res1=r1, no_tma_epilogue=not tma)
    ref = a.float() @ w.float().t() + bias + (r1.float() if res else 0)
    assert _rel(out, ref) < 1e-2, _rel(out, ref)
    again = ops.gemm(a, w, bias=bias, res1=r1, no_tma_epilogue=not tma)
    assert torch.equal(out, again)


def test_epilogue_residual_rowvec_f32():
    from dualdiff_b200 import ops
    n_img, T, K, N = 3, 350, 640, 640
    a = _mk((n_img * T, K), 1); w = _mk((N, K), 2, K ** -0.5)
    r1 = _mk((n_img * T, N), 4); r2 = _mk((n_img * T, N), 5)
    rv = torch.randn(n_img, N, generator=torch.Generator().manual_seed(6)).cuda()
    ref = a.float() @ w.float().t() + r1.float() + r2.float() + rv.repeat_interleave(T, 0)
    out = ops.gemm(a, w, res1=r1, res2=r2, rowvec=rv, rows_per_img=T)
    assert _rel(out, ref) < 1e-2
    out32 = ops.gemm(a, w, res1=r1, res2=r2, rowvec=rv, rows_per_img=T, out_f32=True)
    assert out32.dtype == torch.float32 and _rel(out32, ref) < 2e-3


def test_dual_source_concat():
    from dualdiff_b200 import ops
    M, K1, K2, N = 700, 640, 320, 320
    a1 = _mk((M, K1), 1); a2 = _mk((M, K2), 2); w = _mk((N, K1 + K2), 3, (K1 + K2) ** -0.5)
    out = ops.gemm(a1, w, a2=a2)
    ref = torch.cat([a1, a2], 1).float() @ w.float().t()
    assert _rel(out, ref) < 1e-2


@pytest.mark.parametrize("M,C", [(500, 320), (30000, 320), (9000, 640)])
def test_geglu(M, C):
    from dualdiff_b200 import ops, packing
    a = _mk((M, C), 1); w = _mk((8 * C, C), 2, C ** -0.5)
    b = torch.randn(8 * C, generator=torch.Generator().manual_seed(3)).cuda()
    wp, bp = packing.pack_geglu(w, b)
    out = ops.gemm(a, wp, bias=bp, geglu=True)
    proj = a.float() @ w.float().t() + b
    v, g = proj.chunk(2, -1)
    ref = v * F.gelu(g)
    assert out.shape == (M, 4 * C) and _rel(out, ref) < 1e-2


@pytest.mark.parametrize("n,H,W,ci,co", [(2, 28, 50, 320, 320), (3, 14, 25, 640, 1280), (2, 7, 13, 1280, 1280),
                                         (5, 4, 7, 128, 64), (1, 28, 50, 320, 4), (4, 14, 25, 1280, 640), (12, 28, 50, 640, 320),
                                         (30, 28, 50, 320, 320), (2, 9, 11, 72, 128)])
def test_conv3x3_implicit_gemm(n, H, W, ci, co):
    from dualdiff_b200 import ops, packing
    x = _mk((n, H, W, ci), 1)
    w = _mk((co, ci, 3, 3), 2, (9 * ci) ** -0.5)
    bias = torch.randn(co, generator=torch.Generator().manual_seed(3)).cuda()
    out = ops.gemm(packing.to_padded(x), packing.pack_conv3x3(w), bias=bias, taps=9, conv_hw=(H, W), n_img=n)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(n * H * W, co)
    assert _rel(out, ref) < 1e-2, _rel(out, ref)


def test_conv3x3_dual_source_with_time_vector_and_residual():
    """the up-block resnets: conv over the channel concat of two padded sources (skip connection) with the per-image time
    vector and a residual in the epilogue; 160-wide tiles (one ring slot per kernel row and 64 channels)"""
    from dualdiff_b200 import ops, packing
    n, H, W, c1, c2, co = 7, 28, 50, 640, 320, 320
    x1 = _mk((n, H, W, c1), 1); x2 = _mk((n, H, W, c2), 2)
    w = _mk((co, c1 + c2, 3, 3), 3, (9 * (c1 + c2)) ** -0.5)
    bias = torch.randn(co, generator=torch.Generator().manual_seed(4)).cuda()
    rv = torch.randn(n, co, generator=torch.Generator().manual_seed(5)).cuda()
    r1 = _mk((n * H * W, co), 6)
    out = ops.gemm(packing.to_padded(x1), packing.pack_conv3x3(w), a2=packing.to_padded(x2), bias=bias, rowvec=rv,
                   rows_per_img=H * W, res1=r1, taps=9, conv_hw=(H, W), n_img=n)
    x = torch.cat([x1, x2], -1).float().permute(0, 3, 1, 2)
    conv = F.conv2d(x, w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(n * H * W, co) + r1.float()
    assert _rel(out, conv + rv.repeat_interleave(H * W, 0)) < 1e-2
    # ONE time vector shared by all images (every image of a sampler step has the same timestep: rows_per_img = n * H * W)
    out1 = ops.gemm(packing.to_padded(x1), packing.pack_conv3x3(w), a2=packing.to_padded(x2), bias=bias, rowvec=rv[:1],
                    rows_per_img=n * H * W, res1=r1, taps=9, conv_hw=(H, W), n_img=n)
    assert _rel(out1, conv + rv[:1]) < 1e-2


@pytest.mark.parametrize("M,N,K,res,act", [(134400 // 8, 320, 320, True, 0), (1000, 1088, 320, False, 0), (4200, 640, 640, True, 1),
                                           (257, 1280, 1280, True, 0), (130, 384, 320, False, 0), (5000, 72, 320, True, 0)])
def test_tma_epilogue_equals_register_epilogue(M, N, K, res, act):
    """plain bf16 GEMMs take the TMA load/store epilogue; it must agree with the register path bit-for-bit-ish"""
    from dualdiff_b200 import ops
    a = _mk((M, K), 1); w = _mk((N, K), 2, K ** -0.5)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(3)).cuda()
    r1 = _mk((M, N), 4) if res else None
    out_tma = ops.gemm(a, w, bias=bias, res1=r1, act=act)
    out_reg = ops.gemm(a, w, bias=bias, res1=r1, act=act, no_tma_epilogue=True)
    ref = a.float() @ w.float().t() + bias + (r1.float() if res else 0)
    if act:
        ref = F.silu(ref)
    assert _rel(out_tma, ref) < 1e-2 and _rel(out_reg, ref) < 1e-2
    assert (out_tma.float() - out_reg.float()).abs().max() <= 2e-2 * ref.abs().max()


@pytest.mark.parametrize("M,N,K,taps", [(1000, 1280, 1280, 1), (129, 320, 320, 1), (5000, 640, 2560, 1), (2 * 29 * 51, 320, 320, 9)])
def test_cta_pair_kernel_equals_single_cta_kernel(M, N, K, taps):
    """default = CTA pairs (tcgen05 cta_group::2, 256-row tiles); one_cta=True = the single-CTA kernel"""
    from dualdiff_b200 import ops, packing
    bias = torch.randn(N, generator=torch.Generator().manual_seed(3)).cuda()
    if taps == 9:
        x = _mk((2, 28, 50, K), 1); w = _mk((N, K, 3, 3), 2, (9 * K) ** -0.5)
        a, wp = packing.to_padded(x), packing.pack_conv3x3(w)
        kw = dict(taps=9, conv_hw=(28, 50), n_img=2)
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, N)
    else:
        a = _mk((M, K), 1); wp = _mk((N, K), 2, K ** -0.5); kw = {}
        ref = a.float() @ wp.float().t() + bias
    o2 = ops.gemm(a, wp, bias=bias, stream_k=-1, **kw)
    o1 = ops.gemm(a, wp, bias=bias, one_cta=True, stream_k=-1, **kw)
    assert _rel(o2, ref) < 1e-2 and _rel(o1, ref) < 1e-2
    assert torch.equal(o1, o2)   # same accumulation order per output element -> bit-identical


@pytest.mark.parametrize("case", ["conv_l3", "conv_small", "linear_res", "linear_rowvec_f32", "one_cta", "dual"])
def test_stream_k_equals_classic(case):
    """stream-K (flattened (tile, k-iteration) space cut evenly over the CTA pairs, fp32 partial tiles summed in worker
    order by a second pass that applies the epilogue) against the one-tile-per-worker kernel and torch fp32"""
    from dualdiff_b200 import ops, packing
    kw, kw_ref = {}, {}
    if case in ("conv_l3", "conv_small"):
        n, H, W, ci, co = (24, 4, 7, 1280, 1280) if case == "conv_l3" else (3, 7, 13, 192, 200)
        x = _mk((n, H, W, ci), 1); w = _mk((co, ci, 3, 3), 2, (9 * ci) ** -0.5)
        bias = torch.randn(co, generator=torch.Generator().manual_seed(3)).cuda()
        rv = torch.randn(n, co, generator=torch.Generator().manual_seed(6)).cuda()
        r1 = _mk((n * H * W, co), 4)
        a, wp = packing.to_padded(x), packing.pack_conv3x3(w)
        kw = dict(taps=9, conv_hw=(H, W), n_img=n, bias=bias, rowvec=rv, rows_per_img=H * W, res1=r1)
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, co)
        ref = ref + rv.repeat_interleave(H * W, 0) + r1.float()
    elif case == "dual":
        M, K1, K2, N = 700, 640, 320, 320
        a = _mk((M, K1), 1); a2 = _mk((M, K2), 2); wp = _mk((N, K1 + K2), 3, (K1 + K2) ** -0.5)
        kw = dict(a2=a2)
        ref = torch.cat([a, a2], 1).float() @ wp.float().t()
    else:
        M, N, K = (2688, 1280, 5120) if case != "one_cta" else (100, 328, 2560)
        a = _mk((M, K), 1); wp = _mk((N, K), 2, K ** -0.5)
        bias = torch.randn(N, generator=torch.Generator().manual_seed(3)).cuda()
        r1 = _mk((M, N), 4); r2 = _mk((M, N), 5)
        ref = a.float() @ wp.float().t() + bias + r1.float() + r2.float()
        kw = dict(bias=bias, res1=r1, res2=r2)
        if case == "linear_rowvec_f32":
            rv = torch.randn(M // 28, N, generator=torch.Generator().manual_seed(6)).cuda()
            kw.update(rowvec=rv, rows_per_img=28, out_f32=True, act=1)
            ref = F.silu(ref + rv.repeat_interleave(28, 0))
    o_sk = ops.gemm(a, wp, stream_k=1, **kw)
    o_cl = ops.gemm(a, wp, stream_k=-1, **kw)
    assert _rel(o_sk, ref) < 1e-2 and _rel(o_cl, ref) < 1e-2, (_rel(o_sk, ref), _rel(o_cl, ref))
    assert (o_sk.float() - o_cl.float()).abs().max() <= 1e-2 * ref.abs().max()
    assert torch.equal(o_sk, ops.gemm(a, wp, stream_k=1, **kw))   # fixed summation order -> bit-reproducible
