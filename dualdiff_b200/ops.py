"""Thin torch-tensor wrappers over the C ABI (one Python function per ``dd_*`` entry point).

PyTorch is plumbing here: it owns device memory and the current stream; all arithmetic happens in
libdualdiff_sm100.so.  Every wrapper raises if the tensor is not a CUDA tensor — no CPU fallback.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import GemmArgs, check


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


# ---- optional per-launch instrumentation (bench.py roofline pass; never active on the timed path) ----
_PROF = None


def profile_start():
    global _PROF
    _PROF = []


def profile_stop():
    """-> list of (kernel family, milliseconds, algorithmic flops, algorithmic bytes, tag)"""
    global _PROF
    rec, _PROF = _PROF, None
    torch.cuda.synchronize()
    return [(n, a.elapsed_time(b), f, by, tag) for (n, a, b, f, by, tag) in rec]


class _Rec:
    def __init__(self, name, flops=0.0, nbytes=0.0, tag=""):
        self.args = (name, flops, nbytes, tag)

    def __enter__(self):
        if _PROF is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if _PROF is not None:
            self.b.record()
            n, f, by, tag = self.args
            _PROF.append((n, self.a, self.b, f, by, tag))
        return False


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t, dtype, name):
    if not t.is_cuda:
        raise _lib.DDError(f"{name}: CUDA tensor required (dualdiff_b200 has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")


def padded_rows(n_img, H, W):
    """rows of the zero-haloed pixel layout [img][H+1][W+1] used by the 3x3 implicit-GEMM conv"""
    return n_img * (H + 1) * (W + 1)


_WORKSPACE = {}
_WORKSPACE_BYTES = 64 << 20


def gemm_workspace(device):
    """fp32 scratch for stream-K partial tiles: one buffer per (device, stream) so that the concurrently running
    ControlNet / UNet streams never share one.  Allocated on first use, kept for the life of the process."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _WORKSPACE.get(key)
    if ws is None:
        ws = _WORKSPACE[key] = torch.empty(_WORKSPACE_BYTES, device=device, dtype=torch.uint8)
    return ws


def gemm(a, w, *, out=None, bias=None, rowvec=None, rows_per_img=1, res1=None, res2=None, a2=None,
         taps=1, conv_hw=None, n_img=None, geglu=False, out_f32=False, force_bn=0, act=0, no_tma_epilogue=False, one_cta=False,
         stream_k=0):
    """out = epilogue(A @ W^T).  a: [M, K] bf16 (row stride may exceed K); w: [N, taps*K] bf16.

    taps=9: ``a`` is the padded-pixel activation [n_img*(H+1)*(W+1), K]; the result has n_img*H*W rows.
    """
    _req(a, torch.bfloat16, "a")
    _req(w, torch.bfloat16, "w")
    M, K1 = a.shape
    K = K1 + (a2.shape[1] if a2 is not None else 0)
    N = w.shape[0]
    assert w.shape[1] == taps * K, (w.shape, taps, K)
    assert a.stride(1) == 1 and w.stride(1) == 1
    if taps == 9:
        H, W = conv_hw
        assert M == padded_rows(n_img, H, W), (M, n_img, H, W)
        rows_out = n_img * H * W
    else:
        H = W = 0
        rows_out = M
    n_store = N // 2 if geglu else N
    if out is None:
        out = torch.empty((rows_out, n_store), device=a.device,
                          dtype=torch.float32 if out_f32 else torch.bfloat16)
    assert out.shape[0] == rows_out and out.shape[1] == n_store and out.stride(1) == 1
    args = GemmArgs()
    args.a = _ptr(a); args.a2 = _ptr(a2); args.w = _ptr(w); args.out = _ptr(out)
    args.bias = _ptr(bias); args.rowvec = _ptr(rowvec); args.res1 = _ptr(res1); args.res2 = _ptr(res2)
    args.M = M; args.N = N; args.K = K; args.k1 = K1 if a2 is not None else 0
    args.taps = taps; args.conv_h = H; args.conv_w = W
    args.a_ld = a.stride(0); args.a2_ld = a2.stride(0) if a2 is not None else 0
    args.w_ld = w.stride(0); args.out_ld = out.stride(0)
    args.res1_ld = res1.stride(0) if res1 is not None else 0
    args.res2_ld = res2.stride(0) if res2 is not None else 0
    args.rowvec_ld = rowvec.stride(0) if rowvec is not None else 0
    args.rows_per_img = rows_per_img
    args.out_f32 = 1 if out.dtype == torch.float32 else 0
    args.geglu = 1 if geglu else 0
    args.force_bn = force_bn
    args.act = act
    args.no_tma_epilogue = 1 if no_tma_epilogue else 0
    args.one_cta = 1 if one_cta else 0
    args.stream_k = stream_k
    if args.stream_k >= 0 and not geglu:
        ws = gemm_workspace(a.device)
        args.workspace = _ptr(ws); args.workspace_bytes = ws.numel()
    if bias is not None:
        _req(bias, torch.float32, "bias")
    if rowvec is not None:
        _req(rowvec, torch.float32, "rowvec")
    with _Rec("gemm_tcgen05", 2.0 * rows_out * N * K * taps,
              2.0 * (rows_out * K * (1 if taps == 1 else 1) + N * K * taps + rows_out * n_store),
              f"M{rows_out}_N{N}_K{K * taps}"):
        check(_lib.lib().dd_gemm(C.byref(args), _stream()), "dd_gemm")
    return out


# ---------------------------------------------------------------------------------------------------
from ._lib import (AttentionArgs, GroupNormArgs, LayerNormArgs, LinearF32Args, SeqAttentionArgs,  # noqa: E402
                   TemporalAttentionArgs, ToPaddedArgs)

_L = C.c_longlong
_I = C.c_int


def groupnorm(x1, gamma, beta, *, n_img, hw, x2=None, eps=1e-5, silu=True, padded_out=False, groups=32,
              out=None, stats=None, two_pass=False):
    """GroupNorm(32) [+SiLU] over compact channels-last rows; x2 = second channel source (skip concat)."""
    _req(x1, torch.bfloat16, "x1")
    H, W = hw
    c1 = x1.shape[1]
    c2 = x2.shape[1] if x2 is not None else 0
    C_ = c1 + c2
    assert x1.shape[0] == n_img * H * W
    rows = padded_rows(n_img, H, W) if padded_out else n_img * H * W
    if out is None:
        out = torch.empty((rows, C_), device=x1.device, dtype=torch.bfloat16)
    if stats is None:
        stats = torch.empty((_lib.lib().dd_groupnorm_scratch_floats(n_img, C_, H * W),), device=x1.device,
                            dtype=torch.float32)
    a = GroupNormArgs()
    a.x1 = _ptr(x1); a.x2 = _ptr(x2); a.out = _ptr(out); a.stats = _ptr(stats)
    a.gamma = _ptr(gamma); a.beta = _ptr(beta)
    a.x1_ld = x1.stride(0); a.x2_ld = x2.stride(0) if x2 is not None else 0; a.out_ld = out.stride(0)
    a.n_img = n_img; a.h = H; a.w = W; a.c1 = c1; a.c2 = c2; a.groups = groups
    a.eps = eps; a.silu = 1 if silu else 0; a.padded_out = 1 if padded_out else 0; a.two_pass = 1 if two_pass else 0
    with _Rec("groupnorm_silu", 0.0, 2.0 * (2 * n_img * H * W * C_ + rows * C_), f"C{C_}_HW{H * W}"):
        check(_lib.lib().dd_groupnorm(C.byref(a), _stream()), "dd_groupnorm")
    return out


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    _req(x, torch.bfloat16, "x")
    rows, c = x.shape
    if out is None:
        out = torch.empty((rows, c), device=x.device, dtype=torch.bfloat16)
    a = LayerNormArgs()
    a.x = _ptr(x); a.out = _ptr(out); a.gamma = _ptr(gamma); a.beta = _ptr(beta)
    a.x_ld = x.stride(0); a.out_ld = out.stride(0); a.rows = rows; a.c = c; a.eps = eps
    with _Rec("layernorm", 0.0, 4.0 * rows * c, f"C{c}"):
        check(_lib.lib().dd_layernorm(C.byref(a), _stream()), "dd_layernorm")
    return out


def attention(q, k, v, *, n_img, lq, lk, heads, head_dim, out=None, q_col0=0, k_col0=0, v_col0=0,
              q_hs=None, k_hs=None, v_hs=None, kv_map=None, n_src=1, n_kv_img=None, scale=None,
              q_cols=None, k_cols=None, v_cols=None, variant=0, v_ones=False, concat=False):
    """q: [n_img*lq, *], k/v: [n_kv_img*lk, *] bf16 (may be column views of one fused projection output).
    variant: testing hook of the head_dim-40 kernel (include/dualdiff_b200.h), 0 = auto.
    v_ones (head_dim 40): V heads have a 48-column stride with 1.0 in column 40 -- the softmax denominator comes out of the
    P V product."""
    _req(q, torch.bfloat16, "q")
    hs_qk = 48 if head_dim == 40 else head_dim
    q_hs = hs_qk if q_hs is None else q_hs
    k_hs = hs_qk if k_hs is None else k_hs
    if v_ones and head_dim != 40:
        raise ValueError("v_ones is a head_dim-40 layout")
    v_hs = (48 if v_ones else head_dim) if v_hs is None else v_hs
    n_kv_img = n_img if n_kv_img is None else n_kv_img
    if out is None:
        out = torch.empty((n_img * lq, heads * head_dim), device=q.device, dtype=torch.bfloat16)
    a = AttentionArgs()
    a.q = _ptr(q); a.k = _ptr(k); a.v = _ptr(v); a.out = _ptr(out); a.kv_map = _ptr(kv_map)
    a.q_ld = q.stride(0); a.k_ld = k.stride(0); a.v_ld = v.stride(0); a.out_ld = out.stride(0)
    a.q_cols = q.shape[1] if q_cols is None else q_cols
    a.k_cols = k.shape[1] if k_cols is None else k_cols
    a.v_cols = v.shape[1] if v_cols is None else v_cols
    a.q_col0 = q_col0; a.k_col0 = k_col0; a.v_col0 = v_col0
    a.q_head_stride = q_hs; a.k_head_stride = k_hs; a.v_head_stride = v_hs
    a.n_img = n_img; a.n_kv_img = n_kv_img; a.heads = heads; a.head_dim = head_dim
    a.lq = lq; a.lk = lk; a.n_src = n_src
    a.scale = float(head_dim) ** -0.5 if scale is None else scale
    a.variant = variant
    a.v_ones = 1 if v_ones else 0
    a.concat = 1 if concat else 0     # n_src > 1: one softmax over the concatenated sources instead of a sum of per-source attentions
    with _Rec("attn_tcgen05", 4.0 * n_img * lq * lk * heads * head_dim * n_src,
              2.0 * heads * head_dim * (2 * n_img * lq + 2 * n_kv_img * lk), f"d{head_dim}_Lq{lq}_Lk{lk}_s{n_src}"):
        check(_lib.lib().dd_attention(C.byref(a), _stream()), "dd_attention")
    return out


def temporal_attention(q, k, v, *, n_outer, n_view, tokens, heads, head_dim, frames_q, frames_kv=None,
                       frames_per_rank=None, kv_rank_stride=0, out=None, q_col0=0, k_col0=0, v_col0=0,
                       q_hs=None, k_hs=None, v_hs=None, scale=None, frames_q_per_rank=None, q_rank_stride=0):
    """attention over the frames of a clip at every (outer, view, token).  q: [n_outer*frames_q*n_view*tokens, *];
    k/v: frames_kv frames as blocks of [n_outer, frames_per_rank, n_view] images (kv_rank_stride images apart); with
    frames_q_per_rank / q_rank_stride the queries and the output are laid out in rank blocks the same way."""
    _req(q, torch.bfloat16, "q")
    _req(k, torch.bfloat16, "k")
    hs_qk = 48 if head_dim == 40 else head_dim
    frames_kv = frames_q if frames_kv is None else frames_kv
    frames_per_rank = frames_kv if frames_per_rank is None else frames_per_rank
    n_img = n_outer * frames_q * n_view
    if out is None:
        out = torch.empty((n_img * tokens, heads * head_dim), device=q.device, dtype=torch.bfloat16)
    a = TemporalAttentionArgs()
    a.q = _ptr(q); a.k = _ptr(k); a.v = _ptr(v); a.out = _ptr(out)
    a.q_ld = q.stride(0); a.k_ld = k.stride(0); a.v_ld = v.stride(0); a.out_ld = out.stride(0)
    a.q_col0 = q_col0; a.k_col0 = k_col0; a.v_col0 = v_col0
    a.q_head_stride = hs_qk if q_hs is None else q_hs
    a.k_head_stride = hs_qk if k_hs is None else k_hs
    a.v_head_stride = head_dim if v_hs is None else v_hs
    a.n_outer = n_outer; a.n_view = n_view; a.tokens = tokens; a.heads = heads; a.head_dim = head_dim
    a.frames_q = frames_q; a.frames_kv = frames_kv; a.frames_per_rank = frames_per_rank
    a.kv_rank_stride = kv_rank_stride
    a.frames_q_per_rank = 0 if frames_q_per_rank is None else frames_q_per_rank
    a.q_rank_stride = q_rank_stride
    a.scale = float(head_dim) ** -0.5 if scale is None else scale
    nb = 2.0 * heads * head_dim * tokens * n_outer * n_view * (2 * frames_q + 2 * frames_kv)
    with _Rec("temporal_attn", 0.0, nb, f"d{head_dim}_F{frames_q}of{frames_kv}_T{tokens}"):
        check(_lib.lib().dd_temporal_attention(C.byref(a), _stream()), "dd_temporal_attention")
    return out


def ors_project(origins, dirs, semantics, *, sample_point, sample_step=0.2, want_ids=True, want_rows=False,
                keep_fg=True, keep_bg=True):
    """origins / dirs: fp32 [n_pix, 3]; semantics: uint8 [D, H, W] -> (ids uint8 [n_pix, S] | None, rows bf16 [n_pix, S] | None)"""
    _req(origins, torch.float32, "origins")
    _req(dirs, torch.float32, "dirs")
    _req(semantics, torch.uint8, "semantics")
    assert origins.is_contiguous() and dirs.is_contiguous() and semantics.is_contiguous()
    n_pix = origins.shape[0]
    D, H, W = semantics.shape
    ids = torch.empty((n_pix, sample_point), device=origins.device, dtype=torch.uint8) if want_ids else None
    rows = torch.empty((n_pix, sample_point), device=origins.device, dtype=torch.bfloat16) if want_rows else None
    with _Rec("ors_project", 0.0, n_pix * sample_point * ((1 if want_ids else 0) + (2 if want_rows else 0)) + n_pix * 24,
              f"pix{n_pix}_S{sample_point}"):
        check(_lib.lib().dd_ors_project(_ptr(origins), _ptr(dirs), _ptr(semantics), _ptr(ids), _ptr(rows),
                                        _L(n_pix), _I(sample_point), C.c_float(sample_step), _I(D), _I(H), _I(W),
                                        _I(1 if keep_fg else 0), _I(1 if keep_bg else 0), _stream()), "dd_ors_project")
    return ids, rows


def nchw_to_padded(src, *, n_outer, n_view, c, h, w, cp, stride_outer, stride_view, stride_c, stride_h, out=None):
    assert src.is_cuda and src.dtype in (torch.float32, torch.bfloat16)
    n = n_outer * n_view
    if out is None:
        out = torch.empty((padded_rows(n, h, w), cp), device=src.device, dtype=torch.bfloat16)
    a = ToPaddedArgs()
    a.src = _ptr(src); a.out = _ptr(out)
    a.stride_outer = stride_outer; a.stride_view = stride_view; a.stride_c = stride_c; a.stride_h = stride_h
    a.n_outer = n_outer; a.n_view = n_view; a.c = c; a.h = h; a.w = w; a.cp = cp
    a.src_f32 = 1 if src.dtype == torch.float32 else 0
    check(_lib.lib().dd_nchw_to_padded(C.byref(a), _stream()), "dd_nchw_to_padded")
    return out


def nchw_patches(src, *, n_outer, n_view, c, h, w, cp, stride_outer, stride_view, stride_c, stride_h, out=None):
    """3x3 stride-1 pad-1 patch rows [n*h*w, cp] (tap-major, channel-minor, zero tail) of an NCHW image: conv_in on the
    4-channel latents as ONE K = 40 GEMM (nine 8-channel taps through the conv path ran at 35 TFLOP/s)"""
    assert src.is_cuda and src.dtype in (torch.float32, torch.bfloat16)
    n = n_outer * n_view
    if out is None:
        out = torch.empty((n * h * w, cp), device=src.device, dtype=torch.bfloat16)
    a = ToPaddedArgs()
    a.src = _ptr(src); a.out = _ptr(out)
    a.stride_outer = stride_outer; a.stride_view = stride_view; a.stride_c = stride_c; a.stride_h = stride_h
    a.n_outer = n_outer; a.n_view = n_view; a.c = c; a.h = h; a.w = w; a.cp = cp
    a.src_f32 = 1 if src.dtype == torch.float32 else 0
    with _Rec("layout", 0.0, 4.0 * n * c * h * w + 2.0 * out.numel()):
        check(_lib.lib().dd_nchw_patches(C.byref(a), _stream()), "dd_nchw_patches")
    return out


def im2col_s2(x, *, n_img, hw, out=None):
    _req(x, torch.bfloat16, "x")
    H, W = hw
    c = x.shape[1]
    ho, wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    if out is None:
        out = torch.empty((n_img * ho * wo, 9 * c), device=x.device, dtype=torch.bfloat16)
    with _Rec("layout", 0.0, 2.0 * (n_img * H * W * c + out.numel())):
        check(_lib.lib().dd_im2col_s2(_ptr(x), _L(x.stride(0)), _ptr(out), n_img, H, W, c, _stream()), "dd_im2col_s2")
    return out, (ho, wo)


def upsample_pad(x, *, n_img, hw, hw2, out=None):
    _req(x, torch.bfloat16, "x")
    H, W = hw
    H2, W2 = hw2
    c = x.shape[1]
    if out is None:
        out = torch.empty((padded_rows(n_img, H2, W2), c), device=x.device, dtype=torch.bfloat16)
    with _Rec("layout", 0.0, 2.0 * (n_img * H * W * c + out.numel())):
        check(_lib.lib().dd_upsample_pad(_ptr(x), _L(x.stride(0)), _ptr(out), n_img, H, W, c, H2, W2, _stream()),
              "dd_upsample_pad")
    return out


def pad_rows(x, *, n_img, hw, out=None):
    return upsample_pad(x, n_img=n_img, hw=hw, hw2=hw, out=out)


def linear_f32(x, w, b=None, *, act=0, out=None, out16=None, want_f32=True):
    _req(x, torch.float32, "x")
    _req(w, torch.float32, "w")
    M, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K and w.is_contiguous()
    if out is None and want_f32:
        out = torch.empty((M, N), device=x.device, dtype=torch.float32)
    a = LinearF32Args()
    a.x = _ptr(x); a.w = _ptr(w); a.b = _ptr(b); a.y = _ptr(out); a.y16 = _ptr(out16)
    a.x_ld = x.stride(0); a.y_ld = out.stride(0) if out is not None else 0
    a.y16_ld = out16.stride(0) if out16 is not None else 0
    a.M = M; a.N = N; a.K = K; a.act = act
    check(_lib.lib().dd_linear_f32(C.byref(a), _stream()), "dd_linear_f32")
    return out if out is not None else out16


def timestep_embedding(t, dim):
    _req(t, torch.float32, "t")
    out = torch.empty((t.shape[0], dim), device=t.device, dtype=torch.float32)
    check(_lib.lib().dd_timestep_embedding(_ptr(t), _ptr(out), t.shape[0], dim, _stream()), "dd_timestep_embedding")
    return out


def fourier_embed(x, nfreq=4):
    _req(x, torch.float32, "x")
    assert x.shape[-1] == 3 and x.is_contiguous()
    rows = x.numel() // 3
    out = torch.empty((rows, 3 + 6 * nfreq), device=x.device, dtype=torch.float32)
    check(_lib.lib().dd_fourier_embed(_ptr(x), _ptr(out), _L(rows), nfreq, _stream()), "dd_fourier_embed")
    return out


def box_features(boxes, classes, masks, class_tokens, null_pos, null_cls, pos_out, cls_out):
    """boxes [n, P, 3] fp32, classes [n] int64, masks [n] uint8/bool -> pos_out [n, 27P], cls_out [n, 768].
    Class ids follow the reference's Python indexing of `class_tokens` (bbox_embedder.py:189): ids in [-n_classes, 0)
    count from the end (the collate pads with -1, dataset/utils.py:243,283), anything else raises like the reference's
    IndexError.  Runs once per sample in prepare(), so the range check's host sync is off the step path."""
    _req(boxes, torch.float32, "boxes")
    _req(classes, torch.int64, "classes")
    _req(class_tokens, torch.float32, "class_tokens")
    n, P = boxes.shape[0], boxes.shape[1]
    n_classes = class_tokens.shape[0]
    assert classes.numel() == n and masks.numel() == n and classes.is_contiguous() and class_tokens.is_contiguous()
    if n and bool(((classes < -n_classes) | (classes >= n_classes)).any()):
        raise IndexError(f"box_features: class id outside [-{n_classes}, {n_classes}) (index out of range for class_tokens)")
    m8 = masks.to(torch.uint8).contiguous()
    check(_lib.lib().dd_box_features(_ptr(boxes), _ptr(classes), _ptr(m8), _ptr(class_tokens), _ptr(null_pos),
                                     _ptr(null_cls), _ptr(pos_out), _L(pos_out.stride(0)), _ptr(cls_out),
                                     _L(cls_out.stride(0)), _L(n), P, class_tokens.shape[1], _I(n_classes), _stream()),
          "dd_box_features")


def silu_to_bf16(x, out=None):
    _req(x, torch.float32, "x")
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    check(_lib.lib().dd_silu_to_bf16(_ptr(x), _ptr(out), _L(x.numel()), _stream()), "dd_silu_to_bf16")
    return out


def add_bf16(a, b, c=None, out=None):
    _req(a, torch.bfloat16, "a")
    if out is None:
        out = torch.empty_like(a)
    with _Rec("elementwise", 0.0, 2.0 * a.numel() * (3 if c is None else 4)):
        check(_lib.lib().dd_add_bf16(_ptr(a), _ptr(b), _ptr(c), _ptr(out), _L(a.numel()), _stream()), "dd_add_bf16")
    return out


def softmax_rows(x, scale, out=None):
    """fp32 scores [rows, cols] -> bf16 softmax(x * scale) along the columns"""
    _req(x, torch.float32, "x")
    rows, cols = x.shape
    if out is None:
        out = torch.empty((rows, cols), device=x.device, dtype=torch.bfloat16)
    with _Rec("elementwise", 0.0, 6.0 * rows * cols):
        check(_lib.lib().dd_softmax_rows(_ptr(x), _L(x.stride(0)), _ptr(out), _L(out.stride(0)), _I(rows), _I(cols),
                                         C.c_float(scale), _stream()), "dd_softmax_rows")
    return out


def clip_embed(ids, tok_emb, pos_emb, out=None):
    """CLIPTextEmbeddings: ids int64 [n_seq, L] -> bf16 rows [n_seq*L, C] = tok_emb[ids] + pos_emb[position]"""
    _req(ids, torch.int64, "ids")
    _req(tok_emb, torch.float32, "tok_emb")
    _req(pos_emb, torch.float32, "pos_emb")
    assert ids.is_contiguous() and tok_emb.is_contiguous() and pos_emb.is_contiguous()
    n_seq, L = ids.shape
    vocab, c = tok_emb.shape
    if L > pos_emb.shape[0] or pos_emb.shape[1] != c:
        raise ValueError(f"sequence length {L} exceeds the {pos_emb.shape[0]} position embeddings")
    if out is None:
        out = torch.empty((n_seq * L, c), device=ids.device, dtype=torch.bfloat16)
    with _Rec("elementwise", 0.0, n_seq * L * (8.0 + 10.0 * c)):
        check(_lib.lib().dd_clip_embed(_ptr(ids), _ptr(tok_emb), _ptr(pos_emb), _ptr(out), _L(out.stride(0)),
                                       _L(n_seq * L), _I(L), _I(c), _I(vocab), _stream()), "dd_clip_embed")
    return out


def seq_attention(q, k, v, *, n_seq, seq_len, heads, head_dim, causal=True, out=None, q_col0=0, k_col0=0, v_col0=0,
                  q_hs=None, k_hs=None, v_hs=None, scale=None):
    """attention over short sequences (seq_len <= 128, head_dim 64) with an optional causal mask.
    q/k/v: bf16 [n_seq*seq_len, *] (may be column views of one fused projection output)."""
    _req(q, torch.bfloat16, "q")
    _req(k, torch.bfloat16, "k")
    _req(v, torch.bfloat16, "v")
    if out is None:
        out = torch.empty((n_seq * seq_len, heads * head_dim), device=q.device, dtype=torch.bfloat16)
    a = SeqAttentionArgs()
    a.q = _ptr(q); a.k = _ptr(k); a.v = _ptr(v); a.out = _ptr(out)
    a.q_ld = q.stride(0); a.k_ld = k.stride(0); a.v_ld = v.stride(0); a.out_ld = out.stride(0)
    a.q_col0 = q_col0; a.k_col0 = k_col0; a.v_col0 = v_col0
    a.q_head_stride = head_dim if q_hs is None else q_hs
    a.k_head_stride = head_dim if k_hs is None else k_hs
    a.v_head_stride = head_dim if v_hs is None else v_hs
    a.n_seq = n_seq; a.seq_len = seq_len; a.heads = heads; a.head_dim = head_dim
    a.causal = 1 if causal else 0
    a.scale = float(head_dim) ** -0.5 if scale is None else scale
    with _Rec("seq_attention", 4.0 * n_seq * seq_len * seq_len * heads * head_dim * (0.5 if causal else 1.0),
              8.0 * n_seq * seq_len * heads * head_dim, f"d{head_dim}_L{seq_len}"):
        check(_lib.lib().dd_seq_attention(C.byref(a), _stream()), "dd_seq_attention")
    return out


def quick_gelu(x, out=None):
    """x * sigmoid(1.702 x) over bf16 (in place when out is x)"""
    _req(x, torch.bfloat16, "x")
    assert x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    with _Rec("elementwise", 0.0, 4.0 * x.numel()):
        check(_lib.lib().dd_quick_gelu(_ptr(x), _ptr(out), _L(x.numel()), _stream()), "dd_quick_gelu")
    return out


def nchw_to_rows(src, out=None):
    """[n, C, H, W] fp32/bf16 contiguous -> [n*H*W, C] bf16"""
    assert src.is_cuda and src.is_contiguous()
    n, c, h, w = src.shape
    if out is None:
        out = torch.empty((n * h * w, c), device=src.device, dtype=torch.bfloat16)
    check(_lib.lib().dd_nchw_to_rows(_ptr(src), 1 if src.dtype == torch.float32 else 0, _ptr(out), n, c, h * w,
                                     _stream()), "dd_nchw_to_rows")
    return out


def rows_to_nchw(rows, n_img, hw, out_dtype=torch.float32, c=None):
    H, W = hw
    c = rows.shape[1] if c is None else c
    out = torch.empty((n_img, c, H, W), device=rows.device, dtype=out_dtype)
    check(_lib.lib().dd_rows_to_nchw(_ptr(rows), 1 if rows.dtype == torch.float32 else 0, _L(rows.stride(0)),
                                     _ptr(out), 1 if out_dtype == torch.float32 else 0, n_img, c, H * W, _stream()),
          "dd_rows_to_nchw")
    return out


def cfg_sched_step(eps, x, last, m0, m1, coef, *, n_img, c, hw, cfg=True, eps_nchw=False):
    for t_ in (eps, x, last, m0, m1, coef):
        _req(t_, torch.float32, "cfg_sched_step operand")
    check(_lib.lib().dd_cfg_sched_step(_ptr(eps), _ptr(x), _ptr(last), _ptr(m0), _ptr(m1), _ptr(coef), n_img, c,
                                       hw, 1 if cfg else 0, 1 if eps_nchw else 0, _stream()), "dd_cfg_sched_step")
