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


def gemm(a, w, *, out=None, bias=None, rowvec=None, rows_per_img=1, res1=None, res2=None, a2=None,
         taps=1, conv_hw=None, n_img=None, geglu=False, out_f32=False, force_bn=0):
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
    if bias is not None:
        _req(bias, torch.float32, "bias")
    if rowvec is not None:
        _req(rowvec, torch.float32, "rowvec")
    check(_lib.lib().dd_gemm(C.byref(args), _stream()), "dd_gemm")
    return out
