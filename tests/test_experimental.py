"""EXPERIMENTAL variants that are not on the default path yet (built at the end of round 1): off by
default, their GPU checks only run with DD_EXPERIMENTAL=1 so that a variant still under evaluation can never turn the default
suite red.  The host-side halves are checked here on the CPU.  Status: the conv_in patch-matrix checks below passed on a
B200 (3 of 3) at the end of round 1; the variant has not been timed or run inside the step graph yet.  The stride-1 patch
matrix for small feature maps (second half of this file) has not run on a GPU at all."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
EXPERIMENTAL = bool(int(os.environ.get("DD_EXPERIMENTAL", "0")))


def _patch_rows(x, cp):
    """torch restatement of dd_nchw_patches: [n, C, H, W] -> [n*H*W, cp], columns tap-major / channel-minor, zero tail"""
    n, c, h, w = x.shape
    cols = torch.nn.functional.unfold(x, 3, padding=1)                       # [n, C*9, H*W], channel-major / tap-minor
    cols = cols.reshape(n, c, 9, h * w).permute(0, 3, 2, 1).reshape(n * h * w, 9 * c)
    out = x.new_zeros((n * h * w, cp))
    out[:, :9 * c] = cols
    return out


def test_patch_weight_packing_reproduces_the_convolution():
    from dualdiff_b200.packing import pack_conv3x3_patch
    g = torch.Generator().manual_seed(0)
    w, b = torch.randn(16, 4, 3, 3, generator=g), torch.randn(16, generator=g)
    x = torch.randn(3, 4, 5, 7, generator=g)
    wp = pack_conv3x3_patch(w)
    assert wp.shape == (16, 40) and wp.dtype == torch.bfloat16 and (wp[:, 36:] == 0).all()
    out = _patch_rows(x, 40) @ wp.float().T + b
    ref = torch.nn.functional.conv2d(x, w.to(torch.bfloat16).float(), b, padding=1).permute(0, 2, 3, 1).reshape(-1, 16)
    assert (out - ref).abs().max() < 1e-4


@pytest.mark.gpu
@pytest.mark.skipif(not EXPERIMENTAL, reason="unvalidated variant: run with DD_EXPERIMENTAL=1")
@pytest.mark.parametrize("n_outer,n_view,h,w,shared", [(2, 6, 28, 50, True), (1, 3, 5, 7, False), (2, 2, 8, 12, False)])
def test_nchw_patches_kernel_and_conv_in_gemm(n_outer, n_view, h, w, shared):
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
    # the shipped path: nine 8-channel taps through the implicit-GEMM convolution
    w8 = torch.zeros(320, 8, 3, 3); w8[:, :4] = wt
    pad = ops.nchw_to_padded(lat.cuda(), cp=8, **kw)
    old = ops.gemm(pad, pack_conv3x3(w8).cuda(), bias=b.cuda(), taps=9, conv_hw=(h, w), n_img=n_outer * n_view,
                   res1=res.cuda()).float().cpu()
    ref = torch.nn.functional.conv2d(full.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float(), b, padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, 320) + res.float()
    assert (out - ref).abs().max() <= 2e-2 * ref.abs().max()
    assert (out - old).abs().max() <= 2e-2 * ref.abs().max()


# ---- explicit stride-1 patch matrix for the smallest feature maps (DD_SMALL_CONV_IM2COL) ---------------------------------
def _im2col_s1_rows(x_rows, n, h, w):
    """torch restatement of dd_im2col_s1: compact rows [n*h*w, C] -> [n*h*w, 9*C], tap-major / channel-minor"""
    c = x_rows.shape[1]
    x = x_rows.reshape(n, h, w, c).permute(0, 3, 1, 2)
    cols = torch.nn.functional.unfold(x, 3, padding=1).reshape(n, c, 9, h * w)
    return cols.permute(0, 3, 2, 1).reshape(n * h * w, 9 * c)


def test_im2col_s1_with_the_tap_major_conv_weights_reproduces_the_convolution():
    from dualdiff_b200.packing import pack_conv3x3
    g = torch.Generator().manual_seed(2)
    w, x = torch.randn(24, 16, 3, 3, generator=g), torch.randn(3, 16, 4, 7, generator=g)
    rows = x.permute(0, 2, 3, 1).reshape(-1, 16)
    out = _im2col_s1_rows(rows, 3, 4, 7) @ pack_conv3x3(w).float().T
    ref = torch.nn.functional.conv2d(x, w.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 1).reshape(-1, 24)
    assert (out - ref).abs().max() < 1e-3


@pytest.mark.gpu
@pytest.mark.skipif(not EXPERIMENTAL, reason="unvalidated variant: run with DD_EXPERIMENTAL=1")
@pytest.mark.parametrize("n,h,w,cin,cout", [(96, 4, 7, 1280, 1280), (5, 4, 7, 64, 320), (3, 1, 1, 128, 64), (2, 7, 13, 320, 160)])
def test_im2col_s1_kernel_and_gemm_match_the_implicit_conv(n, h, w, cin, cout):
    from dualdiff_b200 import ops
    from dualdiff_b200.packing import pack_conv3x3, to_padded
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n * h * w, cin, generator=g).to(torch.bfloat16)
    wt, b = torch.randn(cout, cin, 3, 3, generator=g) * (cin * 9) ** -0.5, torch.randn(cout, generator=g)
    rv = torch.randn(n, cout, generator=g)
    res = torch.randn(n * h * w, cout, generator=g).to(torch.bfloat16)
    cols = ops.im2col_s1(x.cuda(), n_img=n, hw=(h, w))
    assert torch.equal(cols.cpu(), _im2col_s1_rows(x, n, h, w))
    W = pack_conv3x3(wt).cuda()
    new = ops.gemm(cols, W, bias=b.cuda(), rowvec=rv.cuda(), rows_per_img=h * w, res1=res.cuda()).float().cpu()
    pad = to_padded(x.reshape(n, h, w, cin)).cuda()
    old = ops.gemm(pad, W, bias=b.cuda(), taps=9, conv_hw=(h, w), n_img=n, rowvec=rv.cuda(), rows_per_img=h * w,
                   res1=res.cuda()).float().cpu()
    xi = x.float().reshape(n, h, w, cin).permute(0, 3, 1, 2)
    ref = torch.nn.functional.conv2d(xi, wt.to(torch.bfloat16).float(), b, padding=1) + rv[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1).reshape(-1, cout) + res.float()
    assert (new - ref).abs().max() <= 2e-2 * ref.abs().max()
    assert (new - old).abs().max() <= 2e-2 * ref.abs().max()
