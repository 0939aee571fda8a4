"""Weight / activation layout packers (done once at load time, diffusers state-dict -> kernel layouts)."""
import torch


def pack_conv3x3(w):
    """OIHW [Cout, Cin, 3, 3] -> [Cout, 9*Cin] bf16, tap-major (kh, kw) then channel (K of the implicit GEMM)."""
    co, ci, kh, kw = w.shape
    assert kh == 3 and kw == 3
    return w.permute(0, 2, 3, 1).reshape(co, 9 * ci).contiguous().to(torch.bfloat16)


def pack_conv3x3_patch(w):
    """OIHW [Cout, Cin, 3, 3] -> [Cout, cp] bf16 for the explicit patch-matrix form (dd_nchw_patches): columns tap-major,
    channel-minor like pack_conv3x3, zero-padded to the next multiple of 8 (Cin = 4: 36 -> 40)."""
    co, ci, kh, kw = w.shape
    assert kh == 3 and kw == 3
    k = 9 * ci
    cp = (k + 7) // 8 * 8
    out = w.new_zeros((co, cp))
    out[:, :k] = w.permute(0, 2, 3, 1).reshape(co, k)
    return out.contiguous().to(torch.bfloat16)


def pack_conv1x1(w):
    co, ci = w.shape[:2]
    return w.reshape(co, ci).contiguous().to(torch.bfloat16)


def pack_linear(w):
    return w.contiguous().to(torch.bfloat16)


GEGLU_HALF = 128  # value/gate columns per 256-wide GEMM tile


def pack_geglu(w, b):
    """diffusers GEGLU: proj = Linear(C, 2*inner); value, gate = proj(x).chunk(2).  Interleave the rows in
    groups of 128 so one 256-column tile of the GEMM holds 128 value columns and their 128 gate columns."""
    two_inner, c = w.shape
    inner = two_inner // 2
    assert inner % GEGLU_HALF == 0, inner
    wv, wg = w[:inner], w[inner:]
    bv, bg = b[:inner], b[inner:]
    g = inner // GEGLU_HALF
    wp = torch.stack([wv.reshape(g, GEGLU_HALF, c), wg.reshape(g, GEGLU_HALF, c)], dim=1).reshape(two_inner, c)
    bp = torch.stack([bv.reshape(g, GEGLU_HALF), bg.reshape(g, GEGLU_HALF)], dim=1).reshape(two_inner)
    return wp.contiguous().to(torch.bfloat16), bp.contiguous().float()


def to_padded(x_nhwc):
    """[n, H, W, C] -> zero-haloed [n*(H+1)*(W+1), C] (test helper; the GN kernel writes this layout directly)."""
    n, H, W, C = x_nhwc.shape
    out = x_nhwc.new_zeros((n, H + 1, W + 1, C))
    out[:, :H, :W] = x_nhwc
    return out.reshape(n * (H + 1) * (W + 1), C)
