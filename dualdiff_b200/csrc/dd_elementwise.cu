// dualdiff_b200 — bandwidth-bound layout / elementwise / small-MLP kernels (sm_100a).
// Coalesced 16-byte channel vectors, grid-stride loops sized from the SM count.
#include "dd_api_internal.h"
#include "dd_common.cuh"

namespace dd {

static inline int grid_for(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------------------------------------------
// NCHW -> padded channels-last bf16 (channel zero-pad to cp, halo zero)
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_padded_kernel(const T* __restrict__ src, bf16* __restrict__ out, long long so,
                                      long long sv, long long sc, long long sh, int n_outer, int n_view,
                                      int C, int H, int W, int Cp) {
  const int Wp = W + 1;
  const long long rows = (long long)n_outer * n_view * (H + 1) * Wp;
  const int vec = Cp >> 3;
  const long long total = rows * vec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vec;
    const int v = (int)(i - r * vec);
    const int rows_img = (H + 1) * Wp;
    const int img = (int)(r / rows_img);
    const int rem = (int)(r - (long long)img * rows_img);
    const int y = rem / Wp, x = rem - y * Wp;
    uint32_t pk[4] = {0, 0, 0, 0};
    if (y < H && x < W) {
      const int outer = img / n_view, view = img - outer * n_view;
      const T* p = src + outer * so + view * sv + (long long)y * sh + x;
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = v * 8 + e;
        f[e] = (c < C) ? (float)p[c * sc] : 0.f;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) pk[e] = pack_bf16(f[2 * e], f[2 * e + 1]);
    }
    *reinterpret_cast<uint4*>(out + r * Cp + v * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

int nchw_to_padded_run(const dd_to_padded_args* a, cudaStream_t stream) {
  DD_CHECK(a && a->cp % 8 == 0 && a->cp >= a->c, -1, "dd_nchw_to_padded: cp must be a multiple of 8 and >= c");
  const long long total = (long long)a->n_outer * a->n_view * (a->h + 1) * (a->w + 1) * (a->cp >> 3);
  const int threads = 256;
  const int grid = grid_for(total, threads);
  if (a->src_f32)
    nchw_to_padded_kernel<float><<<grid, threads, 0, stream>>>(
        reinterpret_cast<const float*>(a->src), reinterpret_cast<bf16*>(a->out), a->stride_outer, a->stride_view,
        a->stride_c, a->stride_h, a->n_outer, a->n_view, a->c, a->h, a->w, a->cp);
  else
    nchw_to_padded_kernel<bf16><<<grid, threads, 0, stream>>>(
        reinterpret_cast<const bf16*>(a->src), reinterpret_cast<bf16*>(a->out), a->stride_outer, a->stride_view,
        a->stride_c, a->stride_h, a->n_outer, a->n_view, a->c, a->h, a->w, a->cp);
  DD_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// 3x3 stride-1 pad-1 patch rows of a few-channel NCHW image (conv_in on the 4-channel latents):
// out[(img, y, x), tap*C + c] = src[img, c, y+kh-1, x+kw-1] (zero outside the image), columns [9C, cp) zero.
// One thread per (row, 8-column group).  With C = 4 a row is 36 + 4 columns: the convolution becomes ONE plain GEMM with
// K = 40 on the TMA epilogue instead of nine 8-channel taps through the conv path, whose register epilogue is fully exposed
// when the K loop is this short (conv_in: 177 us measured against a ~40 us HBM floor).
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_patches_kernel(const T* __restrict__ src, bf16* __restrict__ out, long long so, long long sv,
                                    long long sc, long long sh, int n_outer, int n_view, int C, int H, int W, int Cp) {
  const long long rows = (long long)n_outer * n_view * H * W;
  const int vec = Cp >> 3;
  const long long total = rows * vec;
  const int K9 = 9 * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vec;
    const int v = (int)(i - r * vec);
    const int img = (int)(r / (H * W));
    const int rem = (int)(r - (long long)img * (H * W));
    const int y = rem / W, x = rem - y * W;
    const int outer = img / n_view, view = img - outer * n_view;
    const T* base = src + outer * so + view * sv;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = v * 8 + e;
      float val = 0.f;
      if (col < K9) {
        const int tap = col / C, c = col - tap * C;
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) val = (float)base[c * sc + (long long)yy * sh + xx];
      }
      f[e] = val;
    }
    *reinterpret_cast<uint4*>(out + r * Cp + v * 8) =
        make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  }
}

int nchw_patches_run(const dd_to_padded_args* a, cudaStream_t stream) {
  DD_CHECK(a && a->c > 0 && a->cp % 8 == 0 && a->cp >= 9 * a->c, -1,
           "dd_nchw_patches: cp must be a multiple of 8 and >= 9*c");
  DD_CHECK((long long)a->h * a->w < (1LL << 31), -1, "dd_nchw_patches: image too large");
  const long long total = (long long)a->n_outer * a->n_view * a->h * a->w * (a->cp >> 3);
  const int grid = grid_for(total, 256);
  if (a->src_f32)
    nchw_patches_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(a->src), reinterpret_cast<bf16*>(a->out),
                                                         a->stride_outer, a->stride_view, a->stride_c, a->stride_h,
                                                         a->n_outer, a->n_view, a->c, a->h, a->w, a->cp);
  else
    nchw_patches_kernel<bf16><<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(a->src), reinterpret_cast<bf16*>(a->out),
                                                        a->stride_outer, a->stride_view, a->stride_c, a->stride_h,
                                                        a->n_outer, a->n_view, a->c, a->h, a->w, a->cp);
  DD_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// 3x3 stride-2 pad-1 im2col: out[(img, ho, wo), tap*C + c] = x[(img, 2ho+kh-1, 2wo+kw-1), c]
// ---------------------------------------------------------------------------------------------------
__global__ void im2col_s2_kernel(const bf16* __restrict__ x, long long ld, bf16* __restrict__ out, int n_img,
                                 int H, int W, int C, int Ho, int Wo) {
  const int vec = C >> 3;
  const long long total = (long long)n_img * Ho * Wo * 9 * vec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec);
    long long t = i / vec;
    const int tap = (int)(t % 9);
    t /= 9;
    const int wo = (int)(t % Wo);
    t /= Wo;
    const int ho = (int)(t % Ho);
    const int img = (int)(t / Ho);
    const int y = 2 * ho + tap / 3 - 1, xx = 2 * wo + tap % 3 - 1;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < H && xx >= 0 && xx < W)
      val = *reinterpret_cast<const uint4*>(x + (((long long)img * H + y) * W + xx) * ld + v * 8);
    *reinterpret_cast<uint4*>(out + (((long long)img * Ho + ho) * Wo + wo) * (9LL * C) + tap * C + v * 8) = val;
  }
}

// nearest resize -> padded layout
__global__ void upsample_pad_kernel(const bf16* __restrict__ x, long long ld, bf16* __restrict__ out, int n_img,
                                    int H, int W, int C, int H2, int W2) {
  const int vec = C >> 3;
  const int Wp = W2 + 1;
  const int rows_img = (H2 + 1) * Wp;
  const long long total = (long long)n_img * rows_img * vec;
  const float sy = (float)H / (float)H2, sx = (float)W / (float)W2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec);
    const long long r = i / vec;
    const int img = (int)(r / rows_img);
    const int rem = (int)(r - (long long)img * rows_img);
    const int y = rem / Wp, xx = rem - y * Wp;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (y < H2 && xx < W2) {
      // PyTorch 'nearest': src = min(floor(dst * scale), in - 1), scale = in / out in fp32
      const int ys = min((int)floorf(y * sy), H - 1), xs = min((int)floorf(xx * sx), W - 1);
      val = *reinterpret_cast<const uint4*>(x + (((long long)img * H + ys) * W + xs) * ld + v * 8);
    }
    *reinterpret_cast<uint4*>(out + r * C + v * 8) = val;
  }
}

int im2col_s2_run(const void* x, long long ld, void* out, int n_img, int h, int w, int c, cudaStream_t stream) {
  DD_CHECK(c % 8 == 0, -1, "dd_im2col_s2: C must be a multiple of 8");
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const long long total = (long long)n_img * ho * wo * 9 * (c >> 3);
  im2col_s2_kernel<<<grid_for(total, 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), ld,
                                                             reinterpret_cast<bf16*>(out), n_img, h, w, c, ho, wo);
  DD_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int upsample_pad_run(const void* x, long long ld, void* out, int n_img, int h, int w, int c, int h2, int w2,
                     cudaStream_t stream) {
  DD_CHECK(c % 8 == 0, -1, "dd_upsample_pad: C must be a multiple of 8");
  const long long total = (long long)n_img * (h2 + 1) * (w2 + 1) * (c >> 3);
  upsample_pad_kernel<<<grid_for(total, 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), ld,
                                                                reinterpret_cast<bf16*>(out), n_img, h, w, c, h2, w2);
  DD_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// small fp32 linear: one warp per output element (M*N small; K <= a few thousand)
// ---------------------------------------------------------------------------------------------------
__global__ void linear_f32_kernel(const float* __restrict__ x, long long x_ld, const float* __restrict__ w,
                                  const float* __restrict__ b, float* __restrict__ y, long long y_ld,
                                  bf16* __restrict__ y16, long long y16_ld, int M, int N, int K, int act) {
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long o = warp; o < (long long)M * N; o += nwarps) {
    const int m = (int)(o / N), n = (int)(o - (long long)m * N);
    const float* xr = x + m * x_ld;
    const float* wr = w + (long long)n * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(xr[k], wr[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      if (b) acc += b[n];
      if (act == 1) acc = acc / (1.f + expf(-acc));
      if (y) y[m * y_ld + n] = acc;
      if (y16) y16[m * y16_ld + n] = __float2bfloat16(acc);
    }
  }
}

int linear_f32_run(const dd_linear_f32_args* a, cudaStream_t stream) {
  DD_CHECK(a && a->M > 0 && a->N > 0 && a->K > 0, -1, "dd_linear_f32: bad shape");
  const long long warps = (long long)a->M * a->N;
  linear_f32_kernel<<<grid_for(warps * 32, 256), 256, 0, stream>>>(a->x, a->x_ld, a->w, a->b, a->y, a->y_ld,
                                                                   reinterpret_cast<bf16*>(a->y16), a->y16_ld,
                                                                   a->M, a->N, a->K, a->act);
  DD_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

__global__ void timestep_embedding_kernel(const float* __restrict__ t, float* __restrict__ out, int n, int dim) {
  const int half = dim >> 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * half; i += gridDim.x * blockDim.x) {
    const int r = i / half, j = i - r * half;
    // exp(-ln(10000) * j / half) evaluated in fp32 like torch.exp on an fp32 tensor
    const float f = expf(-9.210340371976184f * (float)j / (float)half);
    const float e = t[r] * f;
    out[r * dim + j] = cosf(e);          // flip_sin_to_cos=True: cos first
    out[r * dim + half + j] = sinf(e);
  }
}

__global__ void fourier_embed_kernel(const float* __restrict__ x, float* __restrict__ out, long long rows, int nfreq) {
  const int od = 3 + 6 * nfreq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows * 3;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / 3;
    const int d = (int)(i - r * 3);
    const float v = x[i];
    float* o = out + r * od;
    o[d] = v;
    float f = 1.f;
    for (int k = 0; k < nfreq; ++k) {
      o[3 + 6 * k + d] = sinf(v * f);
      o[3 + 6 * k + 3 + d] = cosf(v * f);
      f *= 2.f;
    }
  }
}

// box features: one block per box token
__global__ void box_features_kernel(const float* __restrict__ boxes, const long long* __restrict__ classes,
                                    const unsigned char* __restrict__ masks, const float* __restrict__ class_tokens,
                                    const float* __restrict__ null_pos, const float* __restrict__ null_cls,
                                    float* __restrict__ pos_out, long long pos_ld, float* __restrict__ cls_out,
                                    long long cls_ld, int n_pts, int cls_dim, int n_classes) {
  const long long b = blockIdx.x;
  const float m = masks[b] ? 1.f : 0.f;
  const int nfreq = 4, od = 27;
  const float* bx = boxes + b * n_pts * 3;
  for (int i = threadIdx.x; i < n_pts * 3; i += blockDim.x) {
    const int pt = i / 3, d = i - pt * 3;
    const float v = bx[i];
    float* o = pos_out + b * pos_ld + pt * od;
    const float* np = null_pos + pt * od;
    o[d] = v * m + np[d] * (1.f - m);
    float f = 1.f;
    for (int k = 0; k < nfreq; ++k) {
      o[3 + 6 * k + d] = sinf(v * f) * m + np[3 + 6 * k + d] * (1.f - m);
      o[3 + 6 * k + 3 + d] = cosf(v * f) * m + np[3 + 6 * k + 3 + d] * (1.f - m);
      f *= 2.f;
    }
  }
  // The reference's collate pads `classes` with -1 (dataset/utils.py:243,283) and indexes class_tokens[-1], a valid
  // Python wrap-around whose value is then multiplied by the mask 0 (bbox_embedder.py:189-190).  A masked-out slot
  // therefore never reads the token table here; an unmasked id follows the Python rule (negative ids count from the
  // end) and is clamped into the table -- the host wrapper rejects ids outside [-n_classes, n_classes) beforehand.
  if (m == 0.f) {
    for (int i = threadIdx.x; i < cls_dim; i += blockDim.x) cls_out[b * cls_ld + i] = null_cls[i];
    return;
  }
  long long c = classes[b];
  if (c < 0) c += n_classes;
  c = c < 0 ? 0 : (c >= n_classes ? n_classes - 1 : c);
  const float* ct = class_tokens + c * cls_dim;
  for (int i = threadIdx.x; i < cls_dim; i += blockDim.x) cls_out[b * cls_ld + i] = ct[i];
}

__global__ void silu_to_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    out[i] = __float2bfloat16(v / (1.f + expf(-v)));
  }
}

__global__ void add_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, const uint4* __restrict__ c,
                                uint4* __restrict__ out, long long nvec) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const uint4 va = a[i], vb = b[i];
    uint4 vc = make_uint4(0, 0, 0, 0);
    if (c) vc = c[i];
    const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w}, wc[4] = {vc.x, vc.y, vc.z, vc.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = unpack_bf16(wa[e]), fb = unpack_bf16(wb[e]), fc = unpack_bf16(wc[e]);
      o[e] = pack_bf16(fa.x + fb.x + fc.x, fa.y + fb.y + fc.y);
    }
    out[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// NCHW <-> channels-last rows (module-boundary conversions; small tensors or one-off)
template <typename TI>
__global__ void nchw_to_rows_kernel(const TI* __restrict__ src, bf16* __restrict__ out, int n_img, int C, int HW) {
  const long long total = (long long)n_img * C * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long t = i / C;
    const int p = (int)(t % HW);
    const int img = (int)(t / HW);
    out[i] = __float2bfloat16((float)src[((long long)img * C + c) * HW + p]);
  }
}
template <typename TI, typename TO>
__global__ void rows_to_nchw_kernel(const TI* __restrict__ rows, long long ld, TO* __restrict__ out, int n_img, int C,
                                    int HW) {
  const long long total = (long long)n_img * C * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const long long t = i / HW;
    const int c = (int)(t % C);
    const int img = (int)(t / C);
    out[i] = (TO)((float)rows[((long long)img * HW + p) * ld + c]);
  }
}

// ---------------------------------------------------------------------------------------------------
// CFG combine + scheduler update (UniPC bh2 / DDIM) — all coefficients precomputed on the host
// ---------------------------------------------------------------------------------------------------
__global__ void cfg_sched_kernel(const float* __restrict__ eps, float* __restrict__ x, float* __restrict__ last,
                                 float* __restrict__ m0, float* __restrict__ m1, const float* __restrict__ coef,
                                 int n_img, int C, int HW, int cfg, int eps_nchw) {
  const float g = coef[0], sigma = coef[1], inv_alpha = coef[2];
  const float a_last = coef[3], a_m0 = coef[4], a_m1 = coef[5], a_x0 = coef[6], a_x = coef[7];
  const float b_xc = coef[8], b_x0 = coef[9], b_m0 = coef[10];
  const long long total = (long long)n_img * C * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const long long t = i / HW;
    const int c = (int)(t % C);
    const int img = (int)(t / C);
    const long long ei = eps_nchw ? i : ((long long)img * HW + p) * C + c;
    float e = eps[ei];
    if (cfg) {
      const float ec = eps[ei + (long long)n_img * HW * C];
      e = e + g * (ec - e);
    }
    const float xv = x[i], lv = last[i], m0v = m0[i], m1v = m1[i];
    const float x0 = (xv - sigma * e) * inv_alpha;
    const float xc = a_x * xv + a_last * lv + a_m0 * m0v + a_m1 * m1v + a_x0 * x0;
    const float xn = b_xc * xc + b_x0 * x0 + b_m0 * m0v;
    last[i] = xc;
    m1[i] = m0v;
    m0[i] = x0;
    x[i] = xn;
  }
}

// softmax over the rows of an fp32 score matrix -> bf16 probabilities: one warp per row (row read once into registers
// for up to 2048 columns, otherwise re-read).  Used by the single-head, 512-wide mid-block attention of the VAE decoder,
// whose head dimension does not fit the flash kernel's TMEM budget (S 128 + O 512 + P 64 columns > 512).
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ x, long long x_ld, bf16* __restrict__ out, long long out_ld, int rows, int cols,
                    float scale_log2e) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (long long)row * x_ld;
  float m = -INFINITY;
  for (int c = lane; c < cols; c += 32) m = fmaxf(m, xr[c]);
  m = warp_max(m);
  float l = 0.f;
  for (int c = lane; c < cols; c += 32) l += exp2f((xr[c] - m) * scale_log2e);
  l = warp_sum(l);
  const float inv = 1.f / l;
  bf16* orow = out + (long long)row * out_ld;
  for (int c = lane; c < cols; c += 32) orow[c] = __float2bfloat16(exp2f((xr[c] - m) * scale_log2e) * inv);
}
}  // namespace dd

using namespace dd;
extern "C" {
int dd_nchw_to_padded(const dd_to_padded_args* args, void* stream) {
  return nchw_to_padded_run(args, reinterpret_cast<cudaStream_t>(stream));
}
int dd_nchw_patches(const dd_to_padded_args* args, void* stream) {
  return nchw_patches_run(args, reinterpret_cast<cudaStream_t>(stream));
}
int dd_im2col_s2(const void* x, long long x_ld, void* out, int n_img, int h, int w, int c, void* stream) {
  return im2col_s2_run(x, x_ld, out, n_img, h, w, c, reinterpret_cast<cudaStream_t>(stream));
}
int dd_upsample_pad(const void* x, long long x_ld, void* out, int n_img, int h, int w, int c, int h2, int w2,
                    void* stream) {
  return upsample_pad_run(x, x_ld, out, n_img, h, w, c, h2, w2, reinterpret_cast<cudaStream_t>(stream));
}
int dd_pad_rows(const void* x, long long x_ld, void* out, int n_img, int h, int w, int c, void* stream) {
  // identity resize into the padded layout
  return upsample_pad_run(x, x_ld, out, n_img, h, w, c, h, w, reinterpret_cast<cudaStream_t>(stream));
}
int dd_linear_f32(const dd_linear_f32_args* args, void* stream) {
  return linear_f32_run(args, reinterpret_cast<cudaStream_t>(stream));
}
int dd_timestep_embedding(const float* t, float* out, int n, int dim, void* stream) {
  if (n <= 0 || dim <= 0 || (dim & 1)) { set_error("dd_timestep_embedding: bad shape"); return -1; }
  timestep_embedding_kernel<<<grid_for((long long)n * dim / 2, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(t, out, n, dim);
  if (cudaGetLastError() != cudaSuccess) { set_error("dd_timestep_embedding launch failed"); return -2; }
  count_launch();
  return 0;
}
int dd_fourier_embed(const float* x, float* out, long long rows, int nfreq, void* stream) {
  if (rows <= 0) { set_error("dd_fourier_embed: bad shape"); return -1; }
  fourier_embed_kernel<<<grid_for(rows * 3, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, out, rows, nfreq);
  if (cudaGetLastError() != cudaSuccess) { set_error("dd_fourier_embed launch failed"); return -2; }
  count_launch();
  return 0;
}
int dd_box_features(const float* boxes, const long long* classes, const unsigned char* masks,
                    const float* class_tokens, const float* null_pos, const float* null_cls, float* pos_out,
                    long long pos_ld, float* cls_out, long long cls_ld, long long n_box, int n_pts, int cls_dim,
                    int n_classes, void* stream) {
  if (n_box <= 0 || n_classes <= 0) { set_error("dd_box_features: bad shape"); return -1; }
  box_features_kernel<<<(unsigned)n_box, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      boxes, classes, masks, class_tokens, null_pos, null_cls, pos_out, pos_ld, cls_out, cls_ld, n_pts, cls_dim, n_classes);
  if (cudaGetLastError() != cudaSuccess) { set_error("dd_box_features launch failed"); return -2; }
  count_launch();
  return 0;
}
int dd_silu_to_bf16(const float* x, void* out, long long n, void* stream) {
  silu_to_bf16_kernel<<<grid_for(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, reinterpret_cast<bf16*>(out), n);
  if (cudaGetLastError() != cudaSuccess) { set_error("dd_silu_to_bf16 launch failed"); return -2; }
  count_launch();
  return 0;
}
int dd_softmax_rows(const float* x, long long x_ld, void* out, long long out_ld, int rows, int cols, float scale,
                    void* stream) {
  if (rows <= 0 || cols <= 0) { set_error("dd_softmax_rows: bad shape"); return -1; }
  softmax_rows_kernel<<<(unsigned)(((long long)rows * 32 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, x_ld, reinterpret_cast<bf16*>(out), out_ld, rows, cols, scale * 1.4426950408889634f);
  if (cudaGetLastError() != cudaSuccess) { set_error("dd_softmax_rows launch failed"); return -2; }
  count_launch();
  return 0;
}
int dd_add_bf16(const void* a, const void* b, const void* c, void* out, long long n, void* stream) {
  if (n % 8 != 0) { set_error("dd_add_bf16: n must be a multiple of 8"); return -1; }
  add_bf16_kernel<<<grid_for(n / 8, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b), reinterpret_cast<const uint4*>(c),
      reinterpret_cast<uint4*>(out), n / 8);
  if (cudaGetLastError() != cudaSuccess) { set_error("dd_add_bf16 launch failed"); return -2; }
  count_launch();
  return 0;
}
int dd_nchw_to_rows(const void* src, int src_f32, void* out, int n_img, int c, int hw, void* stream) {
  const long long total = (long long)n_img * c * hw;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (src_f32)
    nchw_to_rows_kernel<float><<<grid_for(total, 256), 256, 0, s>>>(reinterpret_cast<const float*>(src), reinterpret_cast<bf16*>(out), n_img, c, hw);
  else
    nchw_to_rows_kernel<bf16><<<grid_for(total, 256), 256, 0, s>>>(reinterpret_cast<const bf16*>(src), reinterpret_cast<bf16*>(out), n_img, c, hw);
  if (cudaGetLastError() != cudaSuccess) { set_error("dd_nchw_to_rows launch failed"); return -2; }
  count_launch();
  return 0;
}
int dd_rows_to_nchw(const void* rows, int rows_f32, long long ld, void* out, int out_f32, int n_img, int c, int hw,
                    void* stream) {
  const long long total = (long long)n_img * c * hw;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int g = grid_for(total, 256);
  if (rows_f32 && out_f32)
    rows_to_nchw_kernel<float, float><<<g, 256, 0, s>>>(reinterpret_cast<const float*>(rows), ld, reinterpret_cast<float*>(out), n_img, c, hw);
  else if (!rows_f32 && out_f32)
    rows_to_nchw_kernel<bf16, float><<<g, 256, 0, s>>>(reinterpret_cast<const bf16*>(rows), ld, reinterpret_cast<float*>(out), n_img, c, hw);
  else if (!rows_f32 && !out_f32)
    rows_to_nchw_kernel<bf16, bf16><<<g, 256, 0, s>>>(reinterpret_cast<const bf16*>(rows), ld, reinterpret_cast<bf16*>(out), n_img, c, hw);
  else { set_error("dd_rows_to_nchw: fp32 -> bf16 not supported"); return -1; }
  if (cudaGetLastError() != cudaSuccess) { set_error("dd_rows_to_nchw launch failed"); return -2; }
  count_launch();
  return 0;
}
int dd_cfg_sched_step(const float* eps, float* x, float* last, float* m0, float* m1, const float* coef, int n_img,
                      int c, int hw, int cfg, int eps_nchw, void* stream) {
  if (n_img <= 0 || c <= 0 || hw <= 0) { set_error("dd_cfg_sched_step: bad shape"); return -1; }
  const long long total = (long long)n_img * c * hw;
  cfg_sched_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(eps, x, last, m0, m1, coef, n_img, c, hw, cfg, eps_nchw);
  if (cudaGetLastError() != cudaSuccess) { set_error("dd_cfg_sched_step launch failed"); return -2; }
  count_launch();
  return 0;
}
}
