// dualdiff_b200 — bandwidth-bound normalisation kernels (sm_100a).
//   * GroupNorm(32) [+ SiLU] over channels-last activations, optional channel-concat of two sources
//     (up-block skip connections), output either compact rows or the zero-haloed padded-pixel layout the
//     3x3 implicit-GEMM conv consumes.  Replaces diffusers ResnetBlock2D.norm1/norm2 + SiLU,
//     Transformer2DModel.norm and conv_norm_out (unet_2d_condition_multiview.py:519-521).
//   * LayerNorm over the channel dim of token rows (blocks.py:163,177,192,225).
// fp32 statistics, bf16 I/O, 16-byte vector accesses, warp-shuffle reductions.
#include <stdlib.h>

#include "dd_api_internal.h"
#include "dd_common.cuh"

namespace dd {

// ---------------------------------------------------------------------------------------------------
// GroupNorm pass 1: per (image, group) sum / sum of squares.  A CTA owns a slab of rows of one image; threads that
// share a channel octet reduce through shared memory, channels fold into their groups, and the CTA stores ONE
// partial row: partial[img][cta][groups][2] (fp32; plain stores, fixed order -> bit-reproducible, no atomics).
// Eight independent 16-byte loads per thread are in flight (2 CTAs x 512 threads x 128 B = 128 KB per SM; HBM
// latency x bandwidth needs ~40 KB per SM).  Images and slabs are walked in REVERSE launch order: the tail of the
// activation is what the producing GEMM wrote last and is still in L2, and the head this pass touches last is what
// the apply pass (forward order) reads first.
// ---------------------------------------------------------------------------------------------------
constexpr int GN_U = 8;

__global__ void __launch_bounds__(512, 2)
gn_stats_kernel(const bf16* __restrict__ x1, long long ld1, int C1, const bf16* __restrict__ x2,
                long long ld2, int C, int HW, int groups, int rows_per_cta, float* __restrict__ stats) {
  extern __shared__ float red[];  // [rpi][C][2], then [C][2] at red2
  const int img = gridDim.y - 1 - blockIdx.y;
  const int slab = gridDim.x - 1 - blockIdx.x;
  const int tpr = C >> 3;                    // threads per row (8 channels each)
  const int rpi = blockDim.x / tpr;          // rows per iteration
  const int lane_c = threadIdx.x % tpr;
  const int sub = threadIdx.x / tpr;
  const int c0 = lane_c * 8;
  const int r_begin = slab * rows_per_cta;
  const int r_end = min(HW, r_begin + rows_per_cta);
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  if (sub < rpi) {
    const bool from1 = c0 < C1;
    const long long ld = from1 ? ld1 : ld2;
    const bf16* base = (from1 ? x1 + c0 : x2 + (c0 - C1)) + (long long)img * HW * ld;
    for (int r = r_begin + sub; r < r_end; r += GN_U * rpi) {
      uint4 v[GN_U];
#pragma unroll
      for (int u = 0; u < GN_U; ++u) {
        const int rr = r + u * rpi;
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        if (rr < r_end) v[u] = *reinterpret_cast<const uint4*>(base + (long long)rr * ld);
      }
#pragma unroll
      for (int u = 0; u < GN_U; ++u) {
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16(w[e]);
          s[2 * e] += f.x; q[2 * e] += f.x * f.x;
          s[2 * e + 1] += f.y; q[2 * e + 1] += f.y * f.y;
        }
      }
    }
    float* o = red + ((size_t)sub * C + c0) * 2;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      o[2 * e] = s[e];
      o[2 * e + 1] = q[e];
    }
  }
  __syncthreads();
  // fold the rpi row-slices (fixed order), per channel
  float* red2 = red + (size_t)rpi * 2 * C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < rpi; ++k) acc += red[(size_t)k * 2 * C + i];
    red2[i] = acc;
  }
  __syncthreads();
  // channels -> groups: one warp per group
  const int cpg = C / groups;
  const int lane = threadIdx.x & 31;
  float* part = stats + ((long long)img * gridDim.x + slab) * 2 * groups;
  for (int g = threadIdx.x >> 5; g < groups; g += blockDim.x >> 5) {
    float gs = 0.f, gq = 0.f;
    for (int c = lane; c < cpg; c += 32) {
      const float2 v = *reinterpret_cast<const float2*>(red2 + 2 * (g * cpg + c));
      gs += v.x;
      gq += v.y;
    }
    gs = warp_sum(gs);
    gq = warp_sum(gq);
    if (lane == 0) *reinterpret_cast<float2*>(part + 2 * g) = make_float2(gs, gq);
  }
}

// ---------------------------------------------------------------------------------------------------
// GroupNorm pass 2: normalise + affine (+ SiLU) into the compact or padded layout.  A thread owns 8 channels (one or
// two groups for the SDv1.5 widths); it sums the n_part per-CTA partials of those groups itself (a handful of
// independent L2 loads, one round trip -- there is no finalize launch), then streams its rows with four loads in flight.
// ---------------------------------------------------------------------------------------------------
constexpr int GN_UA = 4;

__global__ void __launch_bounds__(512, 2)
gn_apply_kernel(const bf16* __restrict__ x1, long long ld1, int C1, const bf16* __restrict__ x2,
                long long ld2, int C, int H, int W, const float* __restrict__ partial, int n_part,
                const float* __restrict__ gamma, const float* __restrict__ beta, int groups, float eps, int silu,
                int padded, bf16* __restrict__ out, long long out_ld, int rows_per_cta) {
  const int img = blockIdx.y;
  const int HW = H * W;
  const int tpr = C >> 3;
  const int rpi = blockDim.x / tpr;
  const int lane_c = threadIdx.x % tpr;
  const int sub = threadIdx.x / tpr;
  if (sub >= rpi) return;
  const int c0 = lane_c * 8;
  float sc[8], sh[8];
  {
    const int cpg = C / groups;
    const float inv_n = 1.f / (float)(cpg * HW);
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + c0), g1 = *reinterpret_cast<const float4*>(gamma + c0 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + c0), b1 = *reinterpret_cast<const float4*>(beta + c0 + 4);
    const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const float* pimg = partial + (long long)img * n_part * groups * 2;
    int g_prev = -1;
    float mean = 0.f, rstd = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int g = (c0 + e) / cpg;
      if (g != g_prev) {
        float gs = 0.f, gq = 0.f;
        for (int k = 0; k < n_part; ++k) {
          const float2 v = *reinterpret_cast<const float2*>(pimg + ((long long)k * groups + g) * 2);
          gs += v.x;
          gq += v.y;
        }
        mean = gs * inv_n;
        rstd = rsqrtf(fmaxf(gq * inv_n - mean * mean, 0.f) + eps);
        g_prev = g;
      }
      sc[e] = ga[e] * rstd;
      sh[e] = be[e] - mean * sc[e];
    }
  }
  const bool from1 = c0 < C1;
  const long long ld = from1 ? ld1 : ld2;
  const bf16* base = (from1 ? x1 + c0 : x2 + (c0 - C1)) + (long long)img * HW * ld;
  const int Wp = W + 1;
  const int rows_img = padded ? (H + 1) * Wp : HW;
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(rows_img, r_begin + rows_per_cta);
  bf16* obase = out + (long long)img * rows_img * out_ld + c0;
  const uint32_t wp_magic = (uint32_t)((0x100000000ull + (uint32_t)Wp - 1) / (uint32_t)Wp);   // rr / Wp, exact for rr * Wp < 2^32
  for (int r = r_begin + sub; r < r_end; r += GN_UA * rpi) {
    uint4 v[GN_UA];
    bool live[GN_UA];
#pragma unroll
    for (int u = 0; u < GN_UA; ++u) {
      const int rr = r + u * rpi;
      int src = rr;
      live[u] = rr < r_end;
      if (padded) {
        const int hp = (int)__umulhi((uint32_t)rr, wp_magic), wp = rr - hp * Wp;
        live[u] = live[u] && (hp < H) && (wp < W);
        src = hp * W + wp;
      }
      v[u] = make_uint4(0u, 0u, 0u, 0u);
      if (live[u]) v[u] = *reinterpret_cast<const uint4*>(base + (long long)src * ld);
    }
#pragma unroll
    for (int u = 0; u < GN_UA; ++u) {
      const int rr = r + u * rpi;
      if (rr >= r_end) break;
      uint4 o = make_uint4(0u, 0u, 0u, 0u);   // halo rows / columns of the padded layout are zeros
      if (live[u]) {
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16(w[e]);
          float a = f.x * sc[2 * e] + sh[2 * e];
          float b = f.y * sc[2 * e + 1] + sh[2 * e + 1];
          if (silu) {
            a = silu_f(a);
            b = silu_f(b);
          }
          pk[e] = pack_bf16(a, b);
        }
        o = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      *reinterpret_cast<uint4*>(obase + (long long)rr * out_ld) = o;
    }
  }
}

// number of stats CTAs (= partial rows) per image: one wave of two 512-thread CTAs per SM, every CTA non-empty
static int groupnorm_partials(int n_img, int HW, int rpi) {
  const int sms = num_sms();
  int per_img = (2 * sms) / n_img;
  if (per_img < 1) per_img = 1;
  int rows_per_cta = (HW + per_img - 1) / per_img;
  const int min_rows = rpi * GN_U;
  if (rows_per_cta < min_rows) rows_per_cta = min_rows;
  return (HW + rows_per_cta - 1) / rows_per_cta;
}

static int groupnorm_threads(int C, int* rpi_out) {
  const int tpr = C >> 3;
  int threads = 512;
  if (tpr > 512) threads = tpr;
  threads = (threads / tpr) * tpr;
  threads = ((threads + 31) / 32) * 32;
  if (threads > 512) threads = 512;
  *rpi_out = threads / tpr;
  return threads;
}

long long groupnorm_scratch_floats(int n_img, int C, int HW) {
  int rpi = 1;
  groupnorm_threads(C, &rpi);
  if (rpi < 1) rpi = 1;
  (void)C;
  return (long long)n_img * 2 * 512 * groupnorm_partials(n_img, HW, rpi);   // [n_img][n_part][groups <= 512][2]
}

int groupnorm_run(const dd_groupnorm_args* a, cudaStream_t stream) {
  DD_CHECK(a != nullptr, -1, "dd_groupnorm: null args");
  const int C = a->c1 + a->c2;
  DD_CHECK(a->n_img > 0 && a->h > 0 && a->w > 0 && C > 0, -1, "dd_groupnorm: bad shape");
  DD_CHECK(C % 8 == 0 && a->c1 % 8 == 0, -1, "dd_groupnorm: channels must be multiples of 8 (C=%d c1=%d)", C, a->c1);
  DD_CHECK(C % a->groups == 0 && a->groups <= 512, -1, "dd_groupnorm: C=%d not divisible by groups=%d (<= 512)", C, a->groups);
  DD_CHECK(C <= 4096, -1, "dd_groupnorm: C=%d too large (max 4096)", C);
  DD_CHECK(a->c2 == 0 || a->x2 != nullptr, -1, "dd_groupnorm: x2 missing");
  // padded-row index / (W+1) by multiply-high is exact while rows * (W+1) < 2^32
  DD_CHECK((long long)(a->h + 1) * (a->w + 1) * (a->w + 1) < (1LL << 32), -1, "dd_groupnorm: image too large (%d x %d)", a->h, a->w);
  const int HW = a->h * a->w;
  int rpi = 0;
  const int threads = groupnorm_threads(C, &rpi);
  DD_CHECK(rpi >= 1, -1, "dd_groupnorm: C=%d too large", C);
  // stats scratch layout: [n_img][n_part][groups][2] per-CTA partial group sums
  const int sms = num_sms();
  const int n_part = groupnorm_partials(a->n_img, HW, rpi);
  float* sums = a->stats;
  {
    // enough CTAs to fill the machine, few enough that the per-CTA atomics stay negligible
    const int rows_per_cta = (HW + n_part - 1) / n_part;
    dim3 grid(n_part, a->n_img);
    const size_t smem = sizeof(float) * 2 * (size_t)C * (rpi + 1);
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(gn_stats_kernel), 96 * 1024)) return rc;
    DD_CHECK(smem <= 96 * 1024, -1, "dd_groupnorm: reduction buffer too large");
    gn_stats_kernel<<<grid, threads, smem, stream>>>(reinterpret_cast<const bf16*>(a->x1), a->x1_ld, a->c1,
                                                     reinterpret_cast<const bf16*>(a->x2), a->x2_ld, C, HW,
                                                     a->groups, rows_per_cta, sums);
    DD_CUDA(cudaGetLastError());
  }
  {
    const int rows_img = a->padded_out ? (a->h + 1) * (a->w + 1) : HW;
    int per_img = (2 * sms) / a->n_img;   // one wave
    if (per_img < 1) per_img = 1;
    int rows_per_cta = (rows_img + per_img - 1) / per_img;
    if (rows_per_cta < rpi * GN_UA) rows_per_cta = rpi * GN_UA;
    dim3 grid((rows_img + rows_per_cta - 1) / rows_per_cta, a->n_img);
    gn_apply_kernel<<<grid, threads, 0, stream>>>(
        reinterpret_cast<const bf16*>(a->x1), a->x1_ld, a->c1, reinterpret_cast<const bf16*>(a->x2),
        a->x2_ld, C, a->h, a->w, sums, n_part, a->gamma, a->beta, a->groups, a->eps, a->silu, a->padded_out,
        reinterpret_cast<bf16*>(a->out), a->out_ld, rows_per_cta);
    DD_CUDA(cudaGetLastError());
  }
  count_launch(2);
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers (C <= 1280 -> <= 5 x uint4 per lane)
// ---------------------------------------------------------------------------------------------------
template <int VPL>  // uint4 vectors per lane
__global__ void __launch_bounds__(256)
layernorm_kernel(const bf16* __restrict__ x, long long x_ld, bf16* __restrict__ out, long long out_ld,
                 const float* __restrict__ gamma, const float* __restrict__ beta, int rows, int C,
                 float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int nvec = C >> 3;
  const bf16* xr = x + (long long)warp * x_ld;
  float v[VPL][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const uint4 u = *reinterpret_cast<const uint4*>(xr + vi * 8);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16(w[e]);
        v[i][2 * e] = f.x;
        v[i][2 * e + 1] = f.y;
        s += f.x + f.y;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[i][e] - mean;
        q += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  bf16* orow = out + (long long)warp * out_ld;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + vi * 8);
      const float4 g1 = *reinterpret_cast<const float4*>(gamma + vi * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(beta + vi * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(beta + vi * 8 + 4);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        pk[e] = pack_bf16((v[i][2 * e] - mean) * rstd * g[2 * e] + b[2 * e],
                          (v[i][2 * e + 1] - mean) * rstd * g[2 * e + 1] + b[2 * e + 1]);
      *reinterpret_cast<uint4*>(orow + vi * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

int layernorm_run(const dd_layernorm_args* a, cudaStream_t stream) {
  DD_CHECK(a != nullptr && a->rows > 0 && a->c > 0, -1, "dd_layernorm: bad args");
  DD_CHECK(a->c % 8 == 0 && a->c <= 2048, -1, "dd_layernorm: C=%d must be a multiple of 8 and <= 2048", a->c);
  const int nvec = a->c >> 3;
  const int vpl = (nvec + 31) / 32;
  const int threads = 256;
  const int grid = (int)(((long long)a->rows * 32 + threads - 1) / threads);
  const bf16* x = reinterpret_cast<const bf16*>(a->x);
  bf16* out = reinterpret_cast<bf16*>(a->out);
#define DD_LN(V)                                                                                         \
  layernorm_kernel<V><<<grid, threads, 0, stream>>>(x, a->x_ld, out, a->out_ld, a->gamma, a->beta, a->rows, \
                                                    a->c, a->eps)
  if (vpl <= 1) DD_LN(1);
  else if (vpl == 2) DD_LN(2);
  else if (vpl == 3) DD_LN(3);
  else if (vpl <= 5) DD_LN(5);
  else DD_LN(8);
#undef DD_LN
  DD_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

}  // namespace dd
