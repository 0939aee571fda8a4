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

// ---------------------------------------------------------------------------------------------------
// GroupNorm in ONE pass over HBM: a CTA owns ALL rows of a few whole groups of one image.
//
// The two-kernel form above reads every activation twice (statistics, then apply): 6 bytes per element against the 4 the
// operator needs, and ncu showed the apply pass half MUFU-bound on top (SiLU = ex2 + rcp per element, XU 54 %).  Groups are
// independent of each other, so the image is cut along the CHANNELS: a CTA pulls the [rows x (a few groups)] slab into
// shared memory with 16-byte cp.async copies (row segments of 80-240 bytes; the CTAs owning the neighbouring channel
// ranges of the same rows run next to it, so DRAM still streams whole lines), and everything else happens on chip with
// CTA-level barriers only: group means, then the sum of squared deviations from the mean (a true two-pass variance -- no
// E[x^2] - mean^2 cancellation), then normalise + affine (+ SiLU as x/2 (1 + tanh(x/2)): one MUFU op per element) straight
// into the compact or zero-haloed output.  One read + one write of HBM, no scratch buffer, no atomics, fixed reduction
// orders (bit-reproducible).  A first version that cut the image along the ROWS over a thread-block cluster (partials
// exchanged through distributed shared memory) was 2-4x SLOWER than the two-kernel form at cluster sizes 8 / 16
// (profiles/r02_gn_notes.md): co-scheduling 16 CTAs and three cluster barriers cost more than the second read.
// The channel range of a CTA starts and ends on a multiple of 8 channels (16-byte vectors); inside it an octet of channels
// may straddle two groups (10 / 20 / 30 / 60 channels per group), which the (lo, hi) bookkeeping below resolves.
// ---------------------------------------------------------------------------------------------------
struct GnSlabDev {
  const bf16* x1; const bf16* x2; bf16* out; const float* gamma; const float* beta;
  long long ld1, ld2, out_ld;
  int C1, C, H, W, groups, cpg, g_per_cta, silu, padded, rows_out_img;
  float eps;
};
constexpr int GN_SLAB_THREADS = 256;            // (512 threads for the slabs that leave one CTA per SM: measured 27 % slower)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
// 16-byte load from a shared-window address (the generic-pointer form made the compiler split it and re-derive the window)
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(GN_SLAB_THREADS, 3)
gn_slab_kernel(const GnSlabDev p) {
  constexpr int NT = GN_SLAB_THREADS;
  extern __shared__ __align__(128) uint8_t gsm[];
  // layout: [group mean: 64 floats][group rstd: 64 floats][scratch: 256 x float2][slab: rows x width bf16]
  float* gmean = reinterpret_cast<float*>(gsm);
  float* grstd = gmean + 64;
  float2* scratch = reinterpret_cast<float2*>(gsm + 512);
  uint8_t* slab = gsm + 512 + GN_SLAB_THREADS * 8;
  const int img = blockIdx.y;
  const int g0 = blockIdx.x * p.g_per_cta;
  const int ng = min(p.g_per_cta, p.groups - g0);       // groups of this CTA
  const int ca = g0 * p.cpg;                             // first channel (a multiple of 8 by construction)
  const int width = ng * p.cpg;                          // channels of this CTA (a multiple of 8)
  const int wv = width >> 3;                             // 16-byte vectors per row
  const int HW = p.H * p.W;
  const int rpi = NT / wv;                               // rows per iteration
  const int j = threadIdx.x % wv, sub = threadIdx.x / wv;
  const bool on = sub < rpi;
  const int c0 = ca + 8 * j;                             // my octet of channels
  const uint32_t pitch = (uint32_t)width * 2;
  // ---- slab in: my column of 16-byte vectors, every rpi-th row
  if (on) {
    const bool from1 = c0 < p.C1;
    const bf16* src = (from1 ? p.x1 + c0 : p.x2 + (c0 - p.C1)) + (long long)img * HW * (from1 ? p.ld1 : p.ld2);
    const long long ld = from1 ? p.ld1 : p.ld2;
    const uint32_t dst = smem_u32(slab) + 16u * (uint32_t)j;
    for (int r = sub; r < HW; r += rpi) cp_async16(dst + (uint32_t)r * pitch, src + (long long)r * ld);
  }
  // the octet's group split: it lies in at most two groups (cpg >= 8), "lo" and "hi"
  const int glo = on ? c0 / p.cpg : g0;
  const int n_lo = on ? min(8, (glo + 1) * p.cpg - c0) : 8;      // channels of the octet that belong to the lo group
  cp_async_wait_all();
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // CTA-level reduction of per-thread (lo, hi) values to per-group totals (one warp per group, fixed order)
  auto group_totals = [&](const float (&v)[8], float* dst, bool finish_rstd) {
    float lo = 0.f, hi = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (e < n_lo) lo += v[e]; else hi += v[e];
    }
    if (on) scratch[sub * wv + j] = make_float2(lo, hi);
    __syncthreads();
    for (int gi = warp; gi < ng; gi += NT / 32) {
      const int g = g0 + gi;
      const int ja = (g * p.cpg - ca) >> 3, jb = ((g + 1) * p.cpg - 1 - ca) >> 3;     // octets overlapping the group
      const int n_oct = jb - ja + 1;
      float acc = 0.f;
      for (int i = lane; i < n_oct * rpi; i += 32) {
        const int sb = i / n_oct, jj = ja + (i - sb * n_oct);
        const float2 f = scratch[sb * wv + jj];
        acc += ((ca + 8 * jj) / p.cpg == g) ? f.x : f.y;
      }
      acc = warp_sum(acc);
      if (lane == 0) {
        const float inv_n = 1.f / (float)(p.cpg * HW);
        dst[gi] = finish_rstd ? rsqrtf(acc * inv_n + p.eps) : acc * inv_n;
      }
    }
    __syncthreads();
  };
  // ---- ONE pass over the slab for both moments.  Every channel is centred on a pivot -- its value in row 0 of the image, the
  //      same for all threads that share the channel -- and a thread accumulates A = sum(x - s) and B = sum((x - s)^2) over its
  //      rows as packed FADD2 / FFMA2 (ncu: the kernel is issue-bound, and the two separate passes were 40 % of its
  //      instructions).  With n rows per thread:   sum x = n s + A,   sum (x - mu)^2 = B + 2 (s - mu) A + n (s - mu)^2   -- exact
  //      algebra; the pivot keeps every term at the scale of the group's spread (no E[x^2] - mean^2 cancellation).
  const uint32_t my_s = smem_u32(slab) + 16u * (uint32_t)j;
  float piv[8], av[8], bv[8];
  float n_rows = 0.f;
  {
    uint64_t NS[4], A2[4], B2[4];
    {
      const uint4 v = lds128(my_s);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16(w[e]);
        piv[2 * e] = f.x;
        piv[2 * e + 1] = f.y;
        NS[e] = pack_f32x2(-f.x, -f.y);
        A2[e] = 0ull;
        B2[e] = 0ull;
      }
    }
    if (on) {
      int cnt = 0;
#pragma unroll 4
      for (int r = sub; r < HW; r += rpi, ++cnt) {
        const uint4 v = lds128(my_s + (uint32_t)r * pitch);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16(w[e]);
          const uint64_t D = add_f32x2(pack_f32x2(f.x, f.y), NS[e]);
          A2[e] = add_f32x2(A2[e], D);
          B2[e] = fma_f32x2(D, D, B2[e]);
        }
      }
      n_rows = (float)cnt;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      unpack_f32x2(A2[e], av[2 * e], av[2 * e + 1]);
      unpack_f32x2(B2[e], bv[2 * e], bv[2 * e + 1]);
    }
  }
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = fmaf(n_rows, piv[e], av[e]);      // this thread's share of sum x per channel
  group_totals(acc, gmean, false);
  float mu[8];                               // mean of each channel's group
  {
    const float mean_lo = gmean[glo - g0], mean_hi = gmean[min(glo - g0 + 1, ng - 1)];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      mu[e] = e < n_lo ? mean_lo : mean_hi;
      const float dm = piv[e] - mu[e];
      acc[e] = bv[e] + dm * (2.f * av[e] + n_rows * dm);                  // ... of sum (x - mu)^2
    }
  }
  group_totals(acc, grstd, true);
  if (!on) return;
  // y = x * sc + sh; with SiLU the halves are kept: h = y / 2, silu(y) = h * tanh(h) + h (one MUFU op per element)
  uint64_t SC[4], SH[4];
  {
    const float rstd_lo = grstd[glo - g0], rstd_hi = grstd[min(glo - g0 + 1, ng - 1)];
    const float4 a0 = *reinterpret_cast<const float4*>(p.gamma + c0), a1 = *reinterpret_cast<const float4*>(p.gamma + c0 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(p.beta + c0), b1 = *reinterpret_cast<const float4*>(p.beta + c0 + 4);
    const float ga[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const float half = p.silu ? 0.5f : 1.f;
    float sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sc[e] = half * ga[e] * (e < n_lo ? rstd_lo : rstd_hi);
      sh[e] = half * be[e] - mu[e] * sc[e];
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      SC[e] = pack_f32x2(sc[2 * e], sc[2 * e + 1]);
      SH[e] = pack_f32x2(sh[2 * e], sh[2 * e + 1]);
    }
  }
  // ---- normalise the image's output rows from shared memory (the padded layout interleaves zero halo rows / columns)
  bf16* optr = p.out + ((size_t)img * p.rows_out_img + sub) * p.out_ld + c0;
  const size_t ostep = (size_t)rpi * p.out_ld;
  const int Wp = p.W + 1;
  int hp = 0, wp = sub;                      // padded (row, column) of output row rr, carried without divisions
  if (p.padded) { hp = sub / Wp; wp = sub - hp * Wp; }
  const int dh = rpi / Wp, dw = rpi - dh * Wp;
  for (int rr = sub; rr < p.rows_out_img; rr += rpi, optr += ostep) {
    int src = rr;
    bool live = true;
    if (p.padded) {
      live = (hp < p.H) && (wp < p.W);
      src = hp * p.W + wp;
      hp += dh;
      wp += dw;
      if (wp >= Wp) { wp -= Wp; ++hp; }
    }
    uint4 o = make_uint4(0u, 0u, 0u, 0u);   // halo rows / columns of the padded layout are zeros
    if (live) {
      const uint4 v = lds128(my_s + (uint32_t)src * pitch);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16(w[e]);
        const uint64_t Y = fma_f32x2(pack_f32x2(f.x, f.y), SC[e], SH[e]);
        float a, b;
        unpack_f32x2(Y, a, b);
        if (p.silu) {
          const uint64_t T = pack_f32x2(tanh_approx(a), tanh_approx(b));
          unpack_f32x2(fma_f32x2(Y, T, Y), a, b);
        }
        pk[e] = pack_bf16(a, b);
      }
      o = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    *reinterpret_cast<uint4*>(optr) = o;
  }
}

// groups per CTA of the single-pass kernel (0 = the two-kernel form must be used): the channel range must start on a
// multiple of 8 channels, an octet may straddle at most two groups, and the [rows x channels] slab must leave room for two
// CTAs per SM; among the admissible sizes the largest one of at most 64 KB is taken as long as the grid still has
// DD_GN_MIN_CTAS_PER_SM CTAs per SM, else the smallest.
#ifndef DD_GN_MIN_CTAS_PER_SM
#define DD_GN_MIN_CTAS_PER_SM 2      // CTAs per SM the plan keeps when it widens the slabs (A/B hook; 1 / 2 / 3 / 4 / 6 / 8 measured: profiles/r02_groupnorm_slab_width.txt)
#endif
static int gn_slab_plan(int C, int groups, int HW, int n_img, size_t* smem_bytes) {
  const int cpg = C / groups;
  if (cpg < 8 || groups > 64 * 64) return 0;
  int unit = 1;
  while ((unit * cpg) % 8 != 0) ++unit;                 // 1, 2 or 4 groups
  if (unit > groups) return 0;
  const size_t hdr = 512 + GN_SLAB_THREADS * 8;
  const size_t limit = 116 * 1024;
  int best = 0;
  for (int g = unit; g <= groups && g <= 64; g += unit) {
    const int width = g * cpg;
    if (width / 8 > GN_SLAB_THREADS) break;
    const size_t need = hdr + (size_t)HW * width * 2;
    if (need > limit) break;
    const long long ctas = (long long)n_img * ((groups + g - 1) / g);
    if (best == 0 || (need <= 64 * 1024 + hdr && ctas >= (long long)DD_GN_MIN_CTAS_PER_SM * num_sms())) best = g;
  }
  if (best == 0) return 0;
  *smem_bytes = hdr + (size_t)HW * best * cpg * 2;
  return best;
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
  // ---- single-pass kernel when a [rows x few groups] slab of the image fits shared memory (see gn_slab_kernel) ----
  if (!a->two_pass && a->x1_ld % 8 == 0 && (a->c2 == 0 || a->x2_ld % 8 == 0) && a->out_ld % 8 == 0) {
    size_t smem = 0;
    const int gpc = gn_slab_plan(C, a->groups, HW, a->n_img, &smem);
    if (gpc > 0) {
      GnSlabDev p;
      p.x1 = reinterpret_cast<const bf16*>(a->x1); p.x2 = reinterpret_cast<const bf16*>(a->x2);
      p.out = reinterpret_cast<bf16*>(a->out); p.gamma = a->gamma; p.beta = a->beta;
      p.ld1 = a->x1_ld; p.ld2 = a->x2_ld; p.out_ld = a->out_ld;
      p.C1 = a->c1; p.C = C; p.H = a->h; p.W = a->w; p.groups = a->groups; p.cpg = C / a->groups; p.g_per_cta = gpc;
      p.silu = a->silu; p.padded = a->padded_out; p.rows_out_img = a->padded_out ? (a->h + 1) * (a->w + 1) : HW;
      p.eps = a->eps;
      dim3 grid((a->groups + gpc - 1) / gpc, a->n_img);
      if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(gn_slab_kernel), 116 * 1024)) return rc;
      gn_slab_kernel<<<grid, GN_SLAB_THREADS, smem, stream>>>(p);
      DD_CUDA(cudaGetLastError());
      count_launch(1);
      return 0;
    }
  }
  // ---- two passes (statistics, then apply) for images whose slabs do not fit shared memory ----
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

// ---------------------------------------------------------------------------------------------------
// LayerNorm, packed form for the row widths of the path (C = 8 * LPR * VPL): 32 / LPR rows per warp, LPR lanes per row, VPL
// 16-byte vectors per lane with EVERY lane busy (320 channels: 4 rows per warp, 8 lanes x 5 vectors; 640: 2 rows, 16 x 5;
// 1280: 1 row, 32 x 5; 768 (CLIP): 1 row, 32 x 3).  ncu on the generic kernel above at 320 channels: 276 warp instructions per
// row, issue slots 76 % busy at 3.4 TB/s -- instruction-bound (a quarter of the lanes idle in the second vector, scalar
// arithmetic, per-row parameter loads).  Here sums, deviations, squares and the affine map run as packed FADD2 / FFMA2, the
// rows of a warp share each shuffle instruction, and the deviations stay in registers between the variance and the output.
// ---------------------------------------------------------------------------------------------------
template <int LPR, int VPL>
__global__ void __launch_bounds__(256)      // 70 registers, 3 CTAs per SM (capping at 64 for a 4th CTA measured 5 % slower)
layernorm_packed_kernel(const bf16* __restrict__ x, long long x_ld, bf16* __restrict__ out, long long out_ld,
                        const float* __restrict__ gamma, const float* __restrict__ beta, int rows, float eps) {
  constexpr int RPW = 32 / LPR;                 // rows per warp
  constexpr int C = 8 * LPR * VPL;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane - sub * LPR;
  const long long row = ((long long)((blockIdx.x * blockDim.x + threadIdx.x) >> 5)) * RPW + sub;
  const bool live = row < rows;
  const bf16* xr = x + (live ? row : 0) * x_ld + l * 8;
  uint64_t v[VPL][4];
  uint64_t S = 0ull;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const uint4 u = live ? *reinterpret_cast<const uint4*>(xr + i * (LPR * 8)) : make_uint4(0u, 0u, 0u, 0u);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_bf16(w[e]);
      v[i][e] = pack_f32x2(f.x, f.y);
      S = add_f32x2(S, v[i][e]);
    }
  }
  float s0, s1;
  unpack_f32x2(S, s0, s1);
  float sum = s0 + s1;
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * (1.f / (float)C);
  const uint64_t NM = pack_f32x2(-mean, -mean);
  uint64_t Q = 0ull;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[i][e] = add_f32x2(v[i][e], NM);         // deviations from the mean, kept for the output
      Q = fma_f32x2(v[i][e], v[i][e], Q);
    }
  }
  unpack_f32x2(Q, s0, s1);
  float sq = s0 + s1;
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq * (1.f / (float)C) + eps);
  const uint64_t R = pack_f32x2(rstd, rstd);
  if (!live) return;
  bf16* orow = out + row * out_ld + l * 8;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (l + i * LPR) * 8;
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + c), g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + c), b1 = *reinterpret_cast<const float4*>(beta + c + 4);
    const uint64_t G[4] = {pack_f32x2(g0.x, g0.y), pack_f32x2(g0.z, g0.w), pack_f32x2(g1.x, g1.y), pack_f32x2(g1.z, g1.w)};
    const uint64_t B[4] = {pack_f32x2(b0.x, b0.y), pack_f32x2(b0.z, b0.w), pack_f32x2(b1.x, b1.y), pack_f32x2(b1.z, b1.w)};
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float y0, y1;
      unpack_f32x2(fma_f32x2(fma_f32x2(v[i][e], R, 0ull), G[e], B[e]), y0, y1);   // ((x - mean) * rstd) * gamma + beta
      pk[e] = pack_bf16(y0, y1);
    }
    *reinterpret_cast<uint4*>(orow + i * (LPR * 8)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
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
  // the packed kernel for the widths of the path (transformer levels 320 / 640 / 1280, CLIP 768), the generic one otherwise
#define DD_LNP(LPR, VPL)                                                                                            \
  do {                                                                                                              \
    const long long warps = ((long long)a->rows + (32 / LPR) - 1) / (32 / LPR);                                     \
    layernorm_packed_kernel<LPR, VPL><<<(unsigned)((warps * 32 + threads - 1) / threads), threads, 0, stream>>>(    \
        x, a->x_ld, out, a->out_ld, a->gamma, a->beta, a->rows, a->eps);                                            \
    DD_CUDA(cudaGetLastError());                                                                                    \
    count_launch(1);                                                                                                \
    return 0;                                                                                                       \
  } while (0)
  if (a->c == 320) DD_LNP(8, 5);
  if (a->c == 640) DD_LNP(16, 5);
  if (a->c == 1280) DD_LNP(32, 5);
  if (a->c == 768) DD_LNP(32, 3);
#undef DD_LNP
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
