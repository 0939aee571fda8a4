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
// GroupNorm pass 1: per (image, channel) sum / sum of squares.  A CTA owns a slab of rows of one image; threads
// that share a channel octet reduce through shared memory and the CTA stores ONE partial row:
// partial[img][cta][C][2] (fp32).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
gn_stats_kernel(const bf16* __restrict__ x1, long long ld1, int C1, const bf16* __restrict__ x2,
                long long ld2, int C, int HW, int rows_per_cta, float* __restrict__ stats) {
  extern __shared__ float red[];  // [rpi][C][2]
  const int img = blockIdx.y;
  const int tpr = C >> 3;                    // threads per row (8 channels each)
  const int rpi = blockDim.x / tpr;          // rows per iteration
  const int lane_c = threadIdx.x % tpr;
  const int sub = threadIdx.x / tpr;
  const int c0 = lane_c * 8;
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(HW, r_begin + rows_per_cta);
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  if (sub < rpi) {
    const bool from1 = c0 < C1;
    const bf16* base = from1 ? x1 + c0 : x2 + (c0 - C1);
    const long long ld = from1 ? ld1 : ld2;
    int r = r_begin + sub;
    // two independent 16-byte loads in flight per thread
    for (; r + rpi < r_end; r += 2 * rpi) {
      const uint4 v0 = *reinterpret_cast<const uint4*>(base + ((long long)img * HW + r) * ld);
      const uint4 v1 = *reinterpret_cast<const uint4*>(base + ((long long)img * HW + r + rpi) * ld);
      const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float2 f = unpack_bf16(w[e]);
        const int j = (e & 3) * 2;
        s[j] += f.x; q[j] += f.x * f.x;
        s[j + 1] += f.y; q[j + 1] += f.y * f.y;
      }
    }
    for (; r < r_end; r += rpi) {
      const uint4 v = *reinterpret_cast<const uint4*>(base + ((long long)img * HW + r) * ld);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16(w[e]);
        s[2 * e] += f.x; q[2 * e] += f.x * f.x;
        s[2 * e + 1] += f.y; q[2 * e + 1] += f.y * f.y;
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
  // one partial per CTA, plain stores: no atomics, no pre-zeroing, bit-reproducible run to run
  float* part = stats + ((long long)img * gridDim.x + blockIdx.x) * 2 * C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < rpi; ++k) acc += red[(size_t)k * 2 * C + i];
    part[i] = acc;
  }
}

// pass 1b: reduce the per-CTA partials, group statistics -> per-(image, channel) affine y = x*scale + shift.
// One warp per group (32 groups -> 1024 threads), lanes stride over the group's channels.
__global__ void __launch_bounds__(1024)
gn_finalize_kernel(const float* __restrict__ partial, int n_part, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ ss, int C, int groups, int HW, float eps) {
  const int img = blockIdx.x;
  const int cpg = C / groups;
  const int lane = threadIdx.x & 31;
  for (int g = threadIdx.x >> 5; g < groups; g += blockDim.x >> 5) {
    float s = 0.f, q = 0.f;
    for (int c = lane; c < cpg; c += 32) {
      const int ch = g * cpg + c;
      for (int k = 0; k < n_part; ++k) {
        const float2 v = *reinterpret_cast<const float2*>(partial + (((long long)img * n_part + k) * C + ch) * 2);
        s += v.x;
        q += v.y;
      }
    }
    s = warp_sum(s);
    q = warp_sum(q);
    const float inv_n = 1.f / (float)(cpg * HW);
    const float mean = s * inv_n;
    const float var = fmaxf(q * inv_n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    for (int c = lane; c < cpg; c += 32) {
      const int ch = g * cpg + c;
      const float ga = gamma[ch] * rstd;
      *reinterpret_cast<float2*>(ss + ((long long)img * C + ch) * 2) = make_float2(ga, beta[ch] - mean * ga);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// GroupNorm pass 2: normalise + affine (+ SiLU), write compact or padded layout
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
gn_apply_kernel(const bf16* __restrict__ x1, long long ld1, int C1, const bf16* __restrict__ x2,
                long long ld2, int C, int H, int W, const float* __restrict__ ss, int silu, int padded,
                bf16* __restrict__ out, long long out_ld, int rows_per_cta) {
  const int img = blockIdx.y;
  const int HW = H * W;
  const int tpr = C >> 3;
  const int rpi = blockDim.x / tpr;
  const int lane_c = threadIdx.x % tpr;
  const int sub = threadIdx.x / tpr;
  if (sub >= rpi) return;
  const int c0 = lane_c * 8;
  const bool from1 = c0 < C1;
  const bf16* base = from1 ? x1 + c0 : x2 + (c0 - C1);
  const long long ld = from1 ? ld1 : ld2;
  const int Wp = W + 1;
  const int rows_img = padded ? (H + 1) * Wp : HW;
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(rows_img, r_begin + rows_per_cta);
  float sc[8], sh[8];
  {
    const float4* p4 = reinterpret_cast<const float4*>(ss + ((long long)img * C + c0) * 2);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float4 t = p4[e];
      sc[2 * e] = t.x; sh[2 * e] = t.y; sc[2 * e + 1] = t.z; sh[2 * e + 1] = t.w;
    }
  }
  for (int r = r_begin + sub; r < r_end; r += rpi) {
    int src = r;
    bool live = true;
    if (padded) {
      const int hp = r / Wp, wp = r - hp * Wp;
      live = (hp < H) && (wp < W);
      src = hp * W + wp;
    }
    uint4 o = make_uint4(0, 0, 0, 0);
    if (live) {
      const uint4 v = *reinterpret_cast<const uint4*>(base + ((long long)img * HW + src) * ld);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
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
    *reinterpret_cast<uint4*>(out + ((long long)img * rows_img + r) * out_ld + c0) = o;
  }
}

// ---------------------------------------------------------------------------------------------------
// Fused GroupNorm: ONE pass over HBM.  A thread-block cluster of K CTAs owns one image; every CTA stages its slab of rows
// in shared memory while accumulating per-channel sums, the 32 group sums of the K CTAs are exchanged through
// distributed shared memory (cluster barrier + ld.shared::cluster, summed in rank order -> bit-reproducible), and
// the slab is normalised straight out of shared memory.  HBM traffic: 1 read + 1 write of the activation instead of
// 2 reads + 1 write and three launches (gn_stats / gn_finalize / gn_apply above remain as the path for images that do
// not fit K <= 8 slabs of shared memory, i.e. the 448x800 configuration and the 960-channel L0 concat).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_addr, uint32_t rank) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(mapa_shared(local_addr, rank)) : "memory");
  return v;
}

__global__ void __launch_bounds__(512)
gn_fused_kernel(const bf16* __restrict__ x1, long long ld1, int C1, const bf16* __restrict__ x2, long long ld2, int C,
                int H, int W, int groups, int rows_per_cta, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, int silu, int padded, bf16* __restrict__ out, long long out_ld,
                int K) {
  extern __shared__ __align__(16) uint8_t gsm[];
  const int HW = H * W;
  const int rank = (int)(blockIdx.x % K);      // == %cluster_ctarank for cluster dims (K, 1, 1)
  const int img = (int)(blockIdx.x / K);
  const int tpr = C >> 3;                      // threads per row (8 channels each)
  const int rpi = blockDim.x / tpr;            // rows per iteration
  const int lane_c = threadIdx.x % tpr;
  const int sub = threadIdx.x / tpr;
  const int c0 = lane_c * 8;
  const int r_begin = min(HW, rank * rows_per_cta);
  const int r_end = min(HW, r_begin + rows_per_cta);
  const int n_rows = r_end - r_begin;
  // smem: slab [rows_per_cta][C] bf16 | red [C][2] fp32 (per-channel sums, later scale/shift) | part [groups][2] fp32
  bf16* slab = reinterpret_cast<bf16*>(gsm);
  float* red = reinterpret_cast<float*>(gsm + (size_t)rows_per_cta * C * 2);
  float* part = red + 2 * C;

  // ---- phase 1: global -> shared, per-thread channel sums (4 independent 16-byte loads in flight) ----
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  if (sub < rpi) {
    const bool from1 = c0 < C1;
    const bf16* base = (from1 ? x1 + c0 : x2 + (c0 - C1)) + ((long long)img * HW + r_begin) * (from1 ? ld1 : ld2);
    const long long ld = from1 ? ld1 : ld2;
    int r = sub;
    for (; r + 3 * rpi < n_rows; r += 4 * rpi) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(base + (long long)(r + u * rpi) * ld);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        *reinterpret_cast<uint4*>(slab + (size_t)(r + u * rpi) * C + c0) = v[u];
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16(w[e]);
          s[2 * e] += f.x; q[2 * e] += f.x * f.x;
          s[2 * e + 1] += f.y; q[2 * e + 1] += f.y * f.y;
        }
      }
    }
    for (; r < n_rows; r += rpi) {
      const uint4 v = *reinterpret_cast<const uint4*>(base + (long long)r * ld);
      *reinterpret_cast<uint4*>(slab + (size_t)r * C + c0) = v;
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16(w[e]);
        s[2 * e] += f.x; q[2 * e] += f.x * f.x;
        s[2 * e + 1] += f.y; q[2 * e + 1] += f.y * f.y;
      }
    }
  }
  // ---- per-channel sums of this CTA: the row-slices add into red[] one after another (fixed order, no atomics) ----
  for (int k = 0; k < rpi; ++k) {
    if (sub == k) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float2* d = reinterpret_cast<float2*>(red + 2 * (c0 + e));
        *d = (k == 0) ? make_float2(s[e], q[e]) : make_float2(d->x + s[e], d->y + q[e]);
      }
    }
    __syncthreads();
  }
  // ---- per-group partial sums of this CTA ----
  const int cpg = C / groups;
  if ((int)threadIdx.x < groups) {
    float gs = 0.f, gq = 0.f;
    for (int c = 0; c < cpg; ++c) {
      const float2 v = *reinterpret_cast<const float2*>(red + 2 * ((int)threadIdx.x * cpg + c));
      gs += v.x;
      gq += v.y;
    }
    *reinterpret_cast<float2*>(part + 2 * threadIdx.x) = make_float2(gs, gq);
  }
  // ---- phase 2: exchange the group sums across the cluster ----
  if (K > 1) cluster_sync_all(); else __syncthreads();
  const float inv_n = 1.f / (float)(cpg * HW);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    float gs = 0.f, gq = 0.f;
    if (K > 1) {
      const uint32_t a = smem_u32(part + 2 * g);
      for (int k = 0; k < K; ++k) {
        gs += ld_dsmem_f32(a, (uint32_t)k);
        gq += ld_dsmem_f32(a + 4, (uint32_t)k);
      }
    } else {
      gs = part[2 * g];
      gq = part[2 * g + 1];
    }
    const float mean = gs * inv_n;
    const float var = fmaxf(gq * inv_n - mean * mean, 0.f);
    const float ga = gamma[c] * rsqrtf(var + eps);
    *reinterpret_cast<float2*>(red + 2 * c) = make_float2(ga, beta[c] - mean * ga);
  }
  // nobody may leave (or overwrite part[]) while a peer can still read it; also publishes red[] inside the CTA
  if (K > 1) cluster_sync_all(); else __syncthreads();

  // ---- phase 3: normalise + affine (+ SiLU) out of shared memory, compact or padded output ----
  if (sub >= rpi) return;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float2 t = *reinterpret_cast<const float2*>(red + 2 * (c0 + e));
    sc[e] = t.x;
    sh[e] = t.y;
  }
  const int Wp = W + 1;
  const int rows_img = padded ? (H + 1) * Wp : HW;
  // output rows owned by this CTA: those of its source rows plus the zero halo entries that follow them
  auto out_index = [&](int src) { return padded ? (src / W) * Wp + (src % W) : src; };
  const int o_begin = (r_begin >= HW) ? rows_img : out_index(r_begin);
  const int o_end = (r_end >= HW) ? rows_img : out_index(r_end);
  bf16* obase = out + (long long)img * rows_img * out_ld + c0;
  for (int r = o_begin + sub; r < o_end; r += rpi) {
    int src = r;
    bool live = true;
    if (padded) {
      const int hp = r / Wp, wp = r - hp * Wp;
      live = (hp < H) && (wp < W);
      src = hp * W + wp;
    }
    uint4 o = make_uint4(0, 0, 0, 0);
    if (live) {
      const uint4 v = *reinterpret_cast<const uint4*>(slab + (size_t)(src - r_begin) * C + c0);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
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
    *reinterpret_cast<uint4*>(obase + (long long)r * out_ld) = o;
  }
}

// number of stats CTAs (= partial rows) per image: enough to fill the machine, every CTA non-empty
static int groupnorm_partials(int n_img, int HW, int rpi) {
  const int sms = num_sms();
  int per_img = (2 * sms + n_img - 1) / n_img;
  if (per_img < 1) per_img = 1;
  int rows_per_cta = (HW + per_img - 1) / per_img;
  const int min_rows = rpi * 4;
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
  return (long long)n_img * C * 2 * (groupnorm_partials(n_img, HW, rpi) + 1);
}

int groupnorm_run(const dd_groupnorm_args* a, cudaStream_t stream) {
  DD_CHECK(a != nullptr, -1, "dd_groupnorm: null args");
  const int C = a->c1 + a->c2;
  DD_CHECK(a->n_img > 0 && a->h > 0 && a->w > 0 && C > 0, -1, "dd_groupnorm: bad shape");
  DD_CHECK(C % 8 == 0 && a->c1 % 8 == 0, -1, "dd_groupnorm: channels must be multiples of 8 (C=%d c1=%d)", C, a->c1);
  DD_CHECK(C % a->groups == 0, -1, "dd_groupnorm: C=%d not divisible by groups=%d", C, a->groups);
  DD_CHECK(C <= 4096, -1, "dd_groupnorm: C=%d too large (max 4096)", C);
  DD_CHECK(a->c2 == 0 || a->x2 != nullptr, -1, "dd_groupnorm: x2 missing");
  const int HW = a->h * a->w;
  const int tpr = C >> 3;
  int rpi = 0;
  const int threads = groupnorm_threads(C, &rpi);
  DD_CHECK(rpi >= 1, -1, "dd_groupnorm: C=%d too large", C);
  {
    // fused single-pass path: K CTAs (one cluster) per image, each holding HW/K rows in shared memory
    static const int fused_on = getenv("DD_GN_FUSED") ? atoi(getenv("DD_GN_FUSED")) : 1;
    const size_t extra = (size_t)C * 8 + (size_t)a->groups * 8 + 16;
    const size_t cap2 = (size_t)(233472 / 2 - 1024) - extra;   // two CTAs per SM
    const size_t cap1 = (size_t)232448 - extra;                // one CTA per SM
    const size_t img_bytes = (size_t)HW * C * 2;
    int K = 0;
    for (int k = 1; k <= 8; k *= 2) {
      const size_t slab = (size_t)((HW + k - 1) / k) * C * 2;
      if (slab <= cap2) { K = k; break; }
    }
    if (K == 0 && (img_bytes + 7) / 8 <= cap1) K = 8;
    // small batches: spread an image over more CTAs so the machine fills
    while (K != 0 && K < 8 && a->n_img * K < 2 * num_sms() && HW / (2 * K) >= rpi) K *= 2;
    if (fused_on && K != 0 && a->groups <= 512) {
      const int rows_per_cta = (HW + K - 1) / K;
      const size_t smem = (size_t)rows_per_cta * C * 2 + extra;
      static bool attr = false;
      if (!attr) {
        DD_CUDA(cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        attr = true;
      }
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(a->n_img * K));
      cfg.blockDim = dim3((unsigned)threads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = stream;
      cudaLaunchAttribute attr1[1];
      attr1[0].id = cudaLaunchAttributeClusterDimension;
      attr1[0].val.clusterDim.x = (unsigned)K;
      attr1[0].val.clusterDim.y = 1;
      attr1[0].val.clusterDim.z = 1;
      cfg.attrs = attr1;
      cfg.numAttrs = 1;
      DD_CUDA(cudaLaunchKernelEx(&cfg, gn_fused_kernel, reinterpret_cast<const bf16*>(a->x1), (long long)a->x1_ld, a->c1,
                                 reinterpret_cast<const bf16*>(a->x2), (long long)a->x2_ld, C, a->h, a->w, a->groups,
                                 rows_per_cta, a->gamma, a->beta, a->eps, a->silu, a->padded_out,
                                 reinterpret_cast<bf16*>(a->out), (long long)a->out_ld, K));
      count_launch(1);
      return 0;
    }
  }
  // stats scratch layout: [n_img][per_img][C][2] per-CTA partial sums followed by [n_img][C][2] (scale, shift)
  const int sms = num_sms();
  const int n_part = groupnorm_partials(a->n_img, HW, rpi);
  float* sums = a->stats;
  float* ss = a->stats + (size_t)a->n_img * n_part * C * 2;
  {
    // enough CTAs to fill the machine, few enough that the per-CTA atomics stay negligible
    const int rows_per_cta = (HW + n_part - 1) / n_part;
    dim3 grid(n_part, a->n_img);
    const size_t smem = sizeof(float) * 2 * (size_t)C * rpi;
    static bool attr = false;
    if (!attr) {
      DD_CUDA(cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr = true;
    }
    DD_CHECK(smem <= 96 * 1024, -1, "dd_groupnorm: reduction buffer too large");
    gn_stats_kernel<<<grid, threads, smem, stream>>>(reinterpret_cast<const bf16*>(a->x1), a->x1_ld, a->c1,
                                                     reinterpret_cast<const bf16*>(a->x2), a->x2_ld, C, HW,
                                                     rows_per_cta, sums);
    DD_CUDA(cudaGetLastError());
  }
  gn_finalize_kernel<<<a->n_img, 1024, 0, stream>>>(sums, n_part, a->gamma, a->beta, ss, C, a->groups, HW, a->eps);
  DD_CUDA(cudaGetLastError());
  {
    const int rows_img = a->padded_out ? (a->h + 1) * (a->w + 1) : HW;
    int per_img = (8 * sms + a->n_img - 1) / a->n_img;
    if (per_img < 1) per_img = 1;
    int rows_per_cta = (rows_img + per_img - 1) / per_img;
    if (rows_per_cta < rpi * 4) rows_per_cta = rpi * 4;
    dim3 grid((rows_img + rows_per_cta - 1) / rows_per_cta, a->n_img);
    gn_apply_kernel<<<grid, threads, 0, stream>>>(
        reinterpret_cast<const bf16*>(a->x1), a->x1_ld, a->c1, reinterpret_cast<const bf16*>(a->x2),
        a->x2_ld, C, a->h, a->w, ss, a->silu, a->padded_out, reinterpret_cast<bf16*>(a->out), a->out_ld,
        rows_per_cta);
    DD_CUDA(cudaGetLastError());
  }
  count_launch(3);
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
