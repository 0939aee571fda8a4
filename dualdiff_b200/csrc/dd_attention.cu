// dualdiff_b200 — flash-style fused attention on tcgen05 / TMEM / TMA (sm_100a).
//
// One kernel serves the four attentions of the DualDiff step:
//   self        (diffusers BasicTransformerBlock.attn1,  networks/blocks.py:166-172)
//   text-cross  (attn2, keys = [camera token | 77 text tokens | box tokens], blocks.py:182-188)
//   cross-view  (attn4, networks/blocks.py:190-222): each view attends to its two ring neighbours; the two
//               softmax-attentions are computed back to back and SUMMED in the epilogue, so the reference's
//               torch.cat of 12 (view, neighbour) pairs, the duplicated projections and the CPU-mask gather
//               disappear (out = A_left + A_right; to_out then adds 2*b_o, see packing.py)
//   SFA         (networks/txt_con_fusion.py:110-177): 320-ch condition feature map queries the 77 text tokens.
//
// CTA = (128 query rows, one head, one image).  S = Q K^T and O_tile = P V run on the tensor cores with
// fp32 accumulators in TMEM; the 128 softmax threads own one query row each (TMEM lane == row), so the online
// softmax needs no shuffles.  K/V tiles arrive by TMA through a 2-stage mbarrier ring; V is consumed as an
// MN-major UMMA operand straight from its row-major [key, dv] layout (no transpose pass).
#include <stdlib.h>

#include "dd_api_internal.h"
#include "dd_common.cuh"

namespace dd {

__device__ __forceinline__ float fast_exp2(float x) {  // inputs are <= 0 after the max subtraction; ftz is fine
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
static constexpr int ATT_BM = 128;
static constexpr int ATT_BN = 128;
static constexpr int ATT_THREADS = 160;     // 4 softmax warps + 1 TMA/UMMA warp
static constexpr int CHUNK_BYTES = 128 * 128;  // 128 rows x 64 bf16 (one SWIZZLE_128B column chunk)

struct AttnDev {
  int Lq, Lk, n_src, n_kv_tiles;
  const int* kv_map;  // [n_img * n_src] kv image per (query image, source) or nullptr (identity)
  float scale_log2e;
  bf16* out;
  long long out_ld;
  int q_col0, k_col0, v_col0, q_hs, k_hs, v_hs, o_hs;
};

template <int DQK, int DV, int DVP, int STAGES, int MINB>
__global__ void __launch_bounds__(ATT_THREADS, MINB)
attn_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnDev p) {
  constexpr int QCH = (DQK + 63) / 64;
  constexpr int VCH = (DVP + 63) / 64;
  constexpr int KV_STAGE_BYTES = (QCH + VCH) * CHUNK_BYTES;
  constexpr int TMEM_COLS = (128 + DVP <= 256) ? 256 : 512;
  constexpr uint32_t IDESC_S = umma_idesc_bf16(ATT_BM, ATT_BN, 0, 0);
  constexpr uint32_t IDESC_O = umma_idesc_bf16(ATT_BM, DVP, 0, 1);  // B (=V) is MN-major

  // No static shared memory: the dynamic segment then starts at the CTA's (1024-byte aligned) window, so the
  // swizzled tiles need no alignment slack and two CTAs fit one SM for head_dim 40.  Barriers live at the tail.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t sQ = smem_base;
  const uint32_t sKV = sQ + QCH * CHUNK_BYTES;
  const uint32_t sP = sKV + STAGES * KV_STAGE_BYTES;
  const uint32_t bar0 = sP + 2 * CHUNK_BYTES;
  uint32_t* tmem_ptr_gen = reinterpret_cast<uint32_t*>(smem_raw + (bar0 - smem_base) + 96);
  const uint32_t q_full = bar0;
  const uint32_t s_full = bar0 + 8;
  const uint32_t p_full = bar0 + 16;
  const uint32_t o_full = bar0 + 24;
  const uint32_t kv_full = bar0 + 32;   // [STAGES]
  const uint32_t kv_empty = bar0 + 48;  // [STAGES]
  const uint32_t s_free = bar0 + 64;    // softmax threads hold S in registers -> the next Q K^T may overwrite TMEM
  if ((smem_base & 1023u) != 0) __trap();

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, head = blockIdx.y, img = blockIdx.z;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_init(s_free, 128);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(kv_full + 8 * s, 1);
      mbar_init(kv_empty + 8 * s, 1);
    }
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(bar0 + 96, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_gen);
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;
  const int total = p.n_src * p.n_kv_tiles;

  if (warp == 4) {
    // whole warp runs the control flow (warp-uniform -> uniform registers for barrier addresses / descriptors);
    // one elected lane issues the TMA loads and UMMAs
    {
      if (elect_one()) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        // Q tile (once)
        mbar_arrive_expect_tx(q_full, QCH * CHUNK_BYTES);
        for (int c = 0; c < QCH; ++c)
          tma_load_3d(sQ + c * CHUNK_BYTES, &tmQ, q_full, p.q_col0 + head * p.q_hs + c * 64, q_tile * ATT_BM, img);
      }
      __syncwarp();
      auto produce = [&](int g) {
        const int s = g % STAGES;
        const uint32_t ph = (g / STAGES) & 1;
        mbar_wait(kv_empty + 8 * s, ph ^ 1);
        const int src = g / p.n_kv_tiles, jt = g - src * p.n_kv_tiles;
        const int kv_img = p.kv_map ? p.kv_map[img * p.n_src + src] : img;
        const uint32_t sK = sKV + s * KV_STAGE_BYTES;
        const uint32_t sV = sK + QCH * CHUNK_BYTES;
        if (elect_one()) {
          mbar_arrive_expect_tx(kv_full + 8 * s, KV_STAGE_BYTES);
          for (int c = 0; c < QCH; ++c)
            tma_load_3d(sK + c * CHUNK_BYTES, &tmK, kv_full + 8 * s, p.k_col0 + head * p.k_hs + c * 64, jt * ATT_BN, kv_img);
          for (int c = 0; c < VCH; ++c)
            tma_load_3d(sV + c * CHUNK_BYTES, &tmV, kv_full + 8 * s, p.v_col0 + head * p.v_hs + c * 64, jt * ATT_BN, kv_img);
        }
        __syncwarp();
      };
      auto issue_qk = [&](int g) {
        const int s = g % STAGES;
        mbar_wait(kv_full + 8 * s, (g / STAGES) & 1);
        tc_fence_after();
        const uint32_t sK = sKV + s * KV_STAGE_BYTES;
        // S = Q K^T  (both operands K-major, 64-column swizzle chunks)
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < DQK / 16; ++kk) {
            const uint32_t off = (kk >> 2) * CHUNK_BYTES + (kk & 3) * 32;
            umma_bf16(tmem_S, umma_smem_desc(sQ + off, 16, 1024, 2), umma_smem_desc(sK + off, 16, 1024, 2), IDESC_S,
                      kk != 0 ? 1u : 0u);
          }
          umma_commit(s_full);
        }
        __syncwarp();
      };
      produce(0);
      if (STAGES > 1 && total > 1) produce(1);
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int g = 0; g < total; ++g) {
        const int s = g % STAGES;
        // Software pipeline: as soon as the softmax threads have pulled S(g) into registers, S(g+1) = Q K(g+1)^T is
        // issued, so it is ready in TMEM when they come back; P(g) V(g) follows once P(g) has been written.
        if (STAGES > 1 && g + 1 < total) {
          mbar_wait(s_free, g & 1);
          tc_fence_after();
          issue_qk(g + 1);
        }
        mbar_wait(p_full, g & 1);
        tc_fence_after();
        const uint32_t sV = sKV + s * KV_STAGE_BYTES + QCH * CHUNK_BYTES;
        // O += P V : A = P (K-major, 2 chunks of 64 keys), B = V (MN-major: rows = keys, 64-wide dv chunks)
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < ATT_BN / 16; ++kk) {
            const uint32_t offP = (kk >> 2) * CHUNK_BYTES + (kk & 3) * 32;
            const uint32_t offV = kk * 16 * 128;  // 16 key rows of 128 B
            umma_bf16(tmem_O, umma_smem_desc(sP + offP, 16, 1024, 2),
                      umma_smem_desc(sV + offV, CHUNK_BYTES, 1024, 2), IDESC_O,
                      (kk != 0 || (g % p.n_kv_tiles) != 0) ? 1u : 0u);  // O accumulates in TMEM over one source
          }
          umma_commit(kv_empty + 8 * s);
          umma_commit(o_full);
        }
        __syncwarp();
        if (STAGES > 1) {
          if (g + 2 < total) produce(g + 2);   // stage s is free once P(g) V(g) retires (kv_empty)
        } else if (g + 1 < total) {
          produce(g + 1);
          mbar_wait(s_free, g & 1);
          tc_fence_after();
          issue_qk(g + 1);
        }
      }
    }
  } else {
    // ------------------------------- softmax / correction / epilogue -------------------------------
    // One pass over S held in registers.  O accumulates in TMEM; the running max is only advanced (and O rescaled
    // in TMEM) when it grows by more than 2^8, so p = exp2(s - m_ref) stays <= 256 and the correction is rare.
    const int row = threadIdx.x;  // == TMEM lane
    const int q_row = q_tile * ATT_BM + row;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    const float sl2 = p.scale_log2e;
    bf16* orow = p.out + ((long long)img * p.Lq + q_row) * p.out_ld + head * p.o_hs;
    const uint32_t p_row_base = sP + row * 128;
    const uint32_t xr = (uint32_t)(row & 7);
    int g = 0;
    for (int src = 0; src < p.n_src; ++src) {
      float m = -INFINITY, l = 0.f;
      for (int jt = 0; jt < p.n_kv_tiles; ++jt, ++g) {
        mbar_wait(s_full, g & 1);
        tc_fence_after();
        const int key0 = jt * ATT_BN;
        uint32_t sv[ATT_BN];
#pragma unroll
        for (int c = 0; c < ATT_BN; c += 32) tmem_ld_32x32(tmem_S + lane_sel + c, *reinterpret_cast<uint32_t(*)[32]>(&sv[c]));
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_free);             // S(g) now lives in registers
        const int nvalid = p.Lk - key0;  // keys >= nvalid in this tile are padding
        float mx = -INFINITY;
        if (nvalid >= ATT_BN) {
#pragma unroll
          for (int j = 0; j < ATT_BN; ++j) mx = fmaxf(mx, __uint_as_float(sv[j]));
        } else {
#pragma unroll
          for (int j = 0; j < ATT_BN; ++j)
            if (j < nvalid) mx = fmaxf(mx, __uint_as_float(sv[j]));
        }
        const float m_cand = fmaxf(m, mx * sl2);
        const bool grow = __any_sync(0xffffffffu, m_cand > m + 8.f);   // warp-uniform (first tile: m = -inf)
        float alpha = 1.f;
        if (grow) {
          alpha = fast_exp2(m - m_cand);
          m = m_cand;
          l *= alpha;
        }
        // p = exp2(s*scale*log2e - m), packed to bf16 (registers: 64)
        uint32_t pk[ATT_BN / 2];
        float lsum = 0.f;
        if (nvalid >= ATT_BN) {
#pragma unroll
          for (int j = 0; j < ATT_BN; j += 2) {
            const float p0 = fast_exp2(__uint_as_float(sv[j]) * sl2 - m);
            const float p1 = fast_exp2(__uint_as_float(sv[j + 1]) * sl2 - m);
            lsum += p0 + p1;
            pk[j >> 1] = pack_bf16(p0, p1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < ATT_BN; j += 2) {
            const float p0 = (j < nvalid) ? fast_exp2(__uint_as_float(sv[j]) * sl2 - m) : 0.f;
            const float p1 = (j + 1 < nvalid) ? fast_exp2(__uint_as_float(sv[j + 1]) * sl2 - m) : 0.f;
            lsum += p0 + p1;
            pk[j >> 1] = pack_bf16(p0, p1);
          }
        }
        l += lsum;
        // the previous tile's P V must have retired before P is overwritten / O is rescaled
        if (g > 0) {
          mbar_wait(o_full, (g - 1) & 1);
          tc_fence_after();
        }
        if (grow && jt > 0) {
#pragma unroll
          for (int c = 0; c < DVP; c += 16) {
            uint32_t ov[16];
            tmem_ld_32x16(tmem_O + lane_sel + c, ov);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) ov[j] = __float_as_uint(__uint_as_float(ov[j]) * alpha);
            tmem_st_32x16(tmem_O + lane_sel + c, ov);
          }
          tmem_st_wait();
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {  // 16-byte units of this row: 8 per 64-key swizzle chunk
          const uint32_t addr = p_row_base + (u >> 3) * CHUNK_BYTES + ((((uint32_t)u & 7) ^ xr) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * u]), "r"(pk[4 * u + 1]),
                       "r"(pk[4 * u + 2]), "r"(pk[4 * u + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(p_full);
      }
      // epilogue of this source: O / l  (second source of the cross-view attention adds onto the first)
      mbar_wait(o_full, (g - 1) & 1);
      tc_fence_after();
      const float inv = 1.f / l;
#pragma unroll
      for (int c = 0; c < DVP; c += 16) {
        uint32_t ov[16];
        tmem_ld_32x16(tmem_O + lane_sel + c, ov);
        tmem_ld_wait();
        if (q_row < p.Lq) {
#pragma unroll
          for (int hh = 0; hh < 16; hh += 8) {
            if (c + hh + 8 <= DV) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(ov[hh + e]) * inv;
              if (src > 0) {
                const uint4 r = *reinterpret_cast<const uint4*>(orow + c + hh);
                float2 t;
                t = unpack_bf16(r.x); v[0] += t.x; v[1] += t.y;
                t = unpack_bf16(r.y); v[2] += t.x; v[3] += t.y;
                t = unpack_bf16(r.z); v[4] += t.x; v[5] += t.y;
                t = unpack_bf16(r.w); v[6] += t.x; v[7] += t.y;
              }
              *reinterpret_cast<uint4*>(orow + c + hh) =
                  make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            }
          }
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// v2: P stays in tensor memory.
//
// ncu on v1 (profiles/r01_ncu_attn_L0_self.txt + source page): per 128x128 tile and SM the kernel needs ~1024 cycles
// of the 16-lane MUFU pipe, but ALSO ~1030 cycles of the 128 B/clk shared-memory port (K/V TMA writes 32 KB, Q/K
// operand reads 24 KB, P written by the softmax threads 32 KB and read back by the P V UMMAs 32 KB, V reads 12 KB),
// and 12 % of the softmax warps' samples sat on the STS / fence.proxy.async / arrive sequence that publishes P.
// Here P is written with tcgen05.st next to S and O in TMEM and consumed as the TMEM A operand of the P V UMMAs
// (tcgen05.mma [d], [a_tmem], b_desc): half of the shared-memory traffic, the proxy fence and the P staging buffer
// disappear (smem per CTA 112 -> 80 KB at head_dim 40, used for a third K/V stage).  The softmax loop runs the
// row maximum on four independent FMNMX3 chains and the scale/subtract and the row sum as packed FFMA2/FADD2 (half
// the issue slots).  POLY of every 8 exponentials can be evaluated as a degree-3 polynomial on the FMA pipe
// instead of the MUFU pipe (Cody-Waite split, max rel. error 7.5e-5); measured slower on B200 so far (ptxas clusters
// the polynomial work instead of interleaving it with the MUFU stream), so POLY = 0 is what ships.
// Key tiles are BN = 128 (head_dim 40) or 64 wide (head_dim 80/160: S + O + P then fit 256 TMEM columns, so two CTAs
// share an SM and cover each other's barrier round trips: 1.45x on those levels).
// Measured dead ends (profiles/README.md): two softmax warpgroups splitting the columns, chunk-wise software
// pipelining of the TMEM loads, staggering the two CTAs of an SM.  TMEM reads are not the limit
// (profiles/micro/tmem_bench.cu: 350 B/clk/SM from 4 warps, 860 from 16).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// 2^x for a pair of fp32 values on the FMA pipe.  x <= ~8 (lazy running max), clamped below at -126.
__device__ __forceinline__ void exp2_poly_pair(float x0, float x1, float& r0, float& r1) {
  x0 = fmaxf(x0, -126.f);
  x1 = fmaxf(x1, -126.f);
  const uint64_t X = pack_f32x2(x0, x1);
  const uint64_t MAGIC = pack_f32x2(12582912.f, 12582912.f);        // 1.5 * 2^23: low mantissa bits = round(x)
  const uint64_t NMAGIC = pack_f32x2(-12582912.f, -12582912.f);
  const uint64_t MONE = pack_f32x2(-1.f, -1.f);
  const uint64_t T = add_f32x2(X, MAGIC);
  const uint64_t N = add_f32x2(T, NMAGIC);                            // round(x) as a float
  const uint64_t F = fma_f32x2(N, MONE, X);                           // x - round(x) in [-0.5, 0.5]
  uint64_t P = fma_f32x2(F, pack_f32x2(0.05517132207751274f, 0.05517132207751274f),
                         pack_f32x2(0.24261054396629333f, 0.24261054396629333f));
  P = fma_f32x2(P, F, pack_f32x2(0.6932609677314758f, 0.6932609677314758f));
  P = fma_f32x2(P, F, pack_f32x2(0.9999281167984009f, 0.9999281167984009f));
  float t0, t1, p0, p1;
  unpack_f32x2(T, t0, t1);
  unpack_f32x2(P, p0, p1);
  r0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));   // p * 2^round(x)
  r1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

template <int DQK, int DV, int DVP, int BN, int STAGES, int MINB, int POLY, bool PIPE>
__global__ void __launch_bounds__(ATT_THREADS, MINB)
attn_v2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnDev p) {
  constexpr int QCH = (DQK + 63) / 64;
  constexpr int VCH = (DVP + 63) / 64;
  constexpr int Q_CHUNK = ATT_BM * 128;                 // 128 query rows x 64 bf16
  constexpr int K_CHUNK = BN * 128;                     // BN key rows x 64 bf16
  constexpr int KV_STAGE_BYTES = (QCH + VCH) * K_CHUNK;
  constexpr int O_COL = BN;                             // TMEM columns: S [0, BN) | O [BN, BN + DVP) | P (bf16x2) BN/2 columns
  constexpr int P_COL = BN + DVP;
  constexpr int TMEM_NEED = P_COL + BN / 2;
  constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
  constexpr int ISSUER = 4;                             // warp index of the TMA / UMMA warp
  constexpr int NSOFT = 128;                            // softmax threads
  constexpr uint32_t IDESC_S = umma_idesc_bf16(ATT_BM, BN, 0, 0);
  constexpr uint32_t IDESC_O = umma_idesc_bf16(ATT_BM, DVP, 0, 1);  // B (=V) is MN-major
  static_assert(DVP % 16 == 0 && BN % 64 == 0 && P_COL % 8 == 0 && (BN == 64 || BN == 128), "tile geometry");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t sQ = smem_base;
  const uint32_t sKV = sQ + QCH * Q_CHUNK;
  const uint32_t bar0 = sKV + STAGES * KV_STAGE_BYTES;
  uint32_t* tmem_ptr_gen = reinterpret_cast<uint32_t*>(smem_raw + (bar0 - smem_base) + 120);
  const uint32_t q_full = bar0;
  const uint32_t s_full = bar0 + 8;
  const uint32_t p_full = bar0 + 16;
  const uint32_t o_full = bar0 + 24;
  const uint32_t s_free = bar0 + 32;
  const uint32_t kv_full = bar0 + 40;                   // [STAGES <= 4]
  const uint32_t kv_empty = bar0 + 72;                  // [STAGES <= 4]
  if ((smem_base & 1023u) != 0) __trap();

  const int warp = threadIdx.x >> 5;
  const int q_tile = blockIdx.x, head = blockIdx.y, img = blockIdx.z;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, NSOFT);
    mbar_init(o_full, 1);
    mbar_init(s_free, NSOFT);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(kv_full + 8 * s, 1);
      mbar_init(kv_empty + 8 * s, 1);
    }
    fence_mbar_init();
  }
  if (warp == ISSUER) {
    tmem_alloc(bar0 + 120, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_gen);
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + O_COL;
  const uint32_t tmem_P = tmem_base + P_COL;
  const int total = p.n_src * p.n_kv_tiles;
  if (warp == ISSUER) {
    // ---------------- TMA producer + UMMA issuer (warp-uniform control flow, one elected lane issues) ----------------
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      mbar_arrive_expect_tx(q_full, QCH * Q_CHUNK);
      for (int c = 0; c < QCH; ++c)
        tma_load_3d(sQ + c * Q_CHUNK, &tmQ, q_full, p.q_col0 + head * p.q_hs + c * 64, q_tile * ATT_BM, img);
    }
    __syncwarp();
    auto produce = [&](int g) {
      const int s = g % STAGES;
      const uint32_t ph = (g / STAGES) & 1;
      mbar_wait(kv_empty + 8 * s, ph ^ 1);
      const int src = g / p.n_kv_tiles, jt = g - src * p.n_kv_tiles;
      const int kv_img = p.kv_map ? p.kv_map[img * p.n_src + src] : img;
      const uint32_t sK = sKV + s * KV_STAGE_BYTES;
      const uint32_t sV = sK + QCH * K_CHUNK;
      if (elect_one()) {
        mbar_arrive_expect_tx(kv_full + 8 * s, KV_STAGE_BYTES);
        for (int c = 0; c < QCH; ++c)
          tma_load_3d(sK + c * K_CHUNK, &tmK, kv_full + 8 * s, p.k_col0 + head * p.k_hs + c * 64, jt * BN, kv_img);
        for (int c = 0; c < VCH; ++c)
          tma_load_3d(sV + c * K_CHUNK, &tmV, kv_full + 8 * s, p.v_col0 + head * p.v_hs + c * 64, jt * BN, kv_img);
      }
      __syncwarp();
    };
    auto issue_qk = [&](int g) {
      const int s = g % STAGES;
      mbar_wait(kv_full + 8 * s, (g / STAGES) & 1);
      tc_fence_after();
      const uint32_t sK = sKV + s * KV_STAGE_BYTES;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < DQK / 16; ++kk) {
          umma_bf16(tmem_S, umma_smem_desc(sQ + (kk >> 2) * Q_CHUNK + (kk & 3) * 32, 16, 1024, 2),
                    umma_smem_desc(sK + (kk >> 2) * K_CHUNK + (kk & 3) * 32, 16, 1024, 2), IDESC_S, kk != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
      }
      __syncwarp();
    };
#pragma unroll 1
    for (int g = 0; g < STAGES && g < total; ++g) produce(g);
    mbar_wait(q_full, 0);
    issue_qk(0);
#pragma unroll 1
    for (int g = 0; g < total; ++g) {
      const int s = g % STAGES;
      // S(g+1) = Q K(g+1)^T is issued as soon as the softmax threads hold S(g) in registers
      if (STAGES > 1 && g + 1 < total) {
        mbar_wait(s_free, g & 1);
        tc_fence_after();
        issue_qk(g + 1);
      }
      mbar_wait(p_full, g & 1);
      tc_fence_after();
      const uint32_t sV = sKV + s * KV_STAGE_BYTES + QCH * K_CHUNK;
      // O += P V : A = P from TMEM (8 packed columns per 16 keys), B = V (MN-major: rows = keys, 64-wide dv chunks)
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
          umma_bf16_ts(tmem_O, tmem_P + kk * 8, umma_smem_desc(sV + kk * 16 * 128, K_CHUNK, 1024, 2), IDESC_O,
                       (kk != 0 || (g % p.n_kv_tiles) != 0) ? 1u : 0u);  // O accumulates in TMEM over one source
        }
        umma_commit(kv_empty + 8 * s);
        umma_commit(o_full);
      }
      __syncwarp();
      if (g + STAGES < total) produce(g + STAGES);   // stage s is free once P(g) V(g) retires (kv_empty)
      if (STAGES == 1 && g + 1 < total) {            // single stage: K(g+1) only lands after P(g) V(g) has retired
        mbar_wait(s_free, g & 1);
        tc_fence_after();
        issue_qk(g + 1);
      }
    }
  } else {
    // ------------------------------- softmax / correction / epilogue -------------------------------
    // Thread = one query row (TMEM lane): the online softmax needs no shuffles.  Per 128x128 tile and SM it costs ~1024
    // cycles of the 16-lane MUFU pipe and ~730 cycles of the TMEM read port (tcgen05.ld moves ~90 B/clk/SM,
    // profiles/micro/tmem_bench.cu; the S tile is 64 KB) — both shared by the two CTAs of an SM.
    const int row = threadIdx.x;                // == TMEM lane
    const int q_row = q_tile * ATT_BM + row;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    const float sl2 = p.scale_log2e;
    const uint64_t SL2 = pack_f32x2(sl2, sl2);
    bf16* orow = p.out + ((long long)img * p.Lq + q_row) * p.out_ld + head * p.o_hs;
    constexpr int NC = BN / 2;                  // S columns per pass (P is published in two passes)
    int g = 0;
    if constexpr (PIPE) {
      // ------------------------------------------------------------------------------------------------------------
      // Pipelined softmax (opt-in variant, DD_ATTN_PIPE=1; slower than the classic loop on B200 as measured -- kept for A/B).
      // The classic loop below reads the whole S tile, then reduces its maximum, then runs the exponentials: the TMEM
      // read port and the MUFU pipe are used one after the other (MUFU ~60 % busy).  Here the exponent reference m is the
      // running maximum of the PREVIOUS tiles (exact for the first tile via a max-only prepass), so a 32-column chunk can
      // be exponentiated as soon as it is in registers while tcgen05.ld fetches the next chunk.  m is advanced at the end
      // of a tile when the tile maximum grew by more than 2^8 (O is rescaled at the start of the next tile, after P V of
      // this tile retired); a chunk whose maximum exceeds m by more than 2^64 (fp32 / bf16 range guard, practically never)
      // takes an in-place rescale of l, the P chunks already published, and O.
      // ------------------------------------------------------------------------------------------------------------
      constexpr int NCH = BN / 32;
      for (int src = 0; src < p.n_src; ++src) {
        float m = -INFINITY, l = 0.f, alpha_pend = 1.f;
        bool pend = false;
#pragma unroll 1
        for (int jt = 0; jt < p.n_kv_tiles; ++jt, ++g) {
          mbar_wait(s_full, g & 1);
          tc_fence_after();
          const int nvalid = p.Lk - jt * BN;         // keys >= nvalid in this tile are padding (warp-uniform)
          uint32_t buf[2][32];
          auto chunk_max = [&](uint32_t (&t)[32], int c) {
            if (nvalid < BN) {
#pragma unroll
              for (int j = 0; j < 32; ++j) t[j] = (c * 32 + j < nvalid) ? t[j] : 0xff800000u;   // -inf
            }
            float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              a0 = fmax3(a0, __uint_as_float(t[j]), __uint_as_float(t[j + 1]));
              a1 = fmax3(a1, __uint_as_float(t[j + 2]), __uint_as_float(t[j + 3]));
              a2 = fmax3(a2, __uint_as_float(t[j + 4]), __uint_as_float(t[j + 5]));
              a3 = fmax3(a3, __uint_as_float(t[j + 6]), __uint_as_float(t[j + 7]));
            }
            return fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
          };
          if (jt == 0) {   // first tile of a source: exact maximum (max-only prepass over the tile)
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
              tmem_ld_32x32(tmem_S + lane_sel + c * 32, buf[0]);
              tmem_ld_wait();
              mx = fmaxf(mx, chunk_max(buf[0], c));
            }
            m = mx * sl2;
          }
          uint64_t NEGM = pack_f32x2(-m, -m);
          uint64_t acc0 = 0ull, acc1 = 0ull, acc2 = 0ull, acc3 = 0ull;
          float tmax = -INFINITY;
          bool o_ready = false;
          auto ensure_o = [&]() {   // P V of the previous tile retired: P may be overwritten, O may be rescaled
            if (o_ready) return;
            o_ready = true;
            if (g > 0) {
              mbar_wait(o_full, (g - 1) & 1);
              tc_fence_after();
            }
            if (pend) {             // warp-uniform: the reference maximum moved at the end of the previous tile
#pragma unroll
              for (int c = 0; c < DVP; c += 16) {
                uint32_t ov[16];
                tmem_ld_32x16(tmem_O + lane_sel + c, ov);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) ov[j] = __float_as_uint(__uint_as_float(ov[j]) * alpha_pend);
                tmem_st_32x16(tmem_O + lane_sel + c, ov);
              }
              pend = false;
            }
          };
          constexpr int HOLD = NCH >= 4 ? 2 : 1;
          uint32_t pk_hold[HOLD][16], pk_cur[16];
          tmem_ld_32x32(tmem_S + lane_sel, buf[0]);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            uint32_t (&cur)[32] = buf[c & 1];
            if (c + 1 < NCH) tmem_ld_32x32(tmem_S + lane_sel + (c + 1) * 32, buf[(c + 1) & 1]);   // overlaps the exponentials
            const float cm = chunk_max(cur, c);
            tmax = fmaxf(tmax, cm);
            if (__any_sync(0xffffffffu, cm * sl2 > m + 64.f)) {
              // range guard (rare): move the reference now and rescale everything accumulated under the old one
              const float m_new = fmaxf(m, cm * sl2);
              const float a = fast_exp2(m - m_new);
              const uint64_t A2 = pack_f32x2(a, a);
              l *= a;
              acc0 = fma_f32x2(acc0, A2, 0ull); acc1 = fma_f32x2(acc1, A2, 0ull);
              acc2 = fma_f32x2(acc2, A2, 0ull); acc3 = fma_f32x2(acc3, A2, 0ull);
              ensure_o();
              tmem_st_wait();
#pragma unroll
              for (int cp = 0; cp < NCH; ++cp) {
                if (cp < c) {
                  if (c < HOLD) {            // still held in registers
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                      const float2 f = unpack_bf16(pk_hold[cp < HOLD ? cp : 0][j]);
                      pk_hold[cp < HOLD ? cp : 0][j] = pack_bf16(f.x * a, f.y * a);
                    }
                  } else {
                    uint32_t pv[16];
                    tmem_ld_32x16(tmem_P + lane_sel + cp * 16, pv);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                      const float2 f = unpack_bf16(pv[j]);
                      pv[j] = pack_bf16(f.x * a, f.y * a);
                    }
                    tmem_st_32x16(tmem_P + lane_sel + cp * 16, pv);
                  }
                }
              }
              if (jt > 0 || c > 0) {
#pragma unroll
                for (int cc = 0; cc < DVP; cc += 16) {
                  uint32_t ov[16];
                  tmem_ld_32x16(tmem_O + lane_sel + cc, ov);
                  tmem_ld_wait();
#pragma unroll
                  for (int j = 0; j < 16; ++j) ov[j] = __float_as_uint(__uint_as_float(ov[j]) * a);
                  tmem_st_32x16(tmem_O + lane_sel + cc, ov);
                }
              }
              m = m_new;
              NEGM = pack_f32x2(-m, -m);
            }
            // P of the first HOLD chunks stays in registers: the wait for P V of the previous tile (which still reads the
            // P buffer) is pushed as late as possible
            uint32_t (&pk)[16] = (c < HOLD) ? pk_hold[c < HOLD ? c : 0] : pk_cur;
#pragma unroll
            for (int jj = 0; jj < 32; jj += 2) {
              const uint64_t X = fma_f32x2(pack_f32x2(__uint_as_float(cur[jj]), __uint_as_float(cur[jj + 1])), SL2, NEGM);
              float x0, x1;
              unpack_f32x2(X, x0, x1);
              const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
              const uint64_t PP = pack_f32x2(p0, p1);
              const int u = (jj >> 1) & 3;
              if (u == 0) acc0 = add_f32x2(acc0, PP);
              if (u == 1) acc1 = add_f32x2(acc1, PP);
              if (u == 2) acc2 = add_f32x2(acc2, PP);
              if (u == 3) acc3 = add_f32x2(acc3, PP);
              pk[jj >> 1] = pack_bf16(p0, p1);
            }
            if (c == HOLD - 1) {
              ensure_o();
#pragma unroll
              for (int cp = 0; cp < HOLD; ++cp) tmem_st_32x16(tmem_P + lane_sel + cp * 16, pk_hold[cp]);
            } else if (c >= HOLD) {
              tmem_st_32x16(tmem_P + lane_sel + c * 16, pk_cur);
            }
            if (c + 1 < NCH) {
              tmem_ld_wait();
              if (c + 2 == NCH) {      // the last chunk of S is in registers -> the issuer may start S(g+1) = Q K(g+1)^T
                tc_fence_before();
                mbar_arrive(s_free);
              }
            }
          }
          if (NCH == 1) {
            tc_fence_before();
            mbar_arrive(s_free);
          }
          {
            float s0, s1;
            unpack_f32x2(add_f32x2(add_f32x2(acc0, acc1), add_f32x2(acc2, acc3)), s0, s1);
            l += s0 + s1;
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(p_full);
          // lazy reference update for the following tiles
          if (jt + 1 < p.n_kv_tiles) {
            const float m_cand = fmaxf(m, tmax * sl2);
            if (__any_sync(0xffffffffu, m_cand > m + 8.f)) {
              alpha_pend = fast_exp2(m - m_cand);
              l *= alpha_pend;
              m = m_cand;
              pend = true;
            }
          }
        }
        // epilogue of this source: O / l  (second source of the cross-view attention adds onto the first)
        mbar_wait(o_full, (g - 1) & 1);
        tc_fence_after();
        const float inv = 1.f / l;
#pragma unroll
        for (int c = 0; c < DVP; c += 16) {
          uint32_t ov[16];
          tmem_ld_32x16(tmem_O + lane_sel + c, ov);
          tmem_ld_wait();
          if (q_row < p.Lq) {
#pragma unroll
            for (int hh = 0; hh < 16; hh += 8) {
              if (c + hh + 8 <= DV) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(ov[hh + e]) * inv;
                if (src > 0) {
                  const uint4 r = *reinterpret_cast<const uint4*>(orow + c + hh);
                  float2 t;
                  t = unpack_bf16(r.x); v[0] += t.x; v[1] += t.y;
                  t = unpack_bf16(r.y); v[2] += t.x; v[3] += t.y;
                  t = unpack_bf16(r.z); v[4] += t.x; v[5] += t.y;
                  t = unpack_bf16(r.w); v[6] += t.x; v[7] += t.y;
                }
                *reinterpret_cast<uint4*>(orow + c + hh) =
                    make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
              }
            }
          }
        }
        tc_fence_before();
      }
    } else
    for (int src = 0; src < p.n_src; ++src) {
      float m = -INFINITY, l = 0.f;
#pragma unroll 1
      for (int jt = 0; jt < p.n_kv_tiles; ++jt, ++g) {
        mbar_wait(s_full, g & 1);
        tc_fence_after();
        uint32_t sv[BN];
#pragma unroll
        for (int c = 0; c < BN; c += 32) tmem_ld_32x32(tmem_S + lane_sel + c, *reinterpret_cast<uint32_t(*)[32]>(&sv[c]));
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_free);                       // S(g) now lives in registers -> the issuer starts S(g+1)
        const int nvalid = p.Lk - jt * BN;         // keys >= nvalid in this tile are padding (warp-uniform)
        if (nvalid < BN) {
#pragma unroll
          for (int j = 0; j < BN; ++j) sv[j] = (j < nvalid) ? sv[j] : 0xff800000u;   // -inf
        }
        // row maximum on four independent FMNMX3 chains
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int j = 0; j < BN; j += 8) {
          mx0 = fmax3(mx0, __uint_as_float(sv[j]), __uint_as_float(sv[j + 1]));
          mx1 = fmax3(mx1, __uint_as_float(sv[j + 2]), __uint_as_float(sv[j + 3]));
          mx2 = fmax3(mx2, __uint_as_float(sv[j + 4]), __uint_as_float(sv[j + 5]));
          mx3 = fmax3(mx3, __uint_as_float(sv[j + 6]), __uint_as_float(sv[j + 7]));
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        // lazy running max: only advanced (and O rescaled) when it grows by more than 2^8, so p <= 256
        const float m_cand = fmaxf(m, mx * sl2);
        const bool grow = __any_sync(0xffffffffu, m_cand > m + 8.f);   // warp-uniform (first tile: m = -inf)
        float alpha = 1.f;
        if (grow) {
          alpha = fast_exp2(m - m_cand);
          m = m_cand;
          l *= alpha;
        }
        // p = exp2(s * scale * log2e - m) packed to bf16; row sum on four packed accumulators.  P is published in two
        // passes of BN/2 columns so that at most BN/4 packed registers are live next to the unconsumed half of S.
        const uint64_t NEGM = pack_f32x2(-m, -m);
        uint64_t acc0 = 0ull, acc1 = 0ull, acc2 = 0ull, acc3 = 0ull;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          uint32_t pk[NC / 2];
#pragma unroll
          for (int jj = pass * NC; jj < (pass + 1) * NC; jj += 2) {
            const uint64_t X = fma_f32x2(pack_f32x2(__uint_as_float(sv[jj]), __uint_as_float(sv[jj + 1])), SL2, NEGM);
            float x0, x1, p0, p1;
            unpack_f32x2(X, x0, x1);
            if (((jj >> 1) & 7) < POLY) {
              exp2_poly_pair(x0, x1, p0, p1);
            } else {
              p0 = fast_exp2(x0);
              p1 = fast_exp2(x1);
            }
            const uint64_t PP = pack_f32x2(p0, p1);
            const int u = (jj >> 1) & 3;
            if (u == 0) acc0 = add_f32x2(acc0, PP);
            if (u == 1) acc1 = add_f32x2(acc1, PP);
            if (u == 2) acc2 = add_f32x2(acc2, PP);
            if (u == 3) acc3 = add_f32x2(acc3, PP);
            pk[(jj >> 1) - pass * (NC / 2)] = pack_bf16(p0, p1);
          }
          if (pass == 0) {
            // the previous tile's P V must have retired before P is overwritten / O is rescaled
            if (g > 0) {
              mbar_wait(o_full, (g - 1) & 1);
              tc_fence_after();
            }
            if (grow && jt > 0) {
#pragma unroll
              for (int c = 0; c < DVP; c += 16) {
                uint32_t ov[16];
                tmem_ld_32x16(tmem_O + lane_sel + c, ov);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) ov[j] = __float_as_uint(__uint_as_float(ov[j]) * alpha);
                tmem_st_32x16(tmem_O + lane_sel + c, ov);
              }
            }
          }
          const uint32_t pdst = tmem_P + lane_sel + (uint32_t)(pass * (NC / 2));
          if constexpr (NC / 2 == 32) {
            tmem_st_32x32(pdst, pk);
          } else {
            tmem_st_32x16(pdst, *reinterpret_cast<const uint32_t(*)[16]>(&pk[0]));
          }
        }
        {
          float s0, s1;
          unpack_f32x2(add_f32x2(add_f32x2(acc0, acc1), add_f32x2(acc2, acc3)), s0, s1);
          l += s0 + s1;
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full);
      }
      // epilogue of this source: O / l  (second source of the cross-view attention adds onto the first)
      mbar_wait(o_full, (g - 1) & 1);
      tc_fence_after();
      const float inv = 1.f / l;
#pragma unroll
      for (int c = 0; c < DVP; c += 16) {
        uint32_t ov[16];
        tmem_ld_32x16(tmem_O + lane_sel + c, ov);
        tmem_ld_wait();
        if (q_row < p.Lq) {
#pragma unroll
          for (int hh = 0; hh < 16; hh += 8) {
            if (c + hh + 8 <= DV) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(ov[hh + e]) * inv;
              if (src > 0) {
                const uint4 r = *reinterpret_cast<const uint4*>(orow + c + hh);
                float2 t;
                t = unpack_bf16(r.x); v[0] += t.x; v[1] += t.y;
                t = unpack_bf16(r.y); v[2] += t.x; v[3] += t.y;
                t = unpack_bf16(r.z); v[4] += t.x; v[5] += t.y;
                t = unpack_bf16(r.w); v[6] += t.x; v[7] += t.y;
              }
              *reinterpret_cast<uint4*>(orow + c + hh) =
                  make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            }
          }
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == ISSUER) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int DQK, int DV, int DVP, int BN, int STAGES, int MINB, int POLY, bool PIPE>
static int launch_attn_v2(const dd_attention_args* a, AttnDev p, cudaStream_t stream) {
  constexpr int QCH = (DQK + 63) / 64, VCH = (DVP + 63) / 64;
  constexpr size_t smem = (size_t)QCH * ATT_BM * 128 + (size_t)STAGES * (QCH + VCH) * BN * 128 + 128;
  static_assert(STAGES <= 4, "barrier block holds 4 stages");
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  rc = make_tmap_3d_bf16(&tmQ, a->q, (uint64_t)a->q_cols, (uint64_t)a->lq, (uint64_t)a->n_img, (uint64_t)a->q_ld,
                         (uint64_t)a->lq * a->q_ld, 64, ATT_BM, 1);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tmK, a->k, (uint64_t)a->k_cols, (uint64_t)a->lk, (uint64_t)a->n_kv_img, (uint64_t)a->k_ld,
                         (uint64_t)a->lk * a->k_ld, 64, BN, 1);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tmV, a->v, (uint64_t)a->v_cols, (uint64_t)a->lk, (uint64_t)a->n_kv_img, (uint64_t)a->v_ld,
                         (uint64_t)a->lk * a->v_ld, 64, BN, 1);
  if (rc) return rc;
  p.n_kv_tiles = (a->lk + BN - 1) / BN;
  static bool attr_done = false;
  if (!attr_done) {
    DD_CUDA(cudaFuncSetAttribute(attn_v2_kernel<DQK, DV, DVP, BN, STAGES, MINB, POLY, PIPE>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid((a->lq + ATT_BM - 1) / ATT_BM, a->heads, a->n_img);
  attn_v2_kernel<DQK, DV, DVP, BN, STAGES, MINB, POLY, PIPE><<<grid, ATT_THREADS, smem, stream>>>(tmQ, tmK, tmV, p);
  DD_CUDA(cudaGetLastError());
  return 0;
}

template <int DQK, int DV, int DVP, int STAGES, int MINB>
static int launch_attn(const dd_attention_args* a, const CUtensorMap& tmQ, const CUtensorMap& tmK,
                       const CUtensorMap& tmV, AttnDev p, cudaStream_t stream) {
  constexpr int QCH = (DQK + 63) / 64, VCH = (DVP + 63) / 64;
  constexpr size_t smem = (size_t)QCH * CHUNK_BYTES + (size_t)STAGES * (QCH + VCH) * CHUNK_BYTES + 2 * CHUNK_BYTES + 128;
  static bool attr_done = false;
  if (!attr_done) {
    DD_CUDA(cudaFuncSetAttribute(attn_tcgen05_kernel<DQK, DV, DVP, STAGES, MINB>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid((a->lq + ATT_BM - 1) / ATT_BM, a->heads, a->n_img);
  attn_tcgen05_kernel<DQK, DV, DVP, STAGES, MINB><<<grid, ATT_THREADS, smem, stream>>>(tmQ, tmK, tmV, p);
  DD_CUDA(cudaGetLastError());
  return 0;
}

int attention_run(const dd_attention_args* a, cudaStream_t stream) {
  DD_CHECK(a != nullptr, -1, "dd_attention: null args");
  DD_CHECK(a->n_img > 0 && a->heads > 0 && a->lq > 0 && a->lk > 0, -1, "dd_attention: bad shape");
  DD_CHECK(a->head_dim == 40 || a->head_dim == 80 || a->head_dim == 160, -1,
           "dd_attention: head_dim %d unsupported (40, 80, 160 = SDv1.5 levels)", a->head_dim);
  DD_CHECK(a->n_src == 1 || a->n_src == 2, -1, "dd_attention: n_src must be 1 or 2");
  DD_CHECK(a->n_src == 1 || a->kv_map != nullptr, -1, "dd_attention: kv_map required for n_src == 2");
  DD_CHECK(a->q_ld % 8 == 0 && a->k_ld % 8 == 0 && a->v_ld % 8 == 0 && a->out_ld % 8 == 0, -1,
           "dd_attention: leading dims must be multiples of 8");
  DD_CHECK(a->n_kv_img > 0, -1, "dd_attention: n_kv_img missing");
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  rc = make_tmap_3d_bf16(&tmQ, a->q, (uint64_t)a->q_cols, (uint64_t)a->lq, (uint64_t)a->n_img, (uint64_t)a->q_ld,
                         (uint64_t)a->lq * a->q_ld, 64, ATT_BM, 1);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tmK, a->k, (uint64_t)a->k_cols, (uint64_t)a->lk, (uint64_t)a->n_kv_img, (uint64_t)a->k_ld,
                         (uint64_t)a->lk * a->k_ld, 64, ATT_BN, 1);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tmV, a->v, (uint64_t)a->v_cols, (uint64_t)a->lk, (uint64_t)a->n_kv_img, (uint64_t)a->v_ld,
                         (uint64_t)a->lk * a->v_ld, 64, ATT_BN, 1);
  if (rc) return rc;
  AttnDev p;
  p.Lq = a->lq; p.Lk = a->lk; p.n_src = a->n_src; p.n_kv_tiles = (a->lk + ATT_BN - 1) / ATT_BN;
  p.kv_map = a->kv_map;
  p.scale_log2e = a->scale * 1.4426950408889634f;
  p.out = reinterpret_cast<bf16*>(a->out); p.out_ld = a->out_ld;
  p.q_col0 = a->q_col0; p.k_col0 = a->k_col0; p.v_col0 = a->v_col0;
  p.q_hs = a->q_head_stride; p.k_hs = a->k_head_stride; p.v_hs = a->v_head_stride; p.o_hs = a->head_dim;
  // DD_ATTN_IMPL=1 selects the v1 kernel (P through shared memory) for A/B measurements; DD_ATTN_POLY = exponentials
  // per 8 evaluated on the FMA pipe at head_dim 40 (0, 2, 3 or 4; default 3)
  static const int impl = getenv("DD_ATTN_IMPL") ? atoi(getenv("DD_ATTN_IMPL")) : 3;
  static const int poly = getenv("DD_ATTN_POLY") ? atoi(getenv("DD_ATTN_POLY")) : 0;
  // DD_ATTN_PIPE=1 selects the chunk-pipelined softmax (exponent reference = running maximum of the previous tiles, 32-column
  // chunks exponentiated while the next chunk is read).  Measured on B200 (profiles/attn_one.py, 96 images, d = 40,
  // L = 1400): 665 us against 585 us for the read-all / max / exponentiate loop, so it stays an opt-in A/B variant.
  static const int pipe = getenv("DD_ATTN_PIPE") ? atoi(getenv("DD_ATTN_PIPE")) : 0;
  if (impl != 1) {
    switch (a->head_dim) {
      case 40: {
        DD_CHECK(a->q_head_stride >= 48 && a->k_head_stride >= 48, -1,
                 "dd_attention: head_dim 40 needs Q/K heads zero-padded to a 48-column stride");
        static const int bn64 = getenv("DD_ATTN_BN64") ? atoi(getenv("DD_ATTN_BN64")) : 0;   // A/B: 64-key tiles at d = 40
        if (bn64) return launch_attn_v2<48, 40, 48, 64, 3, 2, 0, false>(a, p, stream);
        if (poly == 2) return launch_attn_v2<48, 40, 48, 128, 3, 2, 2, false>(a, p, stream);
        if (pipe) return launch_attn_v2<48, 40, 48, 128, 3, 2, 0, true>(a, p, stream);
        return launch_attn_v2<48, 40, 48, 128, 3, 2, 0, false>(a, p, stream);
      }
      case 80:
        if (pipe) return launch_attn_v2<80, 80, 80, 64, 2, 2, 0, true>(a, p, stream);
        return launch_attn_v2<80, 80, 80, 64, 2, 2, 0, false>(a, p, stream);
      case 160:
        if (pipe) return launch_attn_v2<160, 160, 160, 64, 1, 2, 0, true>(a, p, stream);
        return launch_attn_v2<160, 160, 160, 64, 1, 2, 0, false>(a, p, stream);
    }
  }
  switch (a->head_dim) {
    case 40:
      DD_CHECK(a->q_head_stride >= 48 && a->k_head_stride >= 48, -1,
               "dd_attention: head_dim 40 needs Q/K heads zero-padded to a 48-column stride");
      return launch_attn<48, 40, 48, 2, 2>(a, tmQ, tmK, tmV, p, stream);
    case 80: return launch_attn<80, 80, 80, 2, 1>(a, tmQ, tmK, tmV, p, stream);
    case 160: return launch_attn<160, 160, 160, 1, 1>(a, tmQ, tmK, tmV, p, stream);
  }
  return -1;
}

}  // namespace dd
