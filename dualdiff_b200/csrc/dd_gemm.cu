// dualdiff_b200 — tcgen05 / TMEM / TMA GEMM core (sm_100a).
//
// One persistent, warp-specialised kernel computes
//     out[m, n] = epilogue( sum_{tap, k} A[m + shift(tap), k] * W[n, tap*K + k] )
// which covers, on the DualDiff denoising path:
//   * Linear / 1x1 conv   (taps = 1)                         diffusers Attention.to_q/k/v/out,
//                                                            FeedForward, Transformer2DModel.proj_in/out,
//                                                            ResnetBlock2D.conv_shortcut, ControlNet zero convs
//   * 3x3 stride-1 conv as implicit GEMM (taps = 9)          ResnetBlock2D.conv1/conv2, Upsample2D.conv,
//                                                            conv_out.  The activation is stored in a
//                                                            zero-haloed "padded pixel" layout
//                                                            [img][H+1][W+1][C], so tap (kh,kw) is the same
//                                                            rows shifted by (kh-1)*(W+1)+(kw-1); one staged
//                                                            box of 128+2 rows serves the three kw taps of a
//                                                            kernel row (row-shifted UMMA descriptors).
//   * channel concat of two sources along K (skip connections of the up blocks) without torch.cat.
// Epilogues (fused): +bias[n], +per-image vector (time-embedding), +residual(s), GEGLU, fp32/bf16 out,
// padded-pixel -> compact-row scatter for the conv mode.
//
// Roles: warp0 = TMA producer, warp1 = UMMA issuer (+TMEM alloc), warps2-9 = epilogue (TMEM -> regs -> HBM; a TMA
// load/store epilogue for plain bf16 GEMMs, a register epilogue for conv scatter / fp32 / stream-K partials).
// Pipelines: smem rings (full/empty mbarriers; conv mode keeps separate activation and weight rings) and a 2-deep TMEM
// accumulator ring (tmem_full/tmem_empty).  Work: one tile per CTA pair, or stream-K ranges of the flattened
// (tile, k-iteration) space when the tiles fill the last wave badly (second pass: splitk_reduce_kernel).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "dd_api_internal.h"
#include "dd_common.cuh"

namespace dd {

// A/B hooks of profiles/ab (the shipped library is built without -D overrides)
#ifndef DD_GEMM_AREUSE
#define DD_GEMM_AREUSE 1
#endif
#ifndef DD_CONV_MERGED
#define DD_CONV_MERGED 1
#endif

// Timeline instrumentation for profiles/gemm_trace.py: only in a -DDD_GEMM_TRACE build (never in the shipped library).  Lane 0
// of every warp of ONE CTA records (event, warp, tile, SM clock).
#ifdef DD_GEMM_TRACE
__device__ unsigned long long g_gemm_trace[10 * 2 * 2048];
#define DD_GT_DECL unsigned int gt_n_ = 0;
#define DD_GT(ev, tile)                                                                                       \
  do {                                                                                                        \
    if (blockIdx.x == 6 && (threadIdx.x & 31) == 0 && gt_n_ < 2048u) {                                        \
      unsigned long long* e_ = g_gemm_trace + ((threadIdx.x >> 5) * 2048u + gt_n_) * 2u;                      \
      e_[0] = ((unsigned long long)(ev) << 48) | ((unsigned long long)(threadIdx.x >> 5) << 32) | (unsigned)(tile); \
      e_[1] = clock64();                                                                                      \
      ++gt_n_;                                                                                                \
    }                                                                                                         \
  } while (0)
// accumulate the cycles of one statement (barrier waits of the issuing warp) / record an accumulated value as an event
#define DD_GT_TIMED(acc, stmt) do { const long long t_ = clock64(); stmt; acc += clock64() - t_; } while (0)
#define DD_GT_VAL(ev, tile, val)                                                                              \
  do {                                                                                                        \
    if (blockIdx.x == 6 && (threadIdx.x & 31) == 0 && gt_n_ < 2048u) {                                        \
      unsigned long long* e_ = g_gemm_trace + ((threadIdx.x >> 5) * 2048u + gt_n_) * 2u;                      \
      e_[0] = ((unsigned long long)(ev) << 48) | ((unsigned long long)(threadIdx.x >> 5) << 32) | (unsigned)(tile); \
      e_[1] = (unsigned long long)(val) | (1ull << 62);                                                       \
      ++gt_n_;                                                                                                \
    }                                                                                                         \
  } while (0)
#else
#define DD_GT_DECL
#define DD_GT(ev, tile) do { } while (0)
#define DD_GT_TIMED(acc, stmt) stmt
#define DD_GT_VAL(ev, tile, val) do { } while (0)
#endif

static constexpr int BM = 128;
static constexpr int BK = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
static constexpr int GEMM_THREADS_MAX = 320;  // TMA warp + UMMA warp + 8 epilogue warps
static constexpr int A_STAGE_BYTES = BM * BK * 2;
static constexpr int EPI_WARP_BYTES = 32 * 64 * 4;  // per-epilogue-warp transpose buffer (32 rows x 64 fp32)
// 3x3 conv mode: one staged activation box of 128 + 2 halo rows serves the three kw taps of a kernel row (the UMMA
// descriptor of tap kw starts kw rows = kw * 128 B into the box: SWIZZLE_128B phases follow the absolute shared-memory
// address, profiles/micro/desc_shift_test.cu), so the activation crosses the L2 -> SM fabric 3x instead of 9x.
static constexpr int A_CONV_ROWS = BM + 2;
static constexpr int A_CONV_BYTES = 17 * 1024;   // 130 rows x 128 B = 16640, slots kept 1024-byte aligned
static constexpr int A_CONV_TX = A_CONV_ROWS * BK * 2;

struct GemmDev {
  int M, N, K, K1, taps;
  int conv_H, conv_W;  // conv mode (taps == 9): image geometry of the padded layout
  int m_tiles, n_tiles, stages;
  int a_stages;   // conv mode: slots of the activation ring (stages = slots of the weight ring)
  // epilogue
  void* out;
  long long out_ld;
  int out_f32;
  const float* bias;
  const float* rowvec;
  int rowvec_ld;
  int rows_per_img;
  const bf16* res1;
  long long res1_ld;
  const bf16* res2;
  long long res2_ld;
  int geglu;
  int act;      // 0 none, 1 SiLU (applied last)
  int tma_epi;  // 1: epilogue moves residual/output tiles with TMA (plain GEMM, bf16 out)
  int n_store;  // number of valid output columns (N, or N/2 for GEGLU)
  // stream-K (sk_chunk > 0): the flattened (tile, k-iteration) space is cut into equal contiguous ranges, one per worker;
  // every segment's raw fp32 accumulator goes to ws[tile * sk_maxc + contributor][TILE_M][BN] and splitk_reduce_kernel
  // sums the contributors in fixed order and applies the epilogue (bit-reproducible, no atomics).
  int sk_chunk, sk_maxc;
  float* ws;
  // areuse = 1 (short K: the ring has exactly one slot per 64-wide k-block, so slot s always holds k-block s): every worker
  // takes a CONTIGUOUS range of the (m-major, n-fastest) tile order and keeps the A blocks of an m-tile in the ring for all
  // of its n-tiles -- only the B half of a stage is reloaded.  The activation then crosses the L2 -> SM fabric once instead
  // of n_tiles times (ncu: the K = 320 GEMMs move 5500 B/clk through a fabric that caps near 6300).
  int areuse;
};

// Work iterator shared by the three roles of the GEMM kernel.  Classic: tiles worker, worker + n_workers, ... with all
// k-iterations each.  Stream-K: the worker's contiguous range of the flattened (tile, k-iteration) space, cut at tile
// boundaries into segments.
struct WorkIter {
  int tile, it0, it1;
  int pos, end, iters, step, total_tiles, sk;
  // contiguous = 1: the worker owns tiles [worker * T / W, (worker + 1) * T / W) (GemmDev::areuse)
  __device__ __forceinline__ WorkIter(int worker, int n_workers, int total_tiles_, int iters_, int sk_chunk, int contiguous = 0)
      : tile(worker - n_workers), it0(0), it1(iters_), pos(worker * sk_chunk), iters(iters_), step(n_workers),
        total_tiles(total_tiles_), sk(sk_chunk) {
    const int total = total_tiles_ * iters_;
    end = pos + sk_chunk < total ? pos + sk_chunk : total;
    if (contiguous) {
      tile = (int)(((long long)worker * total_tiles_) / n_workers) - 1;
      total_tiles = (int)(((long long)(worker + 1) * total_tiles_) / n_workers);
      step = 1;
    }
  }
  __device__ __forceinline__ bool next() {
    if (sk == 0) {
      tile += step;
      return tile < total_tiles;
    }
    if (pos >= end) return false;
    tile = pos / iters;
    it0 = pos - tile * iters;
    it1 = it0 + (end - pos) < iters ? it0 + (end - pos) : iters;
    pos += it1 - it0;
    return true;
  }
};

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x BN tile — each
// CTA stages its own 128 A rows and HALF of the B tile, the leader's single thread issues M=256 UMMAs that read both
// CTAs' shared memory and write both CTAs' TMEM.  Per SM that halves the B bytes written (TMA) and read (UMMA) per
// MMA cycle: 128 + 0.5*(8192/BN+...) -- the 128 B/clk shared-memory port is what caps the 1-CTA kernel at ~57 %
// tensor-pipe activity (ncu, profiles/r01_ncu_conv_*.txt).
template <int BN, int CG, bool CONV>
__global__ void __launch_bounds__(GEMM_THREADS_MAX, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                    const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                    const __grid_constant__ CUtensorMap tmR1, const GemmDev p) {
  constexpr int B_STAGE_BYTES = (BN / CG) * BK * 2;   // per CTA: the whole B tile, or its half of the pair's tile
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr int TILE_M = BM * CG;
  // 3x3 conv on tiles up to 192 wide: ONE ring whose slot holds everything an iteration (kernel row kh, 64 channels) needs --
  // the activation box and the weight boxes of its three taps -- behind one full / one empty barrier.  The timeline of the
  // issuing warp (profiles/gemm_trace.py conv) showed every barrier wait costing ~90 cycles even when the data is there
  // and every UMMA / commit ~50: with a wait + commit per tap (4 UMMAs) a stage took 383 cycles against the 320 tensor
  // cycles of its four 160-wide UMMAs -- the issuing thread, not the tensor core, set the pace.  One wait + one commit per
  // twelve UMMAs is ~740 cycles of issue against 960 of tensor work.  (256-wide tiles: a UMMA is 128 tensor cycles, the
  // issuer keeps up, and only two such slots would fit: they keep the separate activation / weight rings.)
  constexpr bool MERGED = CONV && BN <= 192 && DD_CONV_MERGED;
  constexpr int CONV_SLOT_BYTES = A_CONV_BYTES + 3 * B_STAGE_BYTES;
  constexpr int ACC_COLS = (BN < 32) ? 32 : BN;      // columns per accumulator buffer
  constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128 : (2 * ACC_COLS <= 256) ? 256 : 512;
  constexpr uint32_t IDESC = umma_idesc_bf16(BM * CG, BN, 0, 0);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[2 * 8 + 4 + 16];
  __shared__ uint32_t tmem_ptr_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  DD_GT_DECL
  const int stages = p.stages;
  const uint32_t crank = (CG == 2) ? cluster_ctarank() : 0u;          // 0 = leader of the CTA pair
  const int worker = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_workers = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const uint32_t full_bar = smem_u32(&bars[0]);
  const uint32_t empty_bar = smem_u32(&bars[8]);
  const uint32_t tfull_bar = smem_u32(&bars[16]);
  const uint32_t tempty_bar = smem_u32(&bars[18]);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    if (p.tma_epi) {
      tma_prefetch_desc(&tmOut);
      tma_prefetch_desc(&tmR1);
    }
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar + 8 * s, CG);   // CG = 2: leader's arrive.expect_tx + the peer's remote arrive
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + 8 * s, 1);
      mbar_init(tempty_bar + 8 * s, CG * ((blockDim.x - 64) / 32));  // one arrive per epilogue warp (of both CTAs)
    }
    if constexpr (CONV) {
      for (int s = 0; s < 4; ++s) {
        mbar_init(smem_u32(&bars[20 + s]), CG);  // activation ring: full
        mbar_init(smem_u32(&bars[24 + s]), 1);   //                  empty
      }
    } else {
      for (int s = 0; s < 16; ++s) mbar_init(smem_u32(&bars[20 + s]), 1);  // per epilogue warp: residual landed in staging tile 0 / 1
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) {
      tmem_alloc_2cta(smem_u32(&tmem_ptr_smem), TMEM_COLS);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc(smem_u32(&tmem_ptr_smem), TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr_smem;
  // programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the
  // tail of the previous kernel in the stream; its results are only visible after this wait (no-op without the attribute)
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const int kchunks = (p.K + BK - 1) / BK;
  // conv mode: an iteration = one (kernel row kh, 64-channel chunk): 1 activation box + 3 weight boxes, 12 UMMAs
  const int iters = CONV ? 3 * kchunks : p.taps * kchunks;
  const uint32_t afull_bar = smem_u32(&bars[20]);
  const uint32_t aempty_bar = smem_u32(&bars[24]);
  const uint32_t ring_b = CONV ? smem_base + (uint32_t)p.a_stages * A_CONV_BYTES : smem_base;   // weight ring (conv)
  const uint32_t epi_base = MERGED ? smem_base + (uint32_t)stages * CONV_SLOT_BYTES
                            : CONV ? ring_b + (uint32_t)stages * B_STAGE_BYTES : smem_base + (uint32_t)stages * STAGE_BYTES;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int pitch = p.conv_W + 1;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    // The whole warp runs the loop (warp-uniform control flow lets ptxas keep barrier addresses, coordinates and
    // descriptors in uniform registers); one elected lane issues.  Inside `if (lane == 0)` every operand of
    // UTMALDG / UTCHMMA had to be moved vector->uniform (R2UR) per instruction, which made the kernel issue-bound.
    {
      // ring positions are carried as (slot, parity) counters: no integer division in the steady-state loops
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      WorkIter wi(worker, n_workers, total_tiles, iters, p.sk_chunk, p.areuse);
      int prev_m0 = -1;
      while (wi.next()) {
        const int tile = wi.tile;
        const int m0 = (tile / p.n_tiles) * TILE_M + (int)crank * BM;
        const int n0 = (tile % p.n_tiles) * BN;
        if constexpr (MERGED) {
          int kh = wi.it0 / kchunks;
          int kc = (wi.it0 - kh * kchunks) * BK;
          for (int it = wi.it0; it < wi.it1; ++it) {
            mbar_wait(empty_bar + 8 * sb, pb ^ 1);
            const int arow = m0 + (kh - 1) * pitch - 1;   // box rows [arow, arow + 130): taps kw = 0, 1, 2 start at row kw
            const uint32_t sA = smem_base + sb * CONV_SLOT_BYTES;
            const uint32_t sB = sA + A_CONV_BYTES;
            const int bcol = kh * 3 * p.K + kc;           // weight columns of tap (kh, kw): (kh * 3 + kw) * K + kc
            if constexpr (CG == 1) {
              if (elect_one()) {
                mbar_arrive_expect_tx(full_bar + 8 * sb, A_CONV_TX + 3 * B_STAGE_BYTES);
                if (kc < p.K1) tma_load_2d(sA, &tmA, full_bar + 8 * sb, kc, arow);
                else tma_load_2d(sA, &tmA2, full_bar + 8 * sb, kc - p.K1, arow);
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) tma_load_2d(sB + kw * B_STAGE_BYTES, &tmB, full_bar + 8 * sb, bcol + kw * p.K, n0);
              }
            } else {
              const uint32_t full_leader = mapa_shared(full_bar + 8 * sb, 0);
              if (elect_one()) {
                if (kc < p.K1) tma_load_2d_2cta(sA, &tmA, full_leader, kc, arow);
                else tma_load_2d_2cta(sA, &tmA2, full_leader, kc - p.K1, arow);
#pragma unroll
                for (int kw = 0; kw < 3; ++kw)
                  tma_load_2d_2cta(sB + kw * B_STAGE_BYTES, &tmB, full_leader, bcol + kw * p.K, n0 + (int)crank * (BN / 2));
                if (crank == 0) mbar_arrive_expect_tx(full_bar + 8 * sb, 2 * (A_CONV_TX + 3 * B_STAGE_BYTES));
                else mbar_arrive_cluster(full_leader);
              }
            }
            __syncwarp();
            if (++sb == (uint32_t)stages) { sb = 0; pb ^= 1; }
            kc += BK;
            if (kc >= p.K) { kc = 0; ++kh; }
          }
        } else if constexpr (CONV) {
          int kh = wi.it0 / kchunks;
          int kc = (wi.it0 - kh * kchunks) * BK;
          for (int it = wi.it0; it < wi.it1; ++it) {
            mbar_wait(aempty_bar + 8 * sa, pa ^ 1);
            const int arow = m0 + (kh - 1) * pitch - 1;   // box rows [arow, arow + 130): taps kw = 0, 1, 2 start at row kw
            const uint32_t sA = smem_base + sa * A_CONV_BYTES;
            if constexpr (CG == 1) {
              if (elect_one()) {
                mbar_arrive_expect_tx(afull_bar + 8 * sa, A_CONV_TX);
                if (kc < p.K1) tma_load_2d(sA, &tmA, afull_bar + 8 * sa, kc, arow);
                else tma_load_2d(sA, &tmA2, afull_bar + 8 * sa, kc - p.K1, arow);
              }
            } else {
              const uint32_t afull_leader = mapa_shared(afull_bar + 8 * sa, 0);
              if (elect_one()) {
                if (kc < p.K1) tma_load_2d_2cta(sA, &tmA, afull_leader, kc, arow);
                else tma_load_2d_2cta(sA, &tmA2, afull_leader, kc - p.K1, arow);
                if (crank == 0) mbar_arrive_expect_tx(afull_bar + 8 * sa, 2 * A_CONV_TX);
                else mbar_arrive_cluster(afull_leader);
              }
            }
            __syncwarp();
            if (++sa == (uint32_t)p.a_stages) { sa = 0; pa ^= 1; }
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              mbar_wait(empty_bar + 8 * sb, pb ^ 1);
              const uint32_t sB = ring_b + sb * B_STAGE_BYTES;
              const int bcol = (kh * 3 + kw) * p.K + kc;
              if constexpr (CG == 1) {
                if (elect_one()) {
                  mbar_arrive_expect_tx(full_bar + 8 * sb, B_STAGE_BYTES);
                  tma_load_2d(sB, &tmB, full_bar + 8 * sb, bcol, n0);
                }
              } else {
                const uint32_t full_leader = mapa_shared(full_bar + 8 * sb, 0);
                if (elect_one()) {
                  tma_load_2d_2cta(sB, &tmB, full_leader, bcol, n0 + (int)crank * (BN / 2));
                  if (crank == 0) mbar_arrive_expect_tx(full_bar + 8 * sb, 2 * B_STAGE_BYTES);
                  else mbar_arrive_cluster(full_leader);
                }
              }
              __syncwarp();
              if (++sb == (uint32_t)stages) { sb = 0; pb ^= 1; }
            }
            kc += BK;
            if (kc >= p.K) { kc = 0; ++kh; }
          }
        } else {
          int tap = wi.it0 / kchunks;
          int kc = (wi.it0 - tap * kchunks) * BK;
          // A reuse: the ring slots still hold the A blocks of this m-tile (loaded for the previous n-tile of this worker)
          const bool load_a = !(p.areuse && m0 == prev_m0);
          prev_m0 = m0;
          const uint32_t tx = load_a ? (uint32_t)STAGE_BYTES : (uint32_t)B_STAGE_BYTES;
          DD_GT(20, tile);
          for (int it = wi.it0; it < wi.it1; ++it) {
            mbar_wait(empty_bar + 8 * sb, pb ^ 1);
            if (it == wi.it0) DD_GT(21, tile);
            const uint32_t sA = smem_base + sb * STAGE_BYTES;
            const uint32_t sB = sA + A_STAGE_BYTES;
            if constexpr (CG == 1) {
              if (elect_one()) {
                mbar_arrive_expect_tx(full_bar + 8 * sb, tx);
                if (load_a) {
                  if (kc < p.K1)
                    tma_load_2d(sA, &tmA, full_bar + 8 * sb, kc, m0);
                  else
                    tma_load_2d(sA, &tmA2, full_bar + 8 * sb, kc - p.K1, m0);
                }
                tma_load_2d(sB, &tmB, full_bar + 8 * sb, tap * p.K + kc, n0);
              }
            } else {
              // both CTAs load into their own smem; every byte is credited to the LEADER's full barrier
              const uint32_t full_leader = mapa_shared(full_bar + 8 * sb, 0);
              if (elect_one()) {
                if (load_a) {
                  if (kc < p.K1)
                    tma_load_2d_2cta(sA, &tmA, full_leader, kc, m0);
                  else
                    tma_load_2d_2cta(sA, &tmA2, full_leader, kc - p.K1, m0);
                }
                tma_load_2d_2cta(sB, &tmB, full_leader, tap * p.K + kc, n0 + (int)crank * (BN / 2));
                if (crank == 0)
                  mbar_arrive_expect_tx(full_bar + 8 * sb, 2 * tx);
                else
                  mbar_arrive_cluster(full_leader);
              }
            }
            __syncwarp();
            if (++sb == (uint32_t)stages) { sb = 0; pb ^= 1; }
            kc += BK;
            if (kc >= p.K) { kc = 0; ++tap; }
          }
          DD_GT(22, tile);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- UMMA issuer --------------------------------
    if (crank == 0) {   // CG = 2: only the leader CTA issues (its UMMAs drive both SMs); warp-uniform loop, elected issue
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;   // ring (slot, parity) counters
      uint32_t local_tile = 0;
      // descriptor = constant fields | (address >> 4): shared-memory addresses are < 256 KB, so no masking is needed
      const uint64_t DESC0 = umma_smem_desc(0, 16, 1024, 2);
      WorkIter wi(worker, n_workers, total_tiles, iters, p.sk_chunk, p.areuse);
      for (; wi.next(); ++local_tile) {
        const uint32_t as = local_tile & 1;
        const uint32_t aph = (local_tile >> 1) & 1;
        DD_GT(10, wi.tile);
        mbar_wait(tempty_bar + 8 * as, aph ^ 1);  // epilogue drained this accumulator
        tc_fence_after();
        DD_GT(11, wi.tile);
        const uint32_t tmem_d = tmem_base + as * ACC_COLS;
        uint32_t fresh = 0;                       // becomes 1 after the first UMMA of this segment
        if constexpr (MERGED) {
          [[maybe_unused]] long long w_b = 0;
          for (int it = wi.it0; it < wi.it1; ++it) {
            DD_GT_TIMED(w_b, mbar_wait(full_bar + 8 * sb, pb));
            tc_fence_after();
            if (it == wi.it0) DD_GT(12, wi.tile);
            const uint32_t sA = smem_base + sb * CONV_SLOT_BYTES;
            if (elect_one()) {
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                // tap kw reads box rows [kw, kw + 128): start address kw * 128 B into the (1024-aligned) slot
                const uint64_t dA = DESC0 + ((sA + kw * 128) >> 4);
                const uint64_t dB = DESC0 + ((sA + A_CONV_BYTES + kw * B_STAGE_BYTES) >> 4);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  const uint32_t acc = (k != 0 || kw != 0) ? 1u : fresh;
                  if constexpr (CG == 2) umma_bf16_2cta(tmem_d, dA + 2 * k, dB + 2 * k, IDESC, acc);
                  else umma_bf16(tmem_d, dA + 2 * k, dB + 2 * k, IDESC, acc);
                }
              }
              // the slot (in both CTAs of a pair) is free once the twelve UMMAs retire
              if constexpr (CG == 2) umma_commit_2cta(empty_bar + 8 * sb, 3); else umma_commit(empty_bar + 8 * sb);
            }
            __syncwarp();
            fresh = 1;
            if (++sb == (uint32_t)stages) { sb = 0; pb ^= 1; }
          }
          DD_GT_VAL(15, wi.tile, w_b);
        } else if constexpr (CONV) {
          [[maybe_unused]] long long w_a = 0, w_b = 0;
          for (int it = wi.it0; it < wi.it1; ++it) {
            DD_GT_TIMED(w_a, mbar_wait(afull_bar + 8 * sa, pa));
            if (it == wi.it0) DD_GT(12, wi.tile);
            const uint32_t sA = smem_base + sa * A_CONV_BYTES;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              DD_GT_TIMED(w_b, mbar_wait(full_bar + 8 * sb, pb));
              tc_fence_after();
              // tap kw reads box rows [kw, kw + 128): start address kw * 128 B into the (1024-aligned) slot
              const uint64_t dA = DESC0 + ((sA + kw * 128) >> 4);
              const uint64_t dB = DESC0 + ((ring_b + sb * B_STAGE_BYTES) >> 4);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  const uint32_t acc = (k != 0) ? 1u : fresh;
                  if constexpr (CG == 2) umma_bf16_2cta(tmem_d, dA + 2 * k, dB + 2 * k, IDESC, acc);
                  else umma_bf16(tmem_d, dA + 2 * k, dB + 2 * k, IDESC, acc);
                }
                if constexpr (CG == 2) umma_commit_2cta(empty_bar + 8 * sb, 3); else umma_commit(empty_bar + 8 * sb);
                if (kw == 2) {   // the activation box is free once the third tap's UMMAs retire
                  if constexpr (CG == 2) umma_commit_2cta(aempty_bar + 8 * sa, 3); else umma_commit(aempty_bar + 8 * sa);
                }
              }
              __syncwarp();
              fresh = 1;
              if (++sb == (uint32_t)stages) { sb = 0; pb ^= 1; }
            }
            if (++sa == (uint32_t)p.a_stages) { sa = 0; pa ^= 1; }
          }
          DD_GT_VAL(14, wi.tile, w_a);
          DD_GT_VAL(15, wi.tile, w_b);
        } else {
          for (int it = wi.it0; it < wi.it1; ++it) {
            mbar_wait(full_bar + 8 * sb, pb);
            tc_fence_after();
            if (it == wi.it0) DD_GT(12, wi.tile);
            const uint32_t sA = smem_base + sb * STAGE_BYTES;
            const uint64_t dA = DESC0 + (sA >> 4);
            const uint64_t dB = DESC0 + ((sA + A_STAGE_BYTES) >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                // advance 16 elements (32 B) along K inside the 128B swizzle atom: +2 in the >>4 address field
                const uint32_t acc = (k != 0) ? 1u : fresh;
                if constexpr (CG == 2) umma_bf16_2cta(tmem_d, dA + 2 * k, dB + 2 * k, IDESC, acc);
                else umma_bf16(tmem_d, dA + 2 * k, dB + 2 * k, IDESC, acc);
              }
              // frees the smem slot (in both CTAs of a pair) when these UMMAs retire
              if constexpr (CG == 2) umma_commit_2cta(empty_bar + 8 * sb, 3); else umma_commit(empty_bar + 8 * sb);
            }
            __syncwarp();
            fresh = 1;
            if (++sb == (uint32_t)stages) { sb = 0; pb ^= 1; }
          }
        }
        // accumulator complete (signalled to the epilogue warps of both CTAs of a pair)
        if (elect_one()) {
          if constexpr (CG == 2) umma_commit_2cta(tfull_bar + 8 * as, 3); else umma_commit(tfull_bar + 8 * as);
        }
        __syncwarp();
        DD_GT(13, wi.tile);
      }
    }
  } else {
    // ------------------------------- epilogue warps ------------------------------
    // TMEM lane == tile row, so a thread owns one output row; storing straight from that mapping writes 16 B per
    // lane at a row pitch of N*2 bytes (every warp store touches 32 cache lines).  Instead each warp transposes
    // 32x64 fp32 sub-tiles through a private, bank-rotated smem buffer so that 8 consecutive lanes own 64 consecutive
    // columns of ONE row: residual/bias loads and the bf16 stores become full 128-byte row segments.
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    uint32_t local_tile = 0;
    const int HW1 = (p.conv_H + 1) * pitch;
    const uint32_t stage_buf = epi_base + (uint32_t)(warp - 2) * EPI_WARP_BYTES;
    const int ehalf = (warp - 2) >> 2;                       // warps 2-5: even 64-column chunks, warps 6-9: odd ones
    const int n_ehalves = (int)(blockDim.x - 64) / 128;
    const int OUTW = p.geglu ? BN / 2 : BN;
    if (p.tma_epi) {
      // ---------------- TMA epilogue (plain GEMM, bf16 out): everything stays in the row-per-thread domain ----------
      // Per 64-column chunk: TMEM -> regs (both 32-column halves up front), + bias, + residual, activation, bf16 pack ->
      // swizzled staging tile -> TMA store.  The timeline of a CTA (profiles/gemm_trace.py, profiles/r02_gemm_trace.txt)
      // showed the short-K GEMMs bound by this loop, not by their operands (M=134400 N=320 K=320: 41 us, 29 us with the
      // accumulators merely handed back): a 192-wide tile has three chunks, so the warps of one half did two per tile
      // while the others idled 40 % of the time; the second tile of N = 320 computed a chunk that lies wholly beyond N;
      // every bulk-async instruction (wait_group.read, fence.proxy.async, store + commit) costs 200-300 cycles, two
      // integer divisions per tile another 300.  Hence:
      //   * a warp owns TWO 4 KB staging tiles used alternately; the residual tile of a chunk is TMA-loaded INTO the tile
      //     its output will be staged in (the thread reads and overwrites its own 16-byte units), a whole chunk ahead:
      //     issued right after the store of the previous chunk from that tile has released it;
      //   * chunks beyond N are skipped, and the odd chunk of a tile alternates between the two halves from tile to tile
      //     (the two accumulator buffers let one half run a tile ahead of the other);
      //   * the accumulator goes back to the UMMA warp as soon as the last chunk's TMEM loads have landed;
      //   * tile coordinates are carried as (m, n) counters.
      if constexpr (BN % 64 == 0 && !CONV) {
        const int half = (warp - 2) >> 2;     // warps 2-5 / 6-9: two warps per TMEM lane quarter (and per scheduler)
        const uint32_t my_stage = epi_base + (uint32_t)(warp - 2) * EPI_WARP_BYTES;   // two 32-row x 128 B tiles, SWIZZLE_128B
        const uint32_t r1_bar = smem_u32(&bars[20 + 2 * (warp - 2)]);                 // [2]: residual landed in tile 0 / 1
        const bool geglu = p.geglu != 0;
        const bool has_r1 = p.res1 != nullptr && !geglu;     // (the host rejects GEGLU with a residual)
        const int act = p.act;
        const float* __restrict__ bias = p.bias;
        const uint32_t xr = (uint32_t)(lane & 7);
        const int nch = geglu ? (BN / 128) : (BN / 64);
        // this worker's tiles: worker, worker + n_workers, ... or (A reuse) a contiguous range of the tile order
        const int t_step = p.areuse ? 1 : n_workers;
        const int t_begin = p.areuse ? (int)(((long long)worker * total_tiles) / n_workers) : worker;
        const int t_end = p.areuse ? (int)(((long long)(worker + 1) * total_tiles) / n_workers) : total_tiles;
        const int step_m = t_step / p.n_tiles, step_n = t_step - step_m * p.n_tiles;
        // chunk cursor: (tile, its (m, n) tile coordinates, index of the tile in this worker's sequence, chunk)
        struct Cur { int tile, m, n, lt, ch; };
        auto n_valid = [&](int n_idx) {            // chunks of this n-tile that hold columns < N
          if (geglu) return nch;
          const int v = (p.N - n_idx * BN + 63) >> 6;
          return v < nch ? v : nch;
        };
        auto next_tile = [&](Cur& c) {
          c.tile += t_step; ++c.lt;
          c.m += step_m; c.n += step_n;
          if (c.n >= p.n_tiles) { c.n -= p.n_tiles; ++c.m; }
        };
        auto next_chunk = [&](Cur& c) {            // the next chunk THIS warp processes; false when there is none
          c.ch += 2;
          while (c.tile < t_end) {
            if (c.ch < n_valid(c.n)) return true;
            next_tile(c);
            c.ch = (half + c.lt) & 1;              // the halves swap the even / odd chunks from tile to tile
          }
          return false;
        };
        auto r1_issue = [&](const Cur& c, uint32_t k) {   // residual box of chunk c -> staging tile k & 1 (lane 0)
          mbar_arrive_expect_tx(r1_bar + 8 * (k & 1u), 4096);
          tma_load_2d(my_stage + (k & 1u) * 4096u, &tmR1, r1_bar + 8 * (k & 1u), c.n * BN + c.ch * 64,
                      c.m * TILE_M + (int)crank * BM + q * 32);
        };
        // one 32-column half: + bias, + residual (in place), activation, bf16 pack -> 4 x 16 B of this thread's staging row
        auto finish_half = [&](uint32_t (&a)[32], int ncol0, int h, uint32_t buf) {
          if (bias) {
            // columns of the last tile beyond N are never stored, but their bias must not be READ either: the overhang
            // (e.g. N = 320 on 192-wide tiles) would run past the end of the bias vector
            const int ncol = p.N - ncol0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
              if (j < ncol) b = __ldg(reinterpret_cast<const float4*>(bias + ncol0 + j));
              a[j] = __float_as_uint(__uint_as_float(a[j]) + b.x);
              a[j + 1] = __float_as_uint(__uint_as_float(a[j + 1]) + b.y);
              a[j + 2] = __float_as_uint(__uint_as_float(a[j + 2]) + b.z);
              a[j + 3] = __float_as_uint(__uint_as_float(a[j + 3]) + b.w);
            }
          }
          if (has_r1) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {  // 4 x 16 B = 32 bf16 of this row
              const uint32_t uu = (uint32_t)((h >> 3) + u);
              uint32_t t0, t1, t2, t3;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3)
                           : "r"(buf + lane * 128 + ((uu ^ xr) << 4)));
              const uint32_t w[4] = {t0, t1, t2, t3};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = unpack_bf16(w[e]);
                const int j = u * 8 + e * 2;
                a[j] = __float_as_uint(__uint_as_float(a[j]) + f.x);
                a[j + 1] = __float_as_uint(__uint_as_float(a[j + 1]) + f.y);
              }
            }
          }
          if (act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) a[j] = __float_as_uint(silu_f(__uint_as_float(a[j])));
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t uu = (uint32_t)((h >> 3) + u);
            const uint32_t addr = buf + lane * 128 + ((uu ^ xr) << 4);
            const uint32_t k0 = pack_bf16(__uint_as_float(a[8 * u]), __uint_as_float(a[8 * u + 1]));
            const uint32_t k1 = pack_bf16(__uint_as_float(a[8 * u + 2]), __uint_as_float(a[8 * u + 3]));
            const uint32_t k2 = pack_bf16(__uint_as_float(a[8 * u + 4]), __uint_as_float(a[8 * u + 5]));
            const uint32_t k3 = pack_bf16(__uint_as_float(a[8 * u + 6]), __uint_as_float(a[8 * u + 7]));
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(k0), "r"(k1), "r"(k2), "r"(k3)
                         : "memory");
          }
        };
        uint32_t cseq = 0;                          // chunks processed by this warp (staging tile = cseq & 1)
        Cur pf;                                     // the chunk whose residual is requested next (one ahead of the work)
        pf.tile = t_begin; pf.lt = 0;
        pf.m = t_begin / p.n_tiles; pf.n = t_begin - pf.m * p.n_tiles;
        pf.ch = ((half + 0) & 1) - 2;
        Cur cur = pf;                               // the tile being drained (cur.ch is set per tile below)
        bool pf_ok = next_chunk(pf);
        if (has_r1 && pf_ok) {
          if (lane == 0) r1_issue(pf, 0);
          pf_ok = next_chunk(pf);
        }
        for (; cur.tile < t_end; next_tile(cur), ++local_tile) {
          const int tile = cur.tile;
          const int m0 = cur.m * TILE_M + (int)crank * BM;
          const int n0 = cur.n * BN;
          const int nout0 = geglu ? cur.n * (BN / 2) : n0;
          const uint32_t as = local_tile & 1;
          const uint32_t aph = (local_tile >> 1) & 1;
          const int nv = n_valid(cur.n);
          // residual rows of the NEXT tile -> L2 now (a whole tile ahead of their TMA loads)
          if (has_r1 && lane == 0 && tile + t_step < t_end) {
            Cur t2 = cur;
            next_tile(t2);
            const int nv2 = n_valid(t2.n);
            for (int ch = (half + t2.lt) & 1; ch < nv2; ch += 2)
              tma_prefetch_2d(&tmR1, t2.n * BN + ch * 64, t2.m * TILE_M + (int)crank * BM + q * 32);
          }
          DD_GT(0, tile);
          mbar_wait(tfull_bar + 8 * as, aph);
          tc_fence_after();
          DD_GT(1, tile);
          const uint32_t tmem_acc = tmem_base + as * ACC_COLS + ((uint32_t)(q * 32) << 16);
          bool released = false;
          auto release_acc = [&]() {   // every TMEM load of this warp for this tile has landed
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (crank == 0) mbar_arrive(tempty_bar + 8 * as);
              else mbar_arrive_cluster(mapa_shared(tempty_bar + 8 * as, 0));
            }
            released = true;
          };
#pragma unroll 1
          for (int ch = (half + (int)local_tile) & 1; ch < nv; ch += 2) {
            const int c = ch * 64;
            const uint32_t buf = my_stage + (cseq & 1u) * 4096u;
            const bool last_chunk = ch + 2 >= nv;
            if (geglu) {
              if (lane == 0) bulk_wait_read1();   // the store issued two chunks ago has finished reading this staging tile
              __syncwarp();
#pragma unroll 1
              for (int h = 0; h < 64; h += 32) {
                uint32_t a[32], g[32];
                tmem_ld_32x32(tmem_acc + c + h, a);
                tmem_ld_32x32(tmem_acc + BN / 2 + c + h, g);
                tmem_ld_wait();
                if (last_chunk && h == 32) release_acc();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), bg = bv;
                  if (bias) {
                    bv = __ldg(reinterpret_cast<const float4*>(bias + n0 + c + h + j));
                    bg = __ldg(reinterpret_cast<const float4*>(bias + n0 + BN / 2 + c + h + j));
                  }
                  float r0, r1;
                  unpack_f32x2(geglu_pair(add_f32x2(pack_f32x2(__uint_as_float(a[j]), __uint_as_float(a[j + 1])), pack_f32x2(bv.x, bv.y)),
                                          add_f32x2(pack_f32x2(__uint_as_float(g[j]), __uint_as_float(g[j + 1])), pack_f32x2(bg.x, bg.y))), r0, r1);
                  a[j] = __float_as_uint(r0); a[j + 1] = __float_as_uint(r1);
                  unpack_f32x2(geglu_pair(add_f32x2(pack_f32x2(__uint_as_float(a[j + 2]), __uint_as_float(a[j + 3])), pack_f32x2(bv.z, bv.w)),
                                          add_f32x2(pack_f32x2(__uint_as_float(g[j + 2]), __uint_as_float(g[j + 3])), pack_f32x2(bg.z, bg.w))), r0, r1);
                  a[j + 2] = __float_as_uint(r0); a[j + 3] = __float_as_uint(r1);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const uint32_t uu = (uint32_t)((h >> 3) + u);
                  const uint32_t addr = buf + lane * 128 + ((uu ^ xr) << 4);
                  const uint32_t k0 = pack_bf16(__uint_as_float(a[8 * u]), __uint_as_float(a[8 * u + 1]));
                  const uint32_t k1 = pack_bf16(__uint_as_float(a[8 * u + 2]), __uint_as_float(a[8 * u + 3]));
                  const uint32_t k2 = pack_bf16(__uint_as_float(a[8 * u + 4]), __uint_as_float(a[8 * u + 5]));
                  const uint32_t k3 = pack_bf16(__uint_as_float(a[8 * u + 6]), __uint_as_float(a[8 * u + 7]));
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(k0), "r"(k1), "r"(k2), "r"(k3)
                               : "memory");
                }
              }
            } else {
              uint32_t a0[32], a1[32];
              tmem_ld_32x32(tmem_acc + c, a0);
              tmem_ld_32x32(tmem_acc + c + 32, a1);
              if (has_r1) {
                // the store of the previous chunk has released the OTHER staging tile: the residual of this warp's next
                // chunk goes there now and has this whole chunk to land
                if (lane == 0) {
                  bulk_wait_read0();
                  if (pf_ok) r1_issue(pf, cseq + 1);
                }
                if (pf_ok) pf_ok = next_chunk(pf);
                __syncwarp();
                mbar_wait(r1_bar + 8 * (cseq & 1u), (cseq >> 1) & 1u);   // this chunk's residual has landed in `buf`
              } else {
                if (lane == 0) bulk_wait_read1();   // the store issued two chunks ago has finished reading this staging tile
                __syncwarp();
              }
              DD_GT(2, tile);
              tmem_ld_wait();
              DD_GT(3, tile);
              if (last_chunk) release_acc();
              finish_half(a0, n0 + c, 0, buf);
              DD_GT(4, tile);
              finish_half(a1, n0 + c + 32, 32, buf);
              DD_GT(5, tile);
            }
            ++cseq;
            fence_proxy_async_smem();
            __syncwarp();
            DD_GT(6, tile);
            if (lane == 0) {
              tma_store_2d(&tmOut, buf, nout0 + c, m0 + q * 32);  // clipped at [M, n_store] by the tensor map
              bulk_commit();
            }
            DD_GT(7, tile);
          }
          if (!released) release_acc();   // no chunk of this tile fell to this warp
        }
        if (lane == 0) bulk_wait0();
      }
    } else {
    // ---------------- register epilogue (conv scatter, fp32 out, per-image vector, stream-K partials) ----------------
    // Per 64-column chunk: TMEM -> regs -> bank-rotated smem (thread = row) -> regs (8 lanes = one row's 64 columns) ->
    // bias / per-image vector / residual / activation -> 16-byte row-segment stores.  Residual rows are pulled into L2
    // one tile ahead (prefetch.global.L2) and into registers one chunk ahead, the first chunk's before the wait on the
    // accumulator, so their HBM latency never sits on the tile's critical path.
    WorkIter wi(worker, n_workers, total_tiles, iters, p.sk_chunk, p.areuse);
    const bool sk = p.sk_chunk != 0;
    const int n_store = sk ? BN : p.n_store;
    const int out_ld = sk ? BN : (int)p.out_ld;
    void* const outp = sk ? (void*)p.ws : p.out;
    const bool has_r1 = p.res1 != nullptr && !p.geglu;
    const int rowt = (int)crank * BM + q * 32 + lane;   // row of this thread inside the tile

    // padded-pixel row m -> (valid, compact output row)
    // (also the row of the per-image vector that belongs to it, output row / rows_per_img -- the caller may share one vector
    // between several images: carried with the row so that the store loop divides nothing; it used to run a 64-bit
    // division per row segment, a third of its instructions)
    auto map_row = [&](int tile, int& valid, int& orow, int& img) {
      const int m = (tile / p.n_tiles) * TILE_M + rowt;
      valid = m < p.M;
      orow = m;
      img = 0;
      if (sk) {
        // raw partial accumulator -> workspace slot of (tile, contributor); every row of the tile is written
        const int slot = tile * p.sk_maxc + (worker - (tile * iters) / p.sk_chunk);
        valid = 1;
        orow = slot * TILE_M + rowt;
      } else if (p.taps == 9) {
        const int im = m / HW1;
        const int rem = m - im * HW1;
        const int hp = rem / pitch;
        const int wp = rem - hp * pitch;
        valid = valid && (hp < p.conv_H) && (wp < p.conv_W);
        orow = (im * p.conv_H + hp) * p.conv_W + wp;
      }
      if (p.rowvec && valid && !sk) img = orow / p.rows_per_img;
    };
    // chunk geometry of this lane: 8 lanes own the 64 columns of one row (4 lanes for a 32-column tail chunk)
    struct Chunk { int lpr, rpi, nit, col; bool vec; };
    auto chunk_of = [&](int nout0, int c) {
      Chunk g;
      const int width = (OUTW - c) < 64 ? (OUTW - c) : 64;
      g.lpr = width >> 3;
      g.rpi = 32 / g.lpr;
      g.nit = 32 / g.rpi;
      g.col = nout0 + c + (lane % g.lpr) * 8;
      g.vec = g.col + 8 <= n_store;
      return g;
    };
    int orow_j[8], img_j[8];
    uint32_t vmask = 0;
    uint4 r1v[8];
    auto prefetch_chunk = [&](const Chunk& g, int valid, int orow, int img) {
      vmask = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < g.nit) {
          const int rr = j * g.rpi + lane / g.lpr;
          orow_j[j] = __shfl_sync(0xffffffffu, orow, rr);
          img_j[j] = __shfl_sync(0xffffffffu, img, rr);
          const int vj = __shfl_sync(0xffffffffu, valid, rr) && (g.col < n_store);
          vmask |= (uint32_t)vj << j;
          r1v[j] = make_uint4(0u, 0u, 0u, 0u);
          if (has_r1 && g.vec && vj) r1v[j] = *reinterpret_cast<const uint4*>(p.res1 + (long long)orow_j[j] * p.res1_ld + g.col);
        }
      }
    };

    bool have = wi.next();
    int valid = 0, orow = 0, img = 0;
    if (have) map_row(wi.tile, valid, orow, img);
    for (; have; ++local_tile) {
      const int tile = wi.tile;
      const int n0 = (tile % p.n_tiles) * BN;
      const int nout0 = sk ? 0 : p.geglu ? (n0 / BN) * (BN / 2) : n0;
      const uint32_t as = local_tile & 1;
      const uint32_t aph = (local_tile >> 1) & 1;
      // the two warp halves swap the even / odd 64-column chunks from tile to tile: a 160-wide tile has 64 + 64 + 32
      // columns, and the two accumulator buffers let one half run a tile ahead of the other
      const int eh = n_ehalves == 2 ? ((ehalf + (int)local_tile) & 1) : 0;
      Chunk g = chunk_of(nout0, eh * 64);
      if (eh * 64 < OUTW) prefetch_chunk(g, valid, orow, img);        // first chunk's residual rows: before the wait
      // next tile: row mapping now, residual rows into L2 (one tile = several microseconds ahead)
      have = wi.next();
      int valid_n = 0, orow_n = 0, img_n = 0;
      if (have) {
        map_row(wi.tile, valid_n, orow_n, img_n);
        if (has_r1 && valid_n) {
          const int nn0 = (wi.tile % p.n_tiles) * BN;
          const bf16* rp = p.res1 + (long long)orow_n * p.res1_ld + nn0;
          for (int cc = (n_ehalves == 2 ? (eh ^ 1) : 0) * 64; cc < OUTW && nn0 + cc < n_store; cc += 64 * n_ehalves)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + cc));
        }
      }
      DD_GT(0, tile);
      mbar_wait(tfull_bar + 8 * as, aph);
      tc_fence_after();
      DD_GT(1, tile);
      const uint32_t tmem_acc = tmem_base + as * ACC_COLS + ((uint32_t)(q * 32) << 16);

#pragma unroll 1
      for (int c = eh * 64; c < OUTW; c += 64 * n_ehalves) {
        const int width = (OUTW - c) < 64 ? (OUTW - c) : 64;  // 64, or 32 for the last chunk of BN = 32/160
        const int k = lane % g.lpr;
        const int col = g.col;
        float4 bia0 = make_float4(0.f, 0.f, 0.f, 0.f), bia1 = bia0;
        if (g.vec && !p.geglu && p.bias) {
          bia0 = *reinterpret_cast<const float4*>(p.bias + col);
          bia1 = *reinterpret_cast<const float4*>(p.bias + col + 4);
        }
        // ---- TMEM -> registers -> rotated smem (thread = row) ----
#pragma unroll 1
        for (int h = 0; h < width; h += 32) {
          uint32_t a[32];
          tmem_ld_32x32(tmem_acc + c + h, a);
          if (p.geglu) {
            uint32_t gt[32];
            tmem_ld_32x32(tmem_acc + BN / 2 + c + h, gt);
            tmem_ld_wait();
#pragma unroll 4
            for (int j = 0; j < 32; ++j) {
              float v = __uint_as_float(a[j]), gg = __uint_as_float(gt[j]);
              if (p.bias) {
                v += p.bias[n0 + c + h + j];
                gg += p.bias[n0 + BN / 2 + c + h + j];
              }
              a[j] = __float_as_uint(v * gelu_erf_f(gg));
            }
          } else {
            tmem_ld_wait();
          }
          const uint32_t rbase = stage_buf + lane * 256;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int uu = (h >> 2) + u;  // 16-byte unit index inside the 256-byte row
            const uint32_t pu = (uint32_t)((uu & 8) | (((uu & 7) + (uu >> 3) + lane) & 7));
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + (pu << 4)), "r"(a[4 * u]),
                         "r"(a[4 * u + 1]), "r"(a[4 * u + 2]), "r"(a[4 * u + 3])
                         : "memory");
          }
        }
        __syncwarp();
        DD_GT(8, tile);
        // ---- smem -> registers (8 lanes = one row's 64 columns) -> epilogue math -> coalesced global store ----
        if (nout0 + c + width > n_store) {
          // ragged right edge inside this chunk (N not a multiple of 8, e.g. conv_out N = 4): every lane takes a compact
          // scalar loop straight from the staging rows (warp-uniform branch; rare)
#pragma unroll 1
          for (int j = 0; j < g.nit; ++j) {
            const int rr = j * g.rpi + lane / g.lpr;
            const long long orow_r = __shfl_sync(0xffffffffu, orow, rr);
            const int vr = __shfl_sync(0xffffffffu, valid, rr);
            const int img_r = __shfl_sync(0xffffffffu, img, rr);
#pragma unroll 1
            for (int e = 0; e < 8; ++e) {
              const int ci = k * 8 + e, uu = ci >> 2;
              if (!vr || col + e >= n_store) continue;
              const uint32_t pu = (uint32_t)((uu & 8) | (((uu & 7) + (uu >> 3) + rr) & 7));
              float x;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(stage_buf + rr * 256 + (pu << 4) + (ci & 3) * 4));
              if (!p.geglu) {
                if (p.bias) x += p.bias[col + e];
                if (p.rowvec) x += p.rowvec[(long long)img_r * p.rowvec_ld + col + e];
                if (p.res1) x += __bfloat162float(p.res1[orow_r * p.res1_ld + col + e]);
                if (p.res2) x += __bfloat162float(p.res2[orow_r * p.res2_ld + col + e]);
                if (p.act == 1) x = silu_f(x);
              }
              if (p.out_f32)
                reinterpret_cast<float*>(outp)[orow_r * out_ld + col + e] = x;
              else
                reinterpret_cast<bf16*>(outp)[orow_r * out_ld + col + e] = __float2bfloat16(x);
            }
          }
        } else if (g.vec) {
          int rv_img = -1;
          float4 rv0 = make_float4(0.f, 0.f, 0.f, 0.f), rv1 = rv0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j < g.nit && ((vmask >> j) & 1u)) {
              const int rr = j * g.rpi + lane / g.lpr;
              const long long orow_r = orow_j[j];
              const uint32_t rbase = stage_buf + rr * 256;
              const int u0 = 2 * k, u1 = 2 * k + 1;
              const uint32_t p0 = (uint32_t)((u0 & 8) | (((u0 & 7) + (u0 >> 3) + rr) & 7));
              const uint32_t p1 = (uint32_t)((u1 & 8) | (((u1 & 7) + (u1 >> 3) + rr) & 7));
              float v[8];
              {
                uint32_t t0, t1, t2, t3;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3) : "r"(rbase + (p0 << 4)));
                v[0] = __uint_as_float(t0); v[1] = __uint_as_float(t1); v[2] = __uint_as_float(t2); v[3] = __uint_as_float(t3);
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3) : "r"(rbase + (p1 << 4)));
                v[4] = __uint_as_float(t0); v[5] = __uint_as_float(t1); v[6] = __uint_as_float(t2); v[7] = __uint_as_float(t3);
              }
              if (!p.geglu) {
                v[0] += bia0.x; v[1] += bia0.y; v[2] += bia0.z; v[3] += bia0.w;
                v[4] += bia1.x; v[5] += bia1.y; v[6] += bia1.z; v[7] += bia1.w;
                if (p.rowvec) {
                  const int im = img_j[j];
                  if (im != rv_img) {   // rows ascend: the per-image vector is reloaded only when the image changes
                    const float* rv = p.rowvec + (long long)im * p.rowvec_ld + col;
                    rv0 = *reinterpret_cast<const float4*>(rv);
                    rv1 = *reinterpret_cast<const float4*>(rv + 4);
                    rv_img = im;
                  }
                  v[0] += rv0.x; v[1] += rv0.y; v[2] += rv0.z; v[3] += rv0.w;
                  v[4] += rv1.x; v[5] += rv1.y; v[6] += rv1.z; v[7] += rv1.w;
                }
                if (p.res1) {
                  float2 t;
                  t = unpack_bf16(r1v[j].x); v[0] += t.x; v[1] += t.y;
                  t = unpack_bf16(r1v[j].y); v[2] += t.x; v[3] += t.y;
                  t = unpack_bf16(r1v[j].z); v[4] += t.x; v[5] += t.y;
                  t = unpack_bf16(r1v[j].w); v[6] += t.x; v[7] += t.y;
                }
                if (p.res2) {   // rare (second residual source): loaded in place
                  const uint4 r = *reinterpret_cast<const uint4*>(p.res2 + orow_r * p.res2_ld + col);
                  float2 t;
                  t = unpack_bf16(r.x); v[0] += t.x; v[1] += t.y;
                  t = unpack_bf16(r.y); v[2] += t.x; v[3] += t.y;
                  t = unpack_bf16(r.z); v[4] += t.x; v[5] += t.y;
                  t = unpack_bf16(r.w); v[6] += t.x; v[7] += t.y;
                }
                if (p.act == 1) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[e] = silu_f(v[e]);
                }
              }
              if (p.out_f32) {
                float* o = reinterpret_cast<float*>(outp) + orow_r * out_ld + col;
                *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
              } else {
                bf16* o = reinterpret_cast<bf16*>(outp) + orow_r * out_ld + col;
                *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                                                          pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
              }
            }
          }
        }
        __syncwarp();
        DD_GT(9, tile);
        // next chunk of this warp: its residual rows go to registers now
        if (c + 64 * n_ehalves < OUTW) {
          g = chunk_of(nout0, c + 64 * n_ehalves);
          prefetch_chunk(g, valid, orow, img);
        }
      }
      // release the accumulator buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (crank == 0) mbar_arrive(tempty_bar + 8 * as);
        else mbar_arrive_cluster(mapa_shared(tempty_bar + 8 * as, 0));
      }
      valid = valid_n;
      orow = orow_n;
      img = img_n;
    }
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();   // pair: nobody leaves while the peer may still signal us
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_2cta(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// stream-K second pass: out = epilogue( sum over the contributors of a tile, in worker order )
// grid (tiles, TILE_M / 32), 256 threads; a thread owns 4 consecutive columns of a row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const GemmDev p, const int BN, const int TILE_M, const int iters) {
  const int tile = blockIdx.x;
  const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
  const int w_first = (tile * iters) / p.sk_chunk;
  const int w_last = ((tile + 1) * iters - 1) / p.sk_chunk;
  const int nc = w_last - w_first + 1;
  const int tpr = BN >> 2;                    // threads per row
  const int rpi = 256 / tpr;                  // rows per iteration
  const int cq = threadIdx.x % tpr;
  const int col = n_tile * BN + cq * 4;
  if (col >= p.n_store || (int)threadIdx.x >= rpi * tpr) return;
  const int pitch = p.conv_W + 1;
  const int HW1 = (p.conv_H + 1) * pitch;
  const float* base = p.ws + (size_t)tile * p.sk_maxc * TILE_M * BN + cq * 4;
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias) bv = *reinterpret_cast<const float4*>(p.bias + col);
  for (int r = blockIdx.y * 32 + threadIdx.x / tpr; r < blockIdx.y * 32 + 32; r += rpi) {
    const int m = m_tile * TILE_M + r;
    if (m >= p.M) break;
    long long orow = m;
    if (p.taps == 9) {
      const int img = m / HW1;
      const int rem = m - img * HW1;
      const int hp = rem / pitch;
      const int wp = rem - hp * pitch;
      if (hp >= p.conv_H || wp >= p.conv_W) continue;
      orow = ((long long)img * p.conv_H + hp) * p.conv_W + wp;
    }
    float4 v = *reinterpret_cast<const float4*>(base + (size_t)r * BN);
    for (int k = 1; k < nc; ++k) {
      const float4 t = *reinterpret_cast<const float4*>(base + ((size_t)k * TILE_M + r) * BN);
      v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
    if (p.rowvec) {
      const float4 t = *reinterpret_cast<const float4*>(p.rowvec + (orow / p.rows_per_img) * (long long)p.rowvec_ld + col);
      v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    if (p.res1) {
      const uint2 t = *reinterpret_cast<const uint2*>(p.res1 + orow * p.res1_ld + col);
      const float2 a = unpack_bf16(t.x), b = unpack_bf16(t.y);
      v.x += a.x; v.y += a.y; v.z += b.x; v.w += b.y;
    }
    if (p.res2) {
      const uint2 t = *reinterpret_cast<const uint2*>(p.res2 + orow * p.res2_ld + col);
      const float2 a = unpack_bf16(t.x), b = unpack_bf16(t.y);
      v.x += a.x; v.y += a.y; v.z += b.x; v.w += b.y;
    }
    if (p.act == 1) {
      v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w);
    }
    if (p.out_f32) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow * p.out_ld + col) = v;
    } else {
      *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out) + orow * p.out_ld + col) =
          make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
template <int BN, int CG, bool CONV>
static int launch_gemm_cg(const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB,
                          const CUtensorMap& tmOut, const CUtensorMap& tmR1, GemmDev p, cudaStream_t stream) {
  constexpr int B_BYTES = (BN / CG) * BK * 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_BYTES;
  const int epi_warps = 8;   // both epilogues run 8 warps (the 4-warp register epilogue measured slower, round 1)
  const int avail = 227 * 1024 - 3072 - epi_warps * EPI_WARP_BYTES;
  const int kchunks = (p.K + BK - 1) / BK;
  const int iters = CONV ? 3 * kchunks : p.taps * kchunks;
  size_t smem;
  if constexpr (CONV && BN <= 192 && DD_CONV_MERGED) {
    // merged slots: activation box + the weight boxes of the three taps of a kernel row
    constexpr int SLOT = A_CONV_BYTES + 3 * B_BYTES;
    int st = avail / SLOT;
    st = st > 4 ? 4 : st;
    p.stages = st;
    p.a_stages = 0;
    smem = (size_t)st * SLOT + (size_t)epi_warps * EPI_WARP_BYTES + 1024;
  } else if constexpr (CONV) {
    int sb = (avail - 3 * A_CONV_BYTES) / B_BYTES;
    sb = sb > 8 ? 8 : sb < 2 ? 2 : sb;
    int sa = (avail - sb * B_BYTES) / A_CONV_BYTES;
    sa = sa > 4 ? 4 : sa < 2 ? 2 : sa;
    p.stages = sb;
    p.a_stages = sa;
    smem = (size_t)sa * A_CONV_BYTES + (size_t)sb * B_BYTES + (size_t)epi_warps * EPI_WARP_BYTES + 1024;
  } else {
    int stages = avail / STAGE_BYTES;
    if (stages > 8) stages = 8;
    if (stages > iters && iters >= 2) stages = iters;
    if (stages < 2) stages = 2;
    p.stages = stages;
    p.a_stages = 0;
    smem = (size_t)stages * STAGE_BYTES + (size_t)epi_warps * EPI_WARP_BYTES + 1024;
  }
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(gemm_tcgen05_kernel<BN, CG, CONV>), 227 * 1024 - 2048)) return rc;
  p.m_tiles = (p.M + BM * CG - 1) / (BM * CG);   // tiles of 128 (one CTA) or 256 (CTA pair) rows
  p.n_tiles = (p.N + BN - 1) / BN;
  // short K (one ring slot per k-block) and several n-tiles per m-tile: contiguous tile ranges with the A blocks kept in the ring
  p.areuse = (DD_GEMM_AREUSE && !CONV && p.taps == 1 && p.sk_chunk == 0 && iters >= 2 && p.stages == iters && p.n_tiles >= 2) ? 1 : 0;
  int workers = p.m_tiles * p.n_tiles;
  const int sms = num_sms();
  if (workers > sms / CG) workers = sms / CG;
  if (p.sk_chunk > 0) workers = (p.m_tiles * p.n_tiles * iters + p.sk_chunk - 1) / p.sk_chunk;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(workers * CG);
  cfg.blockDim = dim3(GEMM_THREADS_MAX);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;   // cluster shape + programmatic dependent launch (prologue overlaps the previous kernel's tail)
  DD_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BN, CG, CONV>, tmA, tmA2, tmB, tmOut, tmR1, p));
  return 0;
}

template <int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB,
                       const CUtensorMap& tmOut, const CUtensorMap& tmR1, GemmDev p, cudaStream_t stream, int cg) {
  if (p.taps == 9) {
    if (cg == 2) return launch_gemm_cg<BN, 2, true>(tmA, tmA2, tmB, tmOut, tmR1, p, stream);
    return launch_gemm_cg<BN, 1, true>(tmA, tmA2, tmB, tmOut, tmR1, p, stream);
  }
  if (cg == 2) return launch_gemm_cg<BN, 2, false>(tmA, tmA2, tmB, tmOut, tmR1, p, stream);
  return launch_gemm_cg<BN, 1, false>(tmA, tmA2, tmB, tmOut, tmR1, p, stream);
}

static int pick_bn_tma(int N, int geglu) {
  // TMA epilogue moves 64-column boxes -> tile widths are multiples of 64.  Small-K GEMMs are bound by operand
  // traffic ~ n_tiles * (128 + BN): prefer the widest tile unless it wastes a whole extra tile of columns.
  if (geglu) return 256;
  if (N <= 64) return 64;
  const int cands[] = {256, 192, 128};
  int best = 128;
  long best_cost = 1L << 60;
  for (int bn : cands) {
    const long cost = (long)((N + bn - 1) / bn) * (128 + bn);
    if (cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

static int pick_bn(int N, int geglu, int force) {
  if (force > 0) return force;
  if (geglu) return 256;
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  // choose the tile width that wastes the fewest columns, preferring wide tiles (less smem traffic / flop)
  const int cands[] = {256, 192, 160, 128};
  int best = 128;
  double best_cost = 1e30;
  for (int bn : cands) {
    const int tiles = (N + bn - 1) / bn;
    const double waste = (double)tiles * bn / N;           // >= 1
    const double smem_pen = (8192.0 / bn + 64.0) / 96.0;   // relative smem bytes per MMA cycle
    const double cost = waste * (0.75 + 0.25 * smem_pen);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

int gemm_run(const dd_gemm_args* a, cudaStream_t stream) {
  DD_CHECK(a != nullptr, -1, "dd_gemm: null args");
  DD_CHECK(a->M > 0 && a->N > 0 && a->K > 0, -1, "dd_gemm: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  DD_CHECK(a->taps == 1 || a->taps == 9, -1, "dd_gemm: taps must be 1 or 9 (got %d)", a->taps);
  DD_CHECK(a->K % 8 == 0, -1, "dd_gemm: K=%d must be a multiple of 8 (16-byte TMA rows)", a->K);
  DD_CHECK(a->a_ld % 8 == 0 && a->w_ld % 8 == 0, -1, "dd_gemm: leading dims must be multiples of 8");
  DD_CHECK(((uintptr_t)a->a & 15) == 0 && ((uintptr_t)a->w & 15) == 0 && ((uintptr_t)a->out & 15) == 0,
           -1, "dd_gemm: pointers must be 16-byte aligned");
  int K1 = a->K;
  if (a->a2 != nullptr) {
    K1 = a->k1;
    DD_CHECK(K1 > 0 && K1 < a->K && K1 % BK == 0, -1, "dd_gemm: dual-source split k1=%d must be a multiple of 64", K1);
    DD_CHECK(a->a2_ld % 8 == 0 && ((uintptr_t)a->a2 & 15) == 0, -1, "dd_gemm: a2 alignment");
  }
  if (a->taps == 9) DD_CHECK(a->conv_h > 0 && a->conv_w > 0, -1, "dd_gemm: conv geometry missing");
  if (a->geglu) DD_CHECK(a->N % 256 == 0 && !a->out_f32, -1, "dd_gemm: GEGLU needs N %% 256 == 0 and bf16 out");
  if (a->geglu) DD_CHECK(a->res1 == nullptr && a->res2 == nullptr && a->rowvec == nullptr && a->act == 0, -1,
                         "dd_gemm: the GEGLU epilogue takes a bias only");
  const int n_store = a->geglu ? a->N / 2 : a->N;
  DD_CHECK(a->out_ld >= n_store, -1, "dd_gemm: out_ld too small");
  if (n_store % 8 != 0) DD_CHECK(a->res1 == nullptr || true, -1, "unreachable");

  const bool tma_ok = a->taps == 1 && !a->out_f32 && n_store % 8 == 0 && a->res2 == nullptr && a->rowvec == nullptr &&
                      a->out_ld % 8 == 0 && (a->res1 == nullptr || (a->res1_ld % 8 == 0 && ((uintptr_t)a->res1 & 15) == 0)) &&
                      a->N > 32 && (a->force_bn == 0 || a->force_bn % 64 == 0) && !a->no_tma_epilogue;
  int bn = pick_bn(a->N, a->geglu, a->force_bn);
  if (tma_ok && a->force_bn == 0) bn = pick_bn_tma(a->N, a->geglu);
  // CTA pairs (cta_group::2, 256-row tiles) whenever there is more than one 128-row tile of work
  const int cg = (a->M > BM && !a->one_cta) ? 2 : 1;
  // stream-K: when the tiles fill the last wave badly (e.g. the 4x7-pixel level: 75 pair-tiles on 74 CTA pairs) and the
  // K loop is long, split the flattened (tile, k-iteration) space evenly over the workers instead
  int sk_chunk = 0, sk_maxc = 0;
  bool tma_epi = tma_ok;
  if (a->workspace != nullptr && a->stream_k >= 0 && !a->geglu && a->N % 4 == 0) {
    const int bn_sk = pick_bn(a->N, 0, a->force_bn);
    const int tile_m = BM * cg;
    const long tiles = (long)((a->M + tile_m - 1) / tile_m) * ((a->N + bn_sk - 1) / bn_sk);
    const long W = num_sms() / cg;
    const long kch = (a->K + BK - 1) / BK;
    const long iters = a->taps == 9 ? 3 * kch : kch;   // conv: one iteration = a kernel row x 64 channels (12 UMMAs)
    const long waves = (tiles + W - 1) / W;
    const double eff = (double)tiles / (double)(waves * W);
    if (a->stream_k == 1 || (iters * (a->taps == 9 ? 3 : 1) >= 32 && eff < 0.75 && tiles <= 4 * W)) {
      const long total = tiles * iters;
      const long chunk = (total + W - 1) / W;
      const long maxc = (iters + chunk - 2) / chunk + 1;
      const long long need = (long long)tiles * maxc * tile_m * bn_sk * 4;
      if (need <= a->workspace_bytes && total < (1L << 30)) {
        sk_chunk = (int)chunk;
        sk_maxc = (int)maxc;
        bn = bn_sk;
        tma_epi = false;
      } else {
        DD_CHECK(a->stream_k != 1, -1, "dd_gemm: stream-K workspace too small (%lld > %lld bytes)", need,
                 (long long)a->workspace_bytes);
      }
    }
  } else {
    DD_CHECK(a->stream_k != 1, -1, "dd_gemm: stream-K needs a workspace, no GEGLU and N %% 4 == 0");
  }
  CUtensorMap tmA, tmA2, tmB, tmOut, tmR1;
  const uint32_t a_box_rows = a->taps == 9 ? A_CONV_ROWS : BM;   // conv: 128 rows + the kw = -1 / +1 halo rows
  int rc = make_tmap_2d_bf16(&tmA, a->a, (uint64_t)a->M, (uint64_t)K1, (uint64_t)a->a_ld, a_box_rows, BK);
  if (rc) return rc;
  if (a->a2 != nullptr) {
    rc = make_tmap_2d_bf16(&tmA2, a->a2, (uint64_t)a->M, (uint64_t)(a->K - K1), (uint64_t)a->a2_ld, a_box_rows, BK);
    if (rc) return rc;
  } else {
    tmA2 = tmA;
  }
  rc = make_tmap_2d_bf16(&tmB, a->w, (uint64_t)a->N, (uint64_t)a->K * a->taps, (uint64_t)a->w_ld, bn / cg, BK);
  if (rc) return rc;

  if (tma_epi) {
    rc = make_tmap_2d_bf16(&tmOut, a->out, (uint64_t)a->M, (uint64_t)n_store, (uint64_t)a->out_ld, 32, 64);
    if (rc) return rc;
    if (a->res1 != nullptr) {
      rc = make_tmap_2d_bf16(&tmR1, a->res1, (uint64_t)a->M, (uint64_t)n_store, (uint64_t)a->res1_ld, 32, 64);
      if (rc) return rc;
    } else {
      tmR1 = tmOut;
    }
  } else {
    tmOut = tmA;
    tmR1 = tmA;
  }
  GemmDev p;
  p.tma_epi = tma_epi ? 1 : 0;
  p.sk_chunk = sk_chunk; p.sk_maxc = sk_maxc; p.ws = reinterpret_cast<float*>(a->workspace);
  p.M = a->M; p.N = a->N; p.K = a->K; p.K1 = K1; p.taps = a->taps;
  p.conv_H = a->conv_h; p.conv_W = a->conv_w;
  p.out = a->out; p.out_ld = a->out_ld; p.out_f32 = a->out_f32;
  p.bias = a->bias; p.rowvec = a->rowvec; p.rowvec_ld = a->rowvec_ld;
  p.rows_per_img = a->rows_per_img > 0 ? a->rows_per_img : 1;
  p.res1 = reinterpret_cast<const bf16*>(a->res1); p.res1_ld = a->res1_ld;
  p.res2 = reinterpret_cast<const bf16*>(a->res2); p.res2_ld = a->res2_ld;
  p.geglu = a->geglu; p.act = a->act; p.n_store = n_store;
  p.m_tiles = p.n_tiles = p.stages = 0; p.areuse = 0;
  if (sk_chunk > 0) {
    // pass 1: raw fp32 partial tiles into the workspace; pass 2: ordered sum + the real epilogue
    GemmDev pk = p;
    pk.bias = nullptr; pk.rowvec = nullptr; pk.res1 = nullptr; pk.res2 = nullptr; pk.act = 0; pk.out_f32 = 1;
    switch (bn) {
      case 32: rc = launch_gemm<32>(tmA, tmA2, tmB, tmOut, tmR1, pk, stream, cg); break;
      case 64: rc = launch_gemm<64>(tmA, tmA2, tmB, tmOut, tmR1, pk, stream, cg); break;
      case 128: rc = launch_gemm<128>(tmA, tmA2, tmB, tmOut, tmR1, pk, stream, cg); break;
      case 160: rc = launch_gemm<160>(tmA, tmA2, tmB, tmOut, tmR1, pk, stream, cg); break;
      case 192: rc = launch_gemm<192>(tmA, tmA2, tmB, tmOut, tmR1, pk, stream, cg); break;
      case 256: rc = launch_gemm<256>(tmA, tmA2, tmB, tmOut, tmR1, pk, stream, cg); break;
      default: DD_CHECK(false, -1, "dd_gemm: unsupported tile width %d", bn);
    }
    if (rc) return rc;
    const int tile_m = BM * cg;
    p.m_tiles = (p.M + tile_m - 1) / tile_m;
    p.n_tiles = (p.N + bn - 1) / bn;
    const int iters = (p.taps == 9 ? 3 : 1) * ((p.K + BK - 1) / BK);
    splitk_reduce_kernel<<<dim3(p.m_tiles * p.n_tiles, tile_m / 32), 256, 0, stream>>>(p, bn, tile_m, iters);
    DD_CUDA(cudaGetLastError());
    count_launch(1);
    return 0;
  }
  switch (bn) {
    case 32: return launch_gemm<32>(tmA, tmA2, tmB, tmOut, tmR1, p, stream, cg);
    case 64: return launch_gemm<64>(tmA, tmA2, tmB, tmOut, tmR1, p, stream, cg);
    case 128: return launch_gemm<128>(tmA, tmA2, tmB, tmOut, tmR1, p, stream, cg);
    case 160: return launch_gemm<160>(tmA, tmA2, tmB, tmOut, tmR1, p, stream, cg);
    case 192: return launch_gemm<192>(tmA, tmA2, tmB, tmOut, tmR1, p, stream, cg);
    case 256: return launch_gemm<256>(tmA, tmA2, tmB, tmOut, tmR1, p, stream, cg);
    default: DD_CHECK(false, -1, "dd_gemm: unsupported tile width %d", bn);
  }
  return 0;
}

}  // namespace dd

#ifdef DD_GEMM_TRACE
// trace builds only (profiles/gemm_trace.py): copy out and clear the per-warp timelines (10 x 2048 events)
extern "C" __attribute__((visibility("default"))) int dd_gemm_trace_read(unsigned long long* dst) {
  if (cudaMemcpyFromSymbol(dst, dd::g_gemm_trace, sizeof(dd::g_gemm_trace)) != cudaSuccess) return -1;
  void* sym = nullptr;
  if (cudaGetSymbolAddress(&sym, dd::g_gemm_trace) != cudaSuccess) return -1;
  cudaMemset(sym, 0, sizeof(dd::g_gemm_trace));
  return 0;
}
#endif
