// dualdiff_b200 — flash-style fused attention on tcgen05 / TMEM / TMA (sm_100a).
//
// One entry point serves the four attentions of the DualDiff step:
//   self        (diffusers BasicTransformerBlock.attn1,  networks/blocks.py:166-172)
//   text-cross  (attn2, keys = [camera token | 77 text tokens | box tokens], blocks.py:182-188)
//   cross-view  (attn4, networks/blocks.py:190-222): each view attends to its two ring neighbours; the two
//               softmax-attentions are computed back to back and SUMMED in the epilogue, so the reference's
//               torch.cat of 12 (view, neighbour) pairs, the duplicated projections and the CPU-mask gather
//               disappear (out = A_left + A_right; to_out then adds 2*b_o, see packing.py)
//   SFA         (networks/txt_con_fusion.py:110-177): 320-ch condition feature map queries the 77 text tokens.
//
// S = Q K^T and O += P V run on the tensor cores with fp32 accumulators in TMEM; a softmax thread owns one query row
// (TMEM lane == row), so the online softmax needs no shuffles; P is written back to TMEM as bf16 and consumed as the
// TMEM A operand of the P V UMMAs; K/V tiles arrive by TMA through an mbarrier ring; V is consumed as an MN-major UMMA
// operand straight from its row-major [key, dv] layout (no transpose pass).
//
// Two kernels:
//   attn_pp_kernel  head_dim 40 (level 0: 1400 / 5600 tokens, 70 % of the attention time): TWO 128-row query tiles per
//                   CTA, one softmax warpgroup each, 48-key tiles, two CTAs per SM -> four softmax warps per scheduler.
//   attn_v2_kernel  head_dim 80 / 160 (levels 1-3): one query tile per CTA, 64-key tiles.
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
static constexpr int ATT_THREADS = 160;     // attn_v2: 4 softmax warps + 1 TMA/UMMA warp
static constexpr int PP_THREADS = 320;      // attn_pp: 2 softmax warpgroups (one query tile each) + 2 UMMA-issuing warps

struct AttnDev {
  int Lq, Lk, n_src, n_kv_tiles;
  const int* kv_map;  // [n_img * n_src] kv image per (query image, source) or nullptr (identity)
  float scale_log2e;
  bf16* out;
  long long out_ld;
  int q_col0, k_col0, v_col0, q_hs, k_hs, v_hs, o_hs;
  int heads, n_img;   // persistent kernel: item -> (image, head, query-tile pair)
  int concat;         // 1: the n_src K/V sources form ONE key sequence (a single softmax over all of them)
};

__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// Timeline instrumentation of attn_pp_kernel for profiles/attn_trace.py: only in a -DDD_ATTN_TRACE build (never in the shipped
// library).  Lane 0 of the first warp of each role of ONE CTA records (event, warp, tile, SM clock) pairs.
#ifdef DD_ATTN_TRACE
// per-warp regions (no atomics: an event is one clock read and two fire-and-forget stores)
__device__ unsigned long long g_attn_trace[10 * 2 * 4096];
#define DD_TR_DECL unsigned int tr_n_ = 0;
#define DD_TR(ev, tile)                                                                                     \
  do {                                                                                                      \
    if (blockIdx.x == 7 && (threadIdx.x & 31) == 0 && tr_n_ < 4096u) {                                      \
      unsigned long long* e_ = g_attn_trace + ((threadIdx.x >> 5) * 4096u + tr_n_) * 2u;                    \
      e_[0] = ((unsigned long long)(ev) << 48) | ((unsigned long long)(threadIdx.x >> 5) << 32) | (unsigned)(tile); \
      e_[1] = clock64();                                                                                    \
      ++tr_n_;                                                                                              \
    }                                                                                                       \
  } while (0)
#else
#define DD_TR_DECL
#define DD_TR(ev, tile) do { } while (0)
#endif

// 2^x for a pair of fp32 values on the FMA pipe (Cody-Waite split + degree-3 polynomial, max rel. error 7.5e-5 -- below the
// bf16 rounding of P).  x <= ~8 (lazy running max), clamped below at -126.
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

// normalise one O row held in TMEM and store (or, for the second source of the cross-view attention, add) it to global
// The tcgen05 fence that must separate these TMEM loads from the thread's next barrier arrival is issued right after the LAST
// load has landed, before the final stores: behind them it waited for the global stores as well (timeline of the text
// cross-attention, profiles/attn_trace.py 2 text: 1000-1700 cycles per item between the last store and the fence's return).
template <int DV, int DVP>
__device__ __forceinline__ void store_o_row(uint32_t tmem_o_row, float inv, bf16* orow, bool valid, bool add) {
#pragma unroll
  for (int c = 0; c < DVP; c += 16) {
    uint32_t ov[16];
    tmem_ld_32x16(tmem_o_row + c, ov);
    tmem_ld_wait();
    if (c + 16 >= DVP) tc_fence_before();
    if (valid) {
#pragma unroll
      for (int hh = 0; hh < 16; hh += 8) {
        if (c + hh + 8 <= DV) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(ov[hh + e]) * inv;
          if (add) {
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
}

// ---------------------------------------------------------------------------------------------------------------
// attn_pp_kernel: persistent CTAs, two query tiles per CTA.
//
// ncu on the one-tile kernel at head_dim 40 (profiles/r01_ncu_attn_L0_self.txt + source page): a 128 x 128 tile needs 16 384
// ex2 = 1024 cycles of the 16-lane MUFU pipe against 384 tensor cycles, so the kernel is MUFU-bound -- but the pipe was only
// 65 % busy.  The softmax warps spend 45 % of their time in the non-exponential part of the loop (TMEM load, row maximum,
// P store, barrier round trips), and with one softmax warp per CTA and scheduler only two warps share a MUFU pipe, often in
// the same phase.  Here:
//   * a work item = TWO 128-row query tiles of one (image, head): softmax warpgroup 0 / 1 (warps 0-3 / 4-7) each run the
//     online softmax of their own tile; with two CTAs per SM FOUR softmax warps share every scheduler, so the MUFU pipe
//     nearly always finds a warp in its exponential phase;
//   * every tile has its own UMMA-issuing warp (8 / 9; warp 8 also runs the TMA loads).  With one warp issuing for both tiles
//     (349 instructions per key tile, integer divisions and descriptor rebuilds included) the softmax warps spent 47 % of
//     their samples waiting for S (profiles/r02_attn_notes.md); now the (item, source, key tile) cursor is carried in
//     counters and the operand descriptors are formed once and advanced by adding to their low word;
//   * both tiles consume the same K/V stage: the K/V bytes crossing the L2 -> SM fabric halve (3.4 GB per level-0 launch
//     before, 5.7 TB/s -- the next limit once the MUFU pipe is busy);
//   * key tiles are 48 wide: S (48 columns) + P (24) + O (48) = 120 TMEM columns per query tile, 256 per CTA, and the
//     48 scores + 24 packed probabilities of a row fit the 96 registers two 320-thread CTAs leave per thread;
//   * the CTAs are PERSISTENT (grid = 2 x SMs): a CTA walks the items  blockIdx.x, blockIdx.x + gridDim.x, ...  as ONE stream
//     of K/V tiles through a six-stage TMA ring, so barrier set-up, TMEM allocation and the first K / V round trip are paid
//     once per CTA and the epilogue of an item overlaps the first key tiles of the next; the query tile of the next item is
//     reloaded while the softmax warpgroup works on the last key tile of the current one;
//   * POLY: one pair of every four is exponentiated on the FMA pipe (exp2_poly_pair) instead of the MUFU pipe, which then
//     has 25 % fewer operations; with four warps per scheduler the polynomial of one warp overlaps the MUFU stream of the
//     others (the all-or-nothing and same-warp forms measured slower in round 1).
//   * ONES: the V heads carry 1.0 in column DV (written by the bias of the value projection), so O[:, DV] = sum_k P[:, k]: the
//     softmax denominator is accumulated by the tensor core together with the numerator -- from the same bf16-rounded P --
//     and is rescaled with it; the softmax warps drop the packed row-sum adds (a sixth of the per-element instructions of
//     a loop that ncu shows issue-limited: 58 % issue slots busy, MUFU 54 %, profiles/r02_ncu_attn_L0_self.txt).
// TMEM columns of query tile t (base t * 128): S [0, BN) | P [BN, BN + BN/2) | O [BN + BN/2, BN + BN/2 + DVP).
// ---------------------------------------------------------------------------------------------------------------
template <int DQK, int DV, int DVP, int BN, int STAGES, int POLY, int ONES>
__global__ void __launch_bounds__(PP_THREADS, 2)
attn_pp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnDev p) {
  static_assert(DQK <= 64 && DVP <= 64 && DQK % 16 == 0 && DVP % 16 == 0, "one 64-column swizzle chunk per operand");
  static_assert(BN % 16 == 0 && BN >= 32 && BN <= 64 && STAGES >= 3 && STAGES <= 8, "key tile / ring geometry");
  static_assert(ONES == 0 || DV < DVP, "the column of ones lives in the padding of the V heads");
  constexpr int Q_TILE = ATT_BM * 128;                  // bytes: 128 query rows x 64 bf16
  constexpr int K_TILE = BN * 128;                      // bytes: BN key rows x 64 bf16 (a multiple of the 1024-byte swizzle atom)
  constexpr int KV_STAGE_BYTES = 2 * K_TILE;
  constexpr int T_STRIDE = 128;                         // TMEM columns per query tile
  constexpr int P_COL = BN, O_COL = BN + BN / 2;
  static_assert(O_COL + DVP <= T_STRIDE && K_TILE % 1024 == 0, "TMEM / swizzle geometry");
  constexpr int ISSUER = 8;                             // warps 8, 9: UMMA issuers of tile 0, 1 (warp 8 also runs the TMA loads)
  constexpr uint32_t IDESC_S = umma_idesc_bf16(ATT_BM, BN, 0, 0);
  constexpr uint32_t IDESC_O = umma_idesc_bf16(ATT_BM, DVP, 0, 1);  // B (=V) is MN-major

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t sQ = smem_base;                        // [2] query tiles
  const uint32_t sKV = sQ + 2 * Q_TILE;                 // [STAGES] {K tile, V tile}
  const uint32_t bar0 = sKV + STAGES * KV_STAGE_BYTES;
  const uint32_t q_full = bar0;                         // [2]  (per query tile)
  const uint32_t s_full = bar0 + 16;                    // [2]
  const uint32_t s_free = bar0 + 32;                    // [2]
  const uint32_t p_full = bar0 + 48;                    // [2]
  const uint32_t o_full = bar0 + 64;                    // [2]
  const uint32_t kv_full = bar0 + 80;                   // [STAGES <= 8]
  const uint32_t kv_empty = bar0 + 144;                 // [STAGES <= 8]
  uint32_t* tmem_ptr_gen = reinterpret_cast<uint32_t*>(smem_raw + (bar0 - smem_base) + 208);
  if ((smem_base & 1023u) != 0) __trap();

  const int warp = threadIdx.x >> 5;
  DD_TR_DECL
  const int n_q_tiles = (p.Lq + ATT_BM - 1) / ATT_BM;
  const int n_pairs = (n_q_tiles + 1) >> 1;
  const int n_items = n_pairs * p.heads * p.n_img;      // item = (image, head, pair of query tiles), pair fastest
  const int total = p.n_src * p.n_kv_tiles;             // key tiles per item
  const int stride = gridDim.x;

  if (threadIdx.x == 0) {
    for (int t = 0; t < 2; ++t) {
      mbar_init(q_full + 8 * t, 1);
      mbar_init(s_full + 8 * t, 1);
      mbar_init(s_free + 8 * t, 128);
      mbar_init(p_full + 8 * t, 128);
      mbar_init(o_full + 8 * t, 1);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(kv_full + 8 * s, 1);
      mbar_init(kv_empty + 8 * s, 2);       // two arrivals per key tile: one per issuing warp
    }
    fence_mbar_init();
  }
  if (warp == ISSUER) {
    tmem_alloc(bar0 + 208, 2 * T_STRIDE);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_gen);

  // tile t of an item exists unless the item is the last pair of an odd tile count and t == 1
  auto tile_active = [&](int item, int t) { return ((item % n_pairs) * 2 + t) < n_q_tiles; };

  if (warp >= ISSUER) {
    // ---------------- UMMA issuer of query tile t = warp - ISSUER (warp-uniform control flow, one elected lane issues) ----------------
    // The timeline of a CTA (profiles/attn_trace.py) showed the issuing warp, not the softmax warpgroups, setting the pace when
    // its loop carried a generic (item, source, tile) cursor: 207 instructions and ~2700 cycles per key tile (a ready
    // mbarrier.try_wait alone costs ~90 cycles, every tcgen05 / TMA instruction ~50) against ~1400 cycles of softmax work.
    // Hence the shape below: all item-level decisions (integer divisions, query-tile loads, single-tile items) sit in the outer
    // loop; the inner loop over the key tiles of one item only waits, issues and counts.
    const int t = warp - ISSUER;
    const bool producer = (t == 0);
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
    }
    __syncwarp();
    // ---- TMA producer (issuer 0): the K/V tiles of ALL items of this CTA in stream order, STAGES - 1 tiles ahead of the UMMAs
    int p_item = blockIdx.x, p_src = 0, p_jt = 0, p_row = 0, p_kcol = 0, p_vcol = 0, p_kvimg = 0;
    uint32_t p_s = 0, p_ph = 1;
    auto p_decode = [&]() {
      const int r = p_item / n_pairs;
      const int hd = r % p.heads, im = r / p.heads;
      p_kcol = p.k_col0 + hd * p.k_hs;
      p_vcol = p.v_col0 + hd * p.v_hs;
      p_kvimg = p.kv_map ? p.kv_map[im * p.n_src + p_src] : im;
    };
    if (producer) p_decode();
    auto produce = [&]() {
      if (p_item >= n_items) return;
      mbar_wait(kv_empty + 8 * p_s, p_ph);
      if (elect_one()) {
        const uint32_t sK = sKV + p_s * KV_STAGE_BYTES;
        mbar_arrive_expect_tx(kv_full + 8 * p_s, KV_STAGE_BYTES);
        tma_load_3d(sK, &tmK, kv_full + 8 * p_s, p_kcol, p_row, p_kvimg);
        tma_load_3d(sK + K_TILE, &tmV, kv_full + 8 * p_s, p_vcol, p_row, p_kvimg);
      }
      __syncwarp();
      if (++p_s == STAGES) { p_s = 0; p_ph ^= 1; }
      p_row += BN;
      if (++p_jt == p.n_kv_tiles) {          // next source / next item: rare
        p_jt = 0;
        p_row = 0;
        if (++p_src == p.n_src) { p_src = 0; p_item += stride; }
        if (p_item < n_items) p_decode();
      }
    };
    // ---- query tile t of an item (one buffer per tile: reloaded when the last Q K^T of the previous item has retired)
    auto load_q = [&](int item) {
      const int r = item / n_pairs;
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full + 8 * t, Q_TILE);
        tma_load_3d(sQ + t * Q_TILE, &tmQ, q_full + 8 * t, p.q_col0 + (r % p.heads) * p.q_hs,
                    ((item % n_pairs) * 2 + t) * ATT_BM, r / p.heads);
      }
      __syncwarp();
    };
    auto next_active_item = [&](int item) {
      item += stride;
      while (item < n_items && !tile_active(item, t)) item += stride;
      return item;
    };
    int item = blockIdx.x;                                   // gridDim.x <= n_items: every CTA has a first item
    {
      const int first = tile_active(item, t) ? item : next_active_item(item);
      if (first < n_items) load_q(first);
    }
    if (producer) {
#pragma unroll 1
      for (int i = 0; i < STAGES - 1; ++i) produce();
    }
    const uint32_t tS = tmem_base + t * T_STRIDE, tP = tS + P_COL, tO = tS + O_COL;
    const uint64_t dQ = umma_smem_desc(sQ + t * Q_TILE, 16, 1024, 2);       // + 2 per 16-column K step
    const uint64_t dK0 = umma_smem_desc(sKV, 16, 1024, 2);                  // + 2 per K step, + STAGE16 per stage
    const uint64_t dV0 = umma_smem_desc(sKV + K_TILE, K_TILE, 1024, 2);     // + 128 per 16 keys, + STAGE16 per stage
    constexpr uint32_t STAGE16 = KV_STAGE_BYTES >> 4;
    const uint32_t b_sfull = s_full + 8 * t, b_sfree = s_free + 8 * t, b_pfull = p_full + 8 * t, b_ofull = o_full + 8 * t;
    uint32_t s_cur = 0, ph_cur = 0;   // ring stage / phase of the key tile whose P V comes next
    uint32_t k = 0;                   // key tiles issued for this query-tile slot (parity of the per-tile barriers)
    uint32_t n_q = 0;                 // query tiles loaded so far by this issuer (parity of q_full)
    // S_t = Q_t K(stage s)^T
    auto issue_qk = [&](uint32_t s, uint32_t ph) {
      mbar_wait(kv_full + 8 * s, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dK = dK0 + (uint64_t)(s * STAGE16);
#pragma unroll
        for (int kk = 0; kk < DQK / 16; ++kk) umma_bf16(tS, dQ + 2 * kk, dK + 2 * kk, IDESC_S, kk != 0 ? 1u : 0u);
        umma_commit(b_sfull);
      }
      __syncwarp();
    };
    // O_t (+)= P_t V(stage s): A = P from TMEM (8 packed columns per 16 keys), B = V (MN-major)
    auto issue_pv = [&](uint32_t s, bool accumulate) {
      mbar_wait(b_pfull, k & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dV = dV0 + (uint64_t)(s * STAGE16);
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) umma_bf16_ts(tO, tP + kk * 8, dV + 128 * kk, IDESC_O, (kk != 0 || accumulate) ? 1u : 0u);
        umma_commit(b_ofull);
        umma_commit(kv_empty + 8 * s);                     // the stage is free once the UMMAs of both issuers retired
      }
      __syncwarp();
    };
#pragma unroll 1
    for (; item < n_items; item += stride) {
      if (!tile_active(item, t)) {
        // tile t of this item does not exist (last pair of an odd tile count): walk its K/V tiles anyway so that this issuer
        // consumes the phases of the kv ring in order (a parity wait is only meaningful within one phase), releasing each
#pragma unroll 1
        for (int i = 0; i < total; ++i) {
          mbar_wait(kv_full + 8 * s_cur, ph_cur);
          if (elect_one()) mbar_arrive(kv_empty + 8 * s_cur);
          __syncwarp();
          if (++s_cur == STAGES) { s_cur = 0; ph_cur ^= 1; }
        }
        continue;
      }
      // first key tile of the item: its Q K^T waits for the query tile (loaded while the previous item was finishing)
      mbar_wait(q_full + 8 * t, n_q & 1);
      ++n_q;
      if (k > 0) mbar_wait(b_sfree, (k - 1) & 1);          // S_t of the previous item's last tile is in registers
      tc_fence_after();
      issue_qk(s_cur, ph_cur);
      int jt = 0;
#pragma unroll 1
      for (int i = 0; i < total; ++i) {
        DD_TR(10, k);
        uint32_t s_nxt = s_cur + 1, ph_nxt = ph_cur;
        if (s_nxt == STAGES) { s_nxt = 0; ph_nxt ^= 1; }
        if (i + 1 < total) {
          // S_t(next tile) is issued as soon as the softmax warpgroup holds S_t(this tile) in registers
          mbar_wait(b_sfree, k & 1);
          tc_fence_after();
          DD_TR(11, k);
          issue_qk(s_nxt, ph_nxt);
          DD_TR(12, k);
        } else {
          // last key tile of the item: once its S is in registers every Q K^T of the item has retired -> reload the query tile
          const int nxt = next_active_item(item);
          if (nxt < n_items) {
            mbar_wait(b_sfree, k & 1);
            load_q(nxt);
          }
        }
        // refill the ring while the softmax warpgroup is still exponentiating this tile
        if (producer) produce();
        DD_TR(15, k);
        issue_pv(s_cur, p.concat ? (i != 0) : (jt != 0));
        DD_TR(16, k);
        if (++jt == p.n_kv_tiles) jt = 0;
        s_cur = s_nxt;
        ph_cur = ph_nxt;
        ++k;
      }
    }
  } else {
    // ------------------------------- softmax / correction / epilogue of query tile t -------------------------------
    const int t = warp >> 2;
    const int row = threadIdx.x & 127;            // == TMEM lane
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tmem_S = tmem_base + t * T_STRIDE + lane_sel;
    const uint32_t tmem_P = tmem_S + P_COL;
    const uint32_t tmem_O = tmem_S + O_COL;
    const uint32_t bs_full = s_full + 8 * t, bs_free = s_free + 8 * t, bp_full = p_full + 8 * t, bo_full = o_full + 8 * t;
    const float sl2 = p.scale_log2e;
    const uint64_t SL2 = pack_f32x2(sl2, sl2);
    uint32_t k = 0;                               // key tiles processed by this warpgroup (parity of the per-tile barriers)
    bool s_ready = false;                         // s_full of tile k already seen complete (probed one tile ahead)
    int item = blockIdx.x;
    while (item < n_items && !tile_active(item, t)) item += stride;
#pragma unroll 1
    for (; item < n_items;) {
      const int r = item / n_pairs;
      const int q_row = ((item % n_pairs) * 2 + t) * ATT_BM + row;
      bf16* orow = p.out + ((long long)(r / p.heads) * p.Lq + q_row) * p.out_ld + (r % p.heads) * p.o_hs;
      float m = -INFINITY, l = 0.f;
      for (int src = 0; src < p.n_src; ++src) {
        if (!p.concat) { m = -INFINITY; l = 0.f; }          // one softmax per source, or one over the concatenated sources
        const bool o_live = p.concat && src > 0;           // O already holds the earlier sources of this softmax
#pragma unroll 1
        for (int jt = 0; jt < p.n_kv_tiles; ++jt, ++k) {
          DD_TR(0, k);
          if (!s_ready) mbar_wait(bs_full, k & 1);   // usually probed already while the previous tile's P store was in flight
          tc_fence_after();
          DD_TR(1, k);
          uint32_t sv[BN];
          if constexpr (BN == 48) {
            tmem_ld_32x32(tmem_S, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
            tmem_ld_32x16(tmem_S + 32, *reinterpret_cast<uint32_t(*)[16]>(&sv[32]));
          } else {
#pragma unroll
            for (int c = 0; c < BN; c += 32) tmem_ld_32x32(tmem_S + c, *reinterpret_cast<uint32_t(*)[32]>(&sv[c]));
          }
          // the previous tile's P V must have retired before P is overwritten / O is rescaled: it was issued a whole tile ago, so
          // this wait only costs the latency of the barrier test -- which hides behind the TMEM load in flight
          if (k > 0) {
            mbar_wait(bo_full, (k - 1) & 1);
            tc_fence_after();
          }
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(bs_free);                      // S_t now lives in registers -> the issuer starts the next Q K^T
          DD_TR(2, k);
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
          // p = exp2(s * scale * log2e - m) packed to bf16; row sum on four packed accumulators
          const uint64_t NEGM = pack_f32x2(-m, -m);
          uint64_t acc0 = 0ull, acc1 = 0ull, acc2 = 0ull, acc3 = 0ull;
          uint32_t pk[BN / 2];
#pragma unroll
          for (int jp = 0; jp < BN / 2; ++jp) {
            const uint64_t X = fma_f32x2(pack_f32x2(__uint_as_float(sv[2 * jp]), __uint_as_float(sv[2 * jp + 1])), SL2, NEGM);
            float x0, x1, p0, p1;
            unpack_f32x2(X, x0, x1);
            if (POLY != 0 && (jp & 3) == 3) {
              exp2_poly_pair(x0, x1, p0, p1);
            } else {
              p0 = fast_exp2(x0);
              p1 = fast_exp2(x1);
            }
            if constexpr (ONES == 0) {
              const uint64_t PP = pack_f32x2(p0, p1);
              const int u = jp & 3;
              if (u == 0) acc0 = add_f32x2(acc0, PP);
              if (u == 1) acc1 = add_f32x2(acc1, PP);
              if (u == 2) acc2 = add_f32x2(acc2, PP);
              if (u == 3) acc3 = add_f32x2(acc3, PP);
            }
            pk[jp] = pack_bf16(p0, p1);
          }
          if constexpr (ONES == 0) {
            float s0, s1;
            unpack_f32x2(add_f32x2(add_f32x2(acc0, acc1), add_f32x2(acc2, acc3)), s0, s1);
            l += s0 + s1;
          }
          DD_TR(3, k);
          DD_TR(4, k);
          if (grow && (jt > 0 || o_live)) {
#pragma unroll
            for (int c = 0; c < DVP; c += 16) {
              uint32_t ov[16];
              tmem_ld_32x16(tmem_O + c, ov);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) ov[j] = __float_as_uint(__uint_as_float(ov[j]) * alpha);
              tmem_st_32x16(tmem_O + c, ov);
            }
          }
          if constexpr (BN == 48) {
            tmem_st_32x16(tmem_P, *reinterpret_cast<const uint32_t(*)[16]>(&pk[0]));
            tmem_st_32x8(tmem_P + 16, *reinterpret_cast<const uint32_t(*)[8]>(&pk[16]));
          } else if constexpr (BN == 64) {
            tmem_st_32x32(tmem_P, pk);
          } else {
            tmem_st_32x16(tmem_P, *reinterpret_cast<const uint32_t(*)[16]>(&pk[0]));
          }
          s_ready = mbar_test_wait(bs_full, (k + 1) & 1);   // S of the next tile (issued when this one reached the registers)
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(bp_full);
          DD_TR(5, k);
        }
        if (p.concat && src + 1 < p.n_src) continue;        // concatenated sources: one epilogue after the last of them
        // epilogue of this source: O / l  (second source of the cross-view attention adds onto the first).  The issuer is
        // already feeding the next source / item: its first P V cannot start before this warpgroup's next p_full arrival.
        mbar_wait(bo_full, (k - 1) & 1);
        tc_fence_after();
        if constexpr (ONES != 0) {       // the denominator sits in column DV of the row's O accumulator
          uint32_t t16[16];
          tmem_ld_32x16(tmem_O + (DV / 16) * 16, t16);
          tmem_ld_wait();
          l = __uint_as_float(t16[DV % 16]);
        }
        store_o_row<DV, DVP>(tmem_O, 1.f / l, orow, q_row < p.Lq, src > 0 && !p.concat);   // (ends with the tcgen05 fence)
        DD_TR(6, k);
      }
      item += stride;
      while (item < n_items && !tile_active(item, t)) item += stride;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == ISSUER) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * T_STRIDE);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// attn_v2_kernel: one query tile per work item (head_dim 80 / 160).  Key tiles are 64 wide so that S + O + P fit 256 TMEM
// columns and two CTAs share an SM; the softmax loop reads the whole S tile, reduces its maximum on four independent
// FMNMX3 chains, and runs scale/subtract and the row sum as packed FFMA2 / FADD2.
// The CTAs walk the items  blockIdx.x, blockIdx.x + gridDim.x, ...  (item = (image, head, query tile), tile fastest) as ONE
// stream of K/V tiles through the TMA ring: levels 1-3 have 1-6 key tiles per item (350 / 91 / 28 tokens), so with one CTA
// per item the barrier set-up, the TMEM allocation and the first Q / K / V round trip were most of a CTA's life (ncu: 102 us
// per level-1 self-attention launch against a 24 us MUFU bound).  Run persistently they are paid once per CTA, the K/V tiles
// of the next item stream in during the epilogue of the current one, and its query tile is reloaded while the softmax
// warps work on the last key tile.  With gridDim.x == number of items the same code is the one-item-per-CTA kernel.
// Measured dead ends (profiles/README.md): two softmax warpgroups splitting the columns of ONE tile, chunk-wise software
// pipelining of the TMEM loads (the former DD_ATTN_PIPE variant), staggering the two CTAs of an SM, ex2.approx.f16x2.
// ---------------------------------------------------------------------------------------------------------------
template <int DQK, int DV, int DVP, int BN, int STAGES, int MINB>
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
  const int n_q_tiles = (p.Lq + ATT_BM - 1) / ATT_BM;
  const int n_items = n_q_tiles * p.heads * p.n_img;    // item = (image, head, query tile), tile fastest
  const int stride = gridDim.x;
  const int total = p.n_src * p.n_kv_tiles;             // key tiles per item

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
  if (warp == ISSUER) {
    // ---------------- TMA producer + UMMA issuer (warp-uniform control flow, one elected lane issues) ----------------
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
    }
    __syncwarp();
    auto load_q = [&](int item) {
      const int r = item / n_q_tiles;
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, QCH * Q_CHUNK);
        for (int c = 0; c < QCH; ++c)
          tma_load_3d(sQ + c * Q_CHUNK, &tmQ, q_full, p.q_col0 + (r % p.heads) * p.q_hs + c * 64, (item % n_q_tiles) * ATT_BM,
                      r / p.heads);
      }
      __syncwarp();
    };
    // K/V tiles of ALL items of this CTA in stream order; gp counts the tiles produced (ring stage / phase)
    int p_item = blockIdx.x, p_t = 0, gp = 0;
    auto produce = [&]() {
      if (p_item >= n_items) return;
      const int s = gp % STAGES;
      mbar_wait(kv_empty + 8 * s, ((gp / STAGES) & 1) ^ 1);
      const int r = p_item / n_q_tiles;
      const int head = r % p.heads, img = r / p.heads;
      const int src = p_t / p.n_kv_tiles, jt = p_t - src * p.n_kv_tiles;
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
      ++gp;
      if (++p_t == total) { p_t = 0; p_item += stride; }
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
    load_q(blockIdx.x);                                  // gridDim.x <= n_items: every CTA has a first item
#pragma unroll 1
    for (int i = 0; i < STAGES; ++i) produce();
    int g = 0;                                           // key tiles issued by this CTA (parity of the per-tile barriers)
    uint32_t n_q = 0;                                    // query tiles loaded so far (parity of q_full)
#pragma unroll 1
    for (int item = blockIdx.x; item < n_items; item += stride) {
      mbar_wait(q_full, n_q & 1);
      ++n_q;
      if (g > 0) {                                       // S of the previous item's last tile is in registers
        mbar_wait(s_free, (g - 1) & 1);
        tc_fence_after();
      }
      issue_qk(g);
      int jt = 0;
#pragma unroll 1
      for (int i = 0; i < total; ++i, ++g) {
        const int s = g % STAGES;
        const bool last = (i + 1 == total);
        // S(g+1) = Q K(g+1)^T is issued as soon as the softmax threads hold S(g) in registers; after the last key tile of
        // the item every Q K^T has retired by then, and the query tile of the next item is loaded instead
        if (!last) {
          if (STAGES > 1) {
            mbar_wait(s_free, g & 1);
            tc_fence_after();
            issue_qk(g + 1);
          }
        } else if (item + stride < n_items) {
          mbar_wait(s_free, g & 1);
          load_q(item + stride);
        }
        mbar_wait(p_full, g & 1);
        tc_fence_after();
        const uint32_t sV = sKV + s * KV_STAGE_BYTES + QCH * K_CHUNK;
        // O += P V : A = P from TMEM (8 packed columns per 16 keys), B = V (MN-major: rows = keys, 64-wide dv chunks)
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk) {
            umma_bf16_ts(tmem_O, tmem_P + kk * 8, umma_smem_desc(sV + kk * 16 * 128, K_CHUNK, 1024, 2), IDESC_O,
                         (kk != 0 || (p.concat ? i != 0 : jt != 0)) ? 1u : 0u);  // O accumulates over one source (all, concatenated)
          }
          umma_commit(kv_empty + 8 * s);
          umma_commit(o_full);
        }
        __syncwarp();
        if (++jt == p.n_kv_tiles) jt = 0;
        produce();                                     // stage s is free once P(g) V(g) retires (kv_empty)
        if (STAGES == 1 && !last) {                    // single stage: K(g+1) only lands after P(g) V(g) has retired
          mbar_wait(s_free, g & 1);
          tc_fence_after();
          issue_qk(g + 1);
        }
      }
    }
  } else {
    // ------------------------------- softmax / correction / epilogue -------------------------------
    // Thread = one query row (TMEM lane): the online softmax needs no shuffles.  Per 128x128 tile and SM it costs ~1024
    // cycles of the 16-lane MUFU pipe and ~730 cycles of the TMEM read port (tcgen05.ld moves ~90 B/clk/SM,
    // profiles/micro/tmem_bench.cu; the S tile is 64 KB) — both shared by the two CTAs of an SM.
    const int row = threadIdx.x;                // == TMEM lane
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    const float sl2 = p.scale_log2e;
    const uint64_t SL2 = pack_f32x2(sl2, sl2);
    constexpr int NC = BN / 2;                  // S columns per pass (P is published in two passes)
    int g = 0;
    bool s_ready = false;                       // s_full of tile g already seen complete (probed one tile ahead)
#pragma unroll 1
    for (int item = blockIdx.x; item < n_items; item += stride) {
      const int r = item / n_q_tiles;
      const int q_row = (item % n_q_tiles) * ATT_BM + row;
      bf16* orow = p.out + ((long long)(r / p.heads) * p.Lq + q_row) * p.out_ld + (r % p.heads) * p.o_hs;
    float m = -INFINITY, l = 0.f;
    for (int src = 0; src < p.n_src; ++src) {
      if (!p.concat) { m = -INFINITY; l = 0.f; }            // one softmax per source, or one over the concatenated sources
      const bool o_live = p.concat && src > 0;
#pragma unroll 1
      for (int jt = 0; jt < p.n_kv_tiles; ++jt, ++g) {
        if (!s_ready) mbar_wait(s_full, g & 1);    // usually probed already while the previous tile's P store was in flight
        tc_fence_after();
        uint32_t sv[BN];
#pragma unroll
        for (int c = 0; c < BN; c += 32) tmem_ld_32x32(tmem_S + lane_sel + c, *reinterpret_cast<uint32_t(*)[32]>(&sv[c]));
        // the previous tile's P V must have retired before P is overwritten / O is rescaled: it was issued a whole tile ago,
        // so this wait costs only the latency of the barrier test, hidden behind the TMEM load in flight
        if (g > 0) {
          mbar_wait(o_full, (g - 1) & 1);
          tc_fence_after();
        }
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
            p0 = fast_exp2(x0);
            p1 = fast_exp2(x1);
            const uint64_t PP = pack_f32x2(p0, p1);
            const int u = (jj >> 1) & 3;
            if (u == 0) acc0 = add_f32x2(acc0, PP);
            if (u == 1) acc1 = add_f32x2(acc1, PP);
            if (u == 2) acc2 = add_f32x2(acc2, PP);
            if (u == 3) acc3 = add_f32x2(acc3, PP);
            pk[(jj >> 1) - pass * (NC / 2)] = pack_bf16(p0, p1);
          }
          if (pass == 0) {
            if (grow && (jt > 0 || o_live)) {
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
        s_ready = mbar_test_wait(s_full, (g + 1) & 1);   // S of the next tile (issued when this one reached the registers)
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full);
      }
      if (p.concat && src + 1 < p.n_src) continue;          // concatenated sources: one epilogue after the last of them
      // epilogue of this source: O / l  (second source of the cross-view attention adds onto the first).  The issuer is
      // already loading the next source / item; its first P V waits for this thread's next p_full arrival.
      mbar_wait(o_full, (g - 1) & 1);
      tc_fence_after();
      store_o_row<DV, DVP>(tmem_O + lane_sel, 1.f / l, orow, q_row < p.Lq, src > 0 && !p.concat);   // (ends with the tcgen05 fence)
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == ISSUER) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


template <int DQK, int DV, int DVP, int BN, int STAGES, int MINB>
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
  if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(attn_v2_kernel<DQK, DV, DVP, BN, STAGES, MINB>), (int)smem)) return e;
  const long long n_items = (long long)((a->lq + ATT_BM - 1) / ATT_BM) * a->heads * a->n_img;
  DD_CHECK(n_items < (1ll << 30), -1, "dd_attention: too many work items");
  // Short K/V streams (levels 1-3: at most 12 key tiles per item) run persistently on MINB CTAs per SM, every CTA taking the
  // same number of items (+-1); long streams launch one CTA per item and leave the balancing to the hardware scheduler.
  const long long slots = (long long)MINB * num_sms();
  long long ctas = n_items;
  if (a->variant != 1 && (long long)a->n_src * p.n_kv_tiles <= 16 && n_items > slots) {
    const long long rounds = (n_items + slots - 1) / slots;
    ctas = (n_items + rounds - 1) / rounds;
  }
  dim3 grid((unsigned)ctas);
  attn_v2_kernel<DQK, DV, DVP, BN, STAGES, MINB><<<grid, ATT_THREADS, smem, stream>>>(tmQ, tmK, tmV, p);
  DD_CUDA(cudaGetLastError());
  return 0;
}

template <int DQK, int DV, int DVP, int BN, int STAGES, int POLY, int ONES>
static int launch_attn_pp(const dd_attention_args* a, AttnDev p, cudaStream_t stream) {
  constexpr size_t smem = (size_t)2 * ATT_BM * 128 + (size_t)STAGES * 2 * BN * 128 + 256;
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
  if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(attn_pp_kernel<DQK, DV, DVP, BN, STAGES, POLY, ONES>), (int)smem)) return e;
  const int n_q_tiles = (a->lq + ATT_BM - 1) / ATT_BM;
  const long long n_items = (long long)((n_q_tiles + 1) / 2) * a->heads * a->n_img;
  DD_CHECK(n_items < (1ll << 30), -1, "dd_attention: too many work items");
  // Short K/V streams (text / SFA cross-attention: 2-3 key tiles per item) run PERSISTENT, two CTAs per SM, so that barrier
  // set-up, TMEM allocation and the first Q / K / V round trip are paid once per CTA (170 -> 124 us per level-0 text launch).
  // Long streams (self / cross-view: 30-120 key tiles per item) launch one CTA per item: the hardware's dynamic CTA scheduling
  // balances the SMs better than a static round-robin (measured 539 vs 565 us per level-0 self-attention launch).
  const int slots = 2 * num_sms();
  const bool persistent = a->n_src * p.n_kv_tiles <= 8;
  dim3 grid((unsigned)((n_items < slots || !persistent) ? n_items : slots));
  attn_pp_kernel<DQK, DV, DVP, BN, STAGES, POLY, ONES><<<grid, PP_THREADS, smem, stream>>>(tmQ, tmK, tmV, p);
  DD_CHECK(cudaGetLastError() == cudaSuccess, -2, "dd_attention: launch failed");
  return 0;
}

int attention_run(const dd_attention_args* a, cudaStream_t stream) {
  DD_CHECK(a != nullptr, -1, "dd_attention: null args");
  DD_CHECK(a->n_img > 0 && a->heads > 0 && a->lq > 0 && a->lk > 0, -1, "dd_attention: bad shape");
  DD_CHECK(a->head_dim == 40 || a->head_dim == 80 || a->head_dim == 160, -1,
           "dd_attention: head_dim %d unsupported (40, 80, 160 = SDv1.5 levels)", a->head_dim);
  DD_CHECK(a->n_src >= 1 && a->n_src <= 8, -1, "dd_attention: n_src must be 1..8");
  DD_CHECK(a->n_src == 1 || a->kv_map != nullptr, -1, "dd_attention: kv_map required for n_src > 1");
  DD_CHECK(a->q_ld % 8 == 0 && a->k_ld % 8 == 0 && a->v_ld % 8 == 0 && a->out_ld % 8 == 0, -1,
           "dd_attention: leading dims must be multiples of 8");
  DD_CHECK(a->n_kv_img > 0, -1, "dd_attention: n_kv_img missing");
  DD_CHECK(a->variant >= 0 && a->variant <= 2, -1, "dd_attention: variant must be 0 (auto), 1 or 2");
  DD_CHECK(a->v_ones == 0 || a->head_dim == 40, -1, "dd_attention: v_ones is a head_dim-40 layout");
  AttnDev p;
  p.Lq = a->lq; p.Lk = a->lk; p.n_src = a->n_src; p.n_kv_tiles = 0;
  p.kv_map = a->kv_map;
  p.scale_log2e = a->scale * 1.4426950408889634f;
  p.out = reinterpret_cast<bf16*>(a->out); p.out_ld = a->out_ld;
  p.q_col0 = a->q_col0; p.k_col0 = a->k_col0; p.v_col0 = a->v_col0;
  p.q_hs = a->q_head_stride; p.k_hs = a->k_head_stride; p.v_hs = a->v_head_stride; p.o_hs = a->head_dim;
  p.heads = a->heads; p.n_img = a->n_img;
  p.concat = a->concat ? 1 : 0;
  switch (a->head_dim) {
    case 40:
      // The 48-wide Q.K^T reads 8 columns beyond a 40-wide head.  Either side may keep unpadded (stride 40) heads as long as the
      // OTHER side is zero-padded to 48: the extra columns then meet zeros (txt_con_XFormersAttn_plus feeds the output of one
      // attention -- 8 heads of 40 columns -- as the queries of the next).
      DD_CHECK(a->q_head_stride >= 40 && a->k_head_stride >= 40 && (a->q_head_stride >= 48 || a->k_head_stride >= 48) &&
                   (a->q_head_stride == 40 || a->q_head_stride >= 48) && (a->k_head_stride == 40 || a->k_head_stride >= 48),
               -1, "dd_attention: head_dim 40 needs the Q or the K heads zero-padded to a 48-column stride");
      DD_CHECK(a->v_ones == 0 || a->v_head_stride >= 48, -1, "dd_attention: v_ones needs V heads on a 48-column stride");
      if (a->variant == 1) return launch_attn_v2<48, 40, 48, 128, 3, 2>(a, p, stream);   // testing hook: one-tile kernel
      if (a->variant == 2)   // testing hook: all ex2 on MUFU
        return a->v_ones ? launch_attn_pp<48, 40, 48, 48, 6, 0, 1>(a, p, stream) : launch_attn_pp<48, 40, 48, 48, 6, 0, 0>(a, p, stream);
      // short K/V streams (text / SFA: <= 8 key tiles, persistent CTAs) run faster with every exponential on MUFU
      // (105 vs 112 us per level-0 text launch); long ones with a quarter of them on the FMA pipe (509 vs 532 us)
      if (a->v_ones && a->n_src * ((a->lk + 47) / 48) <= 8) return launch_attn_pp<48, 40, 48, 48, 6, 0, 1>(a, p, stream);
      if (a->v_ones) return launch_attn_pp<48, 40, 48, 48, 6, 1, 1>(a, p, stream);
      return launch_attn_pp<48, 40, 48, 48, 6, 1, 0>(a, p, stream);
    case 80: return launch_attn_v2<80, 80, 80, 64, 2, 2>(a, p, stream);
    case 160: return launch_attn_v2<160, 160, 160, 64, 1, 2>(a, p, stream);
  }
  return -1;
}

}  // namespace dd

#ifdef DD_ATTN_TRACE
// trace builds only (profiles/attn_trace.py): copy out and clear the per-warp timelines of attn_pp_kernel (10 x 4096 events)
extern "C" __attribute__((visibility("default"))) int dd_attn_trace_read(unsigned long long* dst) {
  if (cudaMemcpyFromSymbol(dst, dd::g_attn_trace, sizeof(dd::g_attn_trace)) != cudaSuccess) return -1;
  void* sym = nullptr;
  if (cudaGetSymbolAddress(&sym, dd::g_attn_trace) != cudaSuccess) return -1;
  cudaMemset(sym, 0, sizeof(dd::g_attn_trace));
  return 10 * 4096;
}
#endif
