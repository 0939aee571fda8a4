// dualdiff_b200 — temporal attention over the frames of a video clip (BASELINE config 5, "DualDiff+": 16 frames x 6 views).
//
// The reference repository ships no temporal block; the block is DEFINED here as
//     x_f += W_o * MHA(q = LN(x_f), k = v = LN(x_f') for every frame f' of the clip) + b_o       (W_o, b_o zero-initialised)
// applied per (scene, camera view, spatial token) after the cross-view attention of BasicMultiviewTransformerBlock,
// bidirectional over frames, 8 heads of C/8 channels, scale (C/8)^-0.5 (oracle/dualdiff_oracle.py:temporal_attention).
//
// A sequence is only F <= 32 frames long, so this is HBM-bound gather work, not tensor-core work: a CTA stages the K and V
// rows of TOK tokens x all frames in shared memory with coalesced 16-byte loads (one frame's row of a token is C contiguous
// bf16), then one thread per (token, head, query frame) runs the F x F attention in registers.  With the frames of a clip
// sharded over ranks, K/V of the other ranks arrive by an NCCL all-gather (dualdiff_b200/sharding.py:FrameShard) and are
// addressed in place through (rank, local frame) strides.
#include "dd_api_internal.h"
#include "dd_common.cuh"

namespace dd {

static constexpr int TMP_MAX_F = 32;

struct TemporalDev {
  const bf16* q; const bf16* k; const bf16* v; bf16* out;
  long long q_ld, k_ld, v_ld, out_ld;
  int q_col0, k_col0, v_col0, q_hs, k_hs, v_hs;
  int n_outer, n_view, T, heads;
  int fq;                 // query frames
  int fq_per_rank;        // ... per rank block of the query / output buffers (= fq when they are one block)
  long long q_rank_stride;   // images between two rank blocks of the query / output buffers
  int fkv, fkv_per_rank;  // key/value frames in total, and per gathered rank block
  long long kv_rank_stride;  // images between two rank blocks of the gathered K/V buffer
  float scale_log2e;
  int tok;                // tokens per CTA
};

template <int D>
__global__ void __launch_bounds__(256)
temporal_attn_kernel(const TemporalDev p) {
  extern __shared__ __align__(16) uint8_t tsm[];   // [tok][fkv][2][heads][D] bf16
  constexpr int VPH = D / 8;                        // 16-byte vectors per head
  const long long n_seq = (long long)p.n_outer * p.n_view * p.T;
  const long long seq0 = (long long)blockIdx.x * p.tok;
  const int C = p.heads * D;
  // ---- stage K and V of every frame for this CTA's tokens ----
  const int vec_per_tok = p.fkv * 2 * p.heads * VPH;
  for (int i = threadIdx.x; i < p.tok * vec_per_tok; i += blockDim.x) {
    const int tl = i / vec_per_tok;
    int r = i - tl * vec_per_tok;
    const long long seq = seq0 + tl;
    if (seq >= n_seq) continue;
    const int f = r / (2 * p.heads * VPH);
    r -= f * 2 * p.heads * VPH;
    const int which = r / (p.heads * VPH);
    r -= which * p.heads * VPH;
    const int h = r / VPH, vec = r - h * VPH;
    const int t = (int)(seq % p.T);
    const long long ov = seq / p.T;
    const int v = (int)(ov % p.n_view);
    const long long o = ov / p.n_view;
    const int rk = f / p.fkv_per_rank, fl = f - rk * p.fkv_per_rank;
    const long long img = rk * p.kv_rank_stride + (o * p.fkv_per_rank + fl) * p.n_view + v;
    const bf16* src = which ? p.v + (img * p.T + t) * p.v_ld + p.v_col0 + h * p.v_hs + vec * 8
                            : p.k + (img * p.T + t) * p.k_ld + p.k_col0 + h * p.k_hs + vec * 8;
    *reinterpret_cast<uint4*>(tsm + ((size_t)i << 4)) = *reinterpret_cast<const uint4*>(src);
  }
  __syncthreads();
  // ---- one thread per (token, head, query frame) ----
  const int per_tok = p.heads * p.fq;
  const int tl = threadIdx.x / per_tok;
  if (tl >= p.tok) return;
  const long long seq = seq0 + tl;
  if (seq >= n_seq) return;
  const int r = threadIdx.x - tl * per_tok;
  const int h = r / p.fq, fi = r - h * p.fq;
  const int t = (int)(seq % p.T);
  const long long ov = seq / p.T;
  const int v = (int)(ov % p.n_view);
  const long long o = ov / p.n_view;
  const int rkq = fi / p.fq_per_rank, flq = fi - rkq * p.fq_per_rank;
  const long long qimg = rkq * p.q_rank_stride + (o * p.fq_per_rank + flq) * p.n_view + v;
  const bf16* qp = p.q + (qimg * p.T + t) * p.q_ld + p.q_col0 + h * p.q_hs;
  const uint8_t* kbase = tsm + (size_t)tl * vec_per_tok * 16;
  auto kv_vec = [&](int f, int which, int vec) {
    return *reinterpret_cast<const uint4*>(kbase + ((size_t)((f * 2 + which) * p.heads + h) * VPH + vec) * 16);
  };
  float s[TMP_MAX_F];
#pragma unroll
  for (int j = 0; j < TMP_MAX_F; ++j) s[j] = 0.f;
#pragma unroll
  for (int vec = 0; vec < VPH; ++vec) {
    const uint4 qv = *reinterpret_cast<const uint4*>(qp + vec * 8);
    const float2 q0 = unpack_bf16(qv.x), q1 = unpack_bf16(qv.y), q2 = unpack_bf16(qv.z), q3 = unpack_bf16(qv.w);
#pragma unroll
    for (int j = 0; j < TMP_MAX_F; ++j) {
      if (j < p.fkv) {
        const uint4 kk = kv_vec(j, 0, vec);
        const float2 k0 = unpack_bf16(kk.x), k1 = unpack_bf16(kk.y), k2 = unpack_bf16(kk.z), k3 = unpack_bf16(kk.w);
        s[j] += q0.x * k0.x + q0.y * k0.y + q1.x * k1.x + q1.y * k1.y + q2.x * k2.x + q2.y * k2.y + q3.x * k3.x + q3.y * k3.y;
      }
    }
  }
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < TMP_MAX_F; ++j)
    if (j < p.fkv) m = fmaxf(m, s[j]);
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < TMP_MAX_F; ++j) {
    if (j < p.fkv) {
      s[j] = exp2f((s[j] - m) * p.scale_log2e);
      l += s[j];
    }
  }
  const float inv = 1.f / l;
  bf16* op = p.out + (qimg * p.T + t) * p.out_ld + h * D;
#pragma unroll
  for (int vec = 0; vec < VPH; ++vec) {
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
    for (int j = 0; j < TMP_MAX_F; ++j) {
      if (j < p.fkv) {
        const uint4 vv = kv_vec(j, 1, vec);
        const float2 v0 = unpack_bf16(vv.x), v1 = unpack_bf16(vv.y), v2 = unpack_bf16(vv.z), v3 = unpack_bf16(vv.w);
        acc[0] += s[j] * v0.x; acc[1] += s[j] * v0.y; acc[2] += s[j] * v1.x; acc[3] += s[j] * v1.y;
        acc[4] += s[j] * v2.x; acc[5] += s[j] * v2.y; acc[6] += s[j] * v3.x; acc[7] += s[j] * v3.y;
      }
    }
    *reinterpret_cast<uint4*>(op + vec * 8) = make_uint4(pack_bf16(acc[0] * inv, acc[1] * inv), pack_bf16(acc[2] * inv, acc[3] * inv),
                                                          pack_bf16(acc[4] * inv, acc[5] * inv), pack_bf16(acc[6] * inv, acc[7] * inv));
  }
}

int temporal_attention_run(const dd_temporal_attention_args* a, cudaStream_t stream) {
  DD_CHECK(a != nullptr, -1, "dd_temporal_attention: null args");
  DD_CHECK(a->n_outer > 0 && a->n_view > 0 && a->tokens > 0 && a->heads > 0, -1, "dd_temporal_attention: bad shape");
  DD_CHECK(a->head_dim == 40 || a->head_dim == 80 || a->head_dim == 160, -1,
           "dd_temporal_attention: head_dim %d unsupported (40, 80, 160)", a->head_dim);
  DD_CHECK(a->frames_q >= 1 && a->frames_kv >= a->frames_q && a->frames_kv <= TMP_MAX_F, -1,
           "dd_temporal_attention: frames_q=%d frames_kv=%d (<= %d)", a->frames_q, a->frames_kv, TMP_MAX_F);
  DD_CHECK(a->frames_per_rank >= 1 && a->frames_kv % a->frames_per_rank == 0, -1,
           "dd_temporal_attention: frames_kv must be a multiple of frames_per_rank");
  DD_CHECK(a->q_ld % 8 == 0 && a->k_ld % 8 == 0 && a->v_ld % 8 == 0 && a->out_ld % 8 == 0 && a->q_col0 % 8 == 0 &&
               a->k_col0 % 8 == 0 && a->v_col0 % 8 == 0 && a->q_head_stride % 8 == 0 && a->k_head_stride % 8 == 0 &&
               a->v_head_stride % 8 == 0, -1, "dd_temporal_attention: 16-byte alignment of columns / strides required");
  DD_CHECK(a->heads * a->frames_q <= 256, -1, "dd_temporal_attention: heads * frames_q must be <= 256");
  TemporalDev p;
  p.q = reinterpret_cast<const bf16*>(a->q); p.k = reinterpret_cast<const bf16*>(a->k);
  p.v = reinterpret_cast<const bf16*>(a->v); p.out = reinterpret_cast<bf16*>(a->out);
  p.q_ld = a->q_ld; p.k_ld = a->k_ld; p.v_ld = a->v_ld; p.out_ld = a->out_ld;
  p.q_col0 = a->q_col0; p.k_col0 = a->k_col0; p.v_col0 = a->v_col0;
  p.q_hs = a->q_head_stride; p.k_hs = a->k_head_stride; p.v_hs = a->v_head_stride;
  p.n_outer = a->n_outer; p.n_view = a->n_view; p.T = a->tokens; p.heads = a->heads;
  p.fq = a->frames_q; p.fkv = a->frames_kv; p.fkv_per_rank = a->frames_per_rank;
  p.kv_rank_stride = a->kv_rank_stride;
  p.fq_per_rank = a->frames_q_per_rank > 0 ? a->frames_q_per_rank : a->frames_q;
  p.q_rank_stride = a->q_rank_stride;
  DD_CHECK(a->frames_q % p.fq_per_rank == 0, -1, "dd_temporal_attention: frames_q must be a multiple of frames_q_per_rank");
  p.scale_log2e = a->scale * 1.4426950408889634f;
  const int per_tok = a->heads * a->frames_q;
  const size_t tok_bytes = (size_t)a->frames_kv * 2 * a->heads * a->head_dim * 2;
  int tok = 256 / per_tok;
  while (tok > 1 && tok * tok_bytes > 96 * 1024) --tok;
  DD_CHECK(tok >= 1 && tok_bytes <= 200 * 1024, -1, "dd_temporal_attention: clip too large for shared memory");
  p.tok = tok;
  const size_t smem = tok * tok_bytes;
  const long long n_seq = (long long)a->n_outer * a->n_view * a->tokens;
  const unsigned grid = (unsigned)((n_seq + tok - 1) / tok);
#define DD_TMP(D)                                                                                              \
  do {                                                                                                         \
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(temporal_attn_kernel<D>), 200 * 1024)) return rc;  \
    temporal_attn_kernel<D><<<grid, 256, smem, stream>>>(p);                                                   \
  } while (0)
  if (a->head_dim == 40) DD_TMP(40);
  else if (a->head_dim == 80) DD_TMP(80);
  else DD_TMP(160);
#undef DD_TMP
  DD_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace dd
