// dualdiff_b200 — the prompt-encoder pieces that are not GEMM / LayerNorm.
//
// The reference encodes the prompt once per sample with transformers' CLIPTextModel (SD-v1.5 text encoder: 12 layers, 768
// wide, 12 heads of 64, 77 tokens, causal mask, quick-GELU MLP) through diffusers' `_encode_prompt`
// (pipeline/pipeline_bev_controlnet.py:273-281).  Its projections and MLP run on the tcgen05 GEMM and its LayerNorms on
// dd_layernorm; this file holds the three remaining pieces:
//   * dd_clip_embed      token-embedding gather + position embedding (fp32 tables -> bf16 rows),
//   * dd_seq_attention   attention over one short sequence (L <= 128 tokens, head_dim 64) with an optional causal mask.
//                        A (sequence, head) problem is 77 x 77 x 64: far below one UMMA tile and launched once per sample,
//                        so it runs on CUDA cores: K/V of the head staged in shared memory, one warp per query row,
//                        lanes own keys for the scores and channel pairs for P V, softmax by warp shuffles,
//   * dd_quick_gelu      x * sigmoid(1.702 x) over bf16 (kept out of the GEMM epilogues: those are sized for the
//                        denoising step's instruction cache, and this runs 12 times per prompt batch).
#include "dd_api_internal.h"
#include "dd_common.cuh"

namespace dd {

static inline int clip_grid_for(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// out[r, :] = tok_emb[ids[r], :] + pos_emb[r % L, :]; an id outside [0, vocab) poisons its row with NaN (the host
// wrapper validates the ids before the copy, torch.nn.Embedding would raise)
__global__ void __launch_bounds__(256)
clip_embed_kernel(const long long* __restrict__ ids, const float* __restrict__ tok, const float* __restrict__ pos,
                  bf16* __restrict__ out, long long out_ld, long long n_tok, int L, int C, int vocab) {
  const int c4 = C >> 2;
  const long long total = n_tok * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c4;
    const int c = (int)(i - r * c4) * 4;
    const long long id = ids[r];
    float4 v;
    if (id < 0 || id >= vocab) {
      v.x = v.y = v.z = v.w = __int_as_float(0x7fc00000);
    } else {
      const float4 t = *reinterpret_cast<const float4*>(tok + id * C + c);
      const float4 p = *reinterpret_cast<const float4*>(pos + (r % L) * C + c);
      v = make_float4(t.x + p.x, t.y + p.y, t.z + p.z, t.w + p.w);
    }
    *reinterpret_cast<uint2*>(out + r * out_ld + c) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

__global__ void __launch_bounds__(256)
quick_gelu_kernel(const uint4* x, uint4* out, long long n8) {   // no __restrict__: in place is allowed
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 u = x[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_bf16(w[e]);
      o[e] = pack_bf16(__fdividef(f.x, 1.f + __expf(-1.702f * f.x)), __fdividef(f.y, 1.f + __expf(-1.702f * f.y)));
    }
    out[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

static constexpr int SEQ_MAX_L = 128;
static constexpr int SEQ_D = 64;
static constexpr int SEQ_KLD = SEQ_D + 8;   // bf16 elements per staged K row: 144 B, 16-byte reads of 8 rows hit 8 bank groups
static constexpr int SEQ_WARPS = 4;

struct SeqAttnDev {
  const bf16* q; const bf16* k; const bf16* v; bf16* out;
  long long q_ld, k_ld, v_ld, out_ld;
  int q_col0, k_col0, v_col0, q_hs, k_hs, v_hs;
  int L, heads, causal;
  float scale_log2e;
};

// grid = n_seq * heads; CTA = 4 warps; warp w takes query rows w, w + 4, ...
__global__ void __launch_bounds__(SEQ_WARPS * 32)
seq_attention_kernel(const SeqAttnDev p) {
  __shared__ __align__(16) bf16 sK[SEQ_MAX_L * SEQ_KLD];
  __shared__ __align__(16) bf16 sV[SEQ_MAX_L * SEQ_D];
  const int seq = blockIdx.x / p.heads, h = blockIdx.x - seq * p.heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)seq * p.L;
  // ---- stage K and V of this (sequence, head): L rows of 64 bf16 = 8 x 16 B each ----
  for (int i = threadIdx.x; i < p.L * 16; i += blockDim.x) {
    const int r = i >> 4, which = (i >> 3) & 1, vec = i & 7;
    if (which == 0) {
      const uint4 u = *reinterpret_cast<const uint4*>(p.k + (row0 + r) * p.k_ld + p.k_col0 + h * p.k_hs + vec * 8);
      *reinterpret_cast<uint4*>(sK + r * SEQ_KLD + vec * 8) = u;
    } else {
      const uint4 u = *reinterpret_cast<const uint4*>(p.v + (row0 + r) * p.v_ld + p.v_col0 + h * p.v_hs + vec * 8);
      *reinterpret_cast<uint4*>(sV + r * SEQ_D + vec * 8) = u;
    }
  }
  __syncthreads();
  for (int r = warp; r < p.L; r += SEQ_WARPS) {
    const int kmax = p.causal ? r + 1 : p.L;       // keys [0, kmax) are visible (warp-uniform)
    // every lane holds the whole query row (the 128-byte row is one broadcast transaction per 16-byte vector)
    float qf[SEQ_D];
    const bf16* qp = p.q + (row0 + r) * p.q_ld + p.q_col0 + h * p.q_hs;
#pragma unroll
    for (int vec = 0; vec < SEQ_D / 8; ++vec) {
      const uint4 u = *reinterpret_cast<const uint4*>(qp + vec * 8);
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      qf[vec * 8 + 0] = a.x; qf[vec * 8 + 1] = a.y; qf[vec * 8 + 2] = b.x; qf[vec * 8 + 3] = b.y;
      qf[vec * 8 + 4] = c.x; qf[vec * 8 + 5] = c.y; qf[vec * 8 + 6] = d.x; qf[vec * 8 + 7] = d.y;
    }
    // ---- scores: lane owns keys lane, lane + 32, lane + 64, lane + 96 ----
    float s[SEQ_MAX_L / 32];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < SEQ_MAX_L / 32; ++i) {
      const int j = lane + 32 * i;
      float acc = -INFINITY;
      if (j < kmax) {
        acc = 0.f;
        const bf16* kr = sK + j * SEQ_KLD;
#pragma unroll
        for (int vec = 0; vec < SEQ_D / 8; ++vec) {
          const uint4 u = *reinterpret_cast<const uint4*>(kr + vec * 8);
          const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
          acc = fmaf(qf[vec * 8 + 0], a.x, acc); acc = fmaf(qf[vec * 8 + 1], a.y, acc);
          acc = fmaf(qf[vec * 8 + 2], b.x, acc); acc = fmaf(qf[vec * 8 + 3], b.y, acc);
          acc = fmaf(qf[vec * 8 + 4], c.x, acc); acc = fmaf(qf[vec * 8 + 5], c.y, acc);
          acc = fmaf(qf[vec * 8 + 6], d.x, acc); acc = fmaf(qf[vec * 8 + 7], d.y, acc);
        }
      }
      s[i] = acc;
      m = fmaxf(m, acc);
    }
    m = warp_max(m);                                // key 0 is always visible -> m is finite
    float l = 0.f;
#pragma unroll
    for (int i = 0; i < SEQ_MAX_L / 32; ++i) {
      s[i] = (lane + 32 * i < kmax) ? exp2f((s[i] - m) * p.scale_log2e) : 0.f;
      l += s[i];
    }
    l = warp_sum(l);
    // ---- O = P V: lane owns channels 2*lane, 2*lane + 1; p_j comes from its owner lane by shuffle ----
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int i = 0; i < SEQ_MAX_L / 32; ++i) {
      if (32 * i < kmax) {                          // warp-uniform
        const int n = min(32, kmax - 32 * i);
        for (int jj = 0; jj < n; ++jj) {
          const float pj = __shfl_sync(0xffffffffu, s[i], jj);
          const float2 vv = unpack_bf16(*reinterpret_cast<const uint32_t*>(sV + (32 * i + jj) * SEQ_D + 2 * lane));
          o0 = fmaf(pj, vv.x, o0);
          o1 = fmaf(pj, vv.y, o1);
        }
      }
    }
    const float inv = 1.f / l;
    *reinterpret_cast<uint32_t*>(p.out + (row0 + r) * p.out_ld + h * SEQ_D + 2 * lane) = pack_bf16(o0 * inv, o1 * inv);
  }
}

int seq_attention_run(const dd_seq_attention_args* a, cudaStream_t stream) {
  DD_CHECK(a != nullptr, -1, "dd_seq_attention: null args");
  DD_CHECK(a->n_seq > 0 && a->heads > 0 && a->seq_len > 0, -1, "dd_seq_attention: bad shape");
  DD_CHECK(a->seq_len <= SEQ_MAX_L, -1, "dd_seq_attention: seq_len %d > %d", a->seq_len, SEQ_MAX_L);
  DD_CHECK(a->head_dim == SEQ_D, -1, "dd_seq_attention: head_dim %d unsupported (64 = CLIP ViT-L/14 text encoder)", a->head_dim);
  DD_CHECK(a->q_ld % 8 == 0 && a->k_ld % 8 == 0 && a->v_ld % 8 == 0 && a->out_ld % 2 == 0 && a->q_col0 % 8 == 0 &&
               a->k_col0 % 8 == 0 && a->v_col0 % 8 == 0 && a->q_head_stride % 8 == 0 && a->k_head_stride % 8 == 0 &&
               a->v_head_stride % 8 == 0, -1, "dd_seq_attention: 16-byte alignment of columns / strides required");
  DD_CHECK((((uintptr_t)a->q | (uintptr_t)a->k | (uintptr_t)a->v) & 15) == 0 && ((uintptr_t)a->out & 3) == 0, -1,
           "dd_seq_attention: pointer alignment");
  DD_CHECK((long long)a->n_seq * a->heads < (1LL << 31), -1, "dd_seq_attention: too many (sequence, head) problems");
  SeqAttnDev p;
  p.q = reinterpret_cast<const bf16*>(a->q); p.k = reinterpret_cast<const bf16*>(a->k);
  p.v = reinterpret_cast<const bf16*>(a->v); p.out = reinterpret_cast<bf16*>(a->out);
  p.q_ld = a->q_ld; p.k_ld = a->k_ld; p.v_ld = a->v_ld; p.out_ld = a->out_ld;
  p.q_col0 = a->q_col0; p.k_col0 = a->k_col0; p.v_col0 = a->v_col0;
  p.q_hs = a->q_head_stride; p.k_hs = a->k_head_stride; p.v_hs = a->v_head_stride;
  p.L = a->seq_len; p.heads = a->heads; p.causal = a->causal ? 1 : 0;
  p.scale_log2e = a->scale * 1.4426950408889634f;
  seq_attention_kernel<<<(unsigned)(a->n_seq * a->heads), SEQ_WARPS * 32, 0, stream>>>(p);
  DD_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace dd

using namespace dd;
extern "C" {
int dd_clip_embed(const long long* ids, const float* tok_emb, const float* pos_emb, void* out, long long out_ld,
                  long long n_tok, int seq_len, int c, int vocab, void* stream) {
  if (ids == nullptr || tok_emb == nullptr || pos_emb == nullptr || out == nullptr || n_tok <= 0 || seq_len <= 0 ||
      vocab <= 0 || c <= 0 || c % 4 != 0 || out_ld % 4 != 0) {
    set_error("dd_clip_embed: bad arguments (c and out_ld must be multiples of 4)");
    return -1;
  }
  clip_embed_kernel<<<clip_grid_for(n_tok * (c >> 2), 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      ids, tok_emb, pos_emb, reinterpret_cast<bf16*>(out), out_ld, n_tok, seq_len, c, vocab);
  if (cudaGetLastError() != cudaSuccess) { set_error("dd_clip_embed launch failed"); return -2; }
  count_launch();
  return 0;
}
int dd_quick_gelu(const void* x, void* out, long long n, void* stream) {
  if (n <= 0 || n % 8 != 0) { set_error("dd_quick_gelu: n must be a positive multiple of 8"); return -1; }
  quick_gelu_kernel<<<clip_grid_for(n / 8, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), n / 8);
  if (cudaGetLastError() != cudaSuccess) { set_error("dd_quick_gelu launch failed"); return -2; }
  count_launch();
  return 0;
}
int dd_seq_attention(const dd_seq_attention_args* args, void* stream) {
  int rc = seq_attention_run(args, reinterpret_cast<cudaStream_t>(stream));
  if (rc == 0) count_launch();
  return rc;
}
}
