// dualdiff_b200 — C-ABI entry points, error plumbing, TMA descriptor encode.
#include <stdarg.h>
#include <stdio.h>
#include <atomic>
#include <mutex>
#include <vector>

#include "dd_api_internal.h"
#include "dd_common.cuh"

namespace dd {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  static std::atomic<int> sms_of[64];   // per device, 0 = not queried yet (benign race: every thread stores the same value)
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  int sms = sms_of[dev].load(std::memory_order_relaxed);
  if (sms == 0) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    sms_of[dev].store(sms, std::memory_order_relaxed);
  }
  return sms;
}

int ensure_dyn_smem(const void* func, int bytes) {
  struct Done { const void* func; int dev; int bytes; };
  static std::mutex mu;
  static std::vector<Done> done;
  int dev = 0;
  DD_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  for (const Done& d : done)
    if (d.func == func && d.dev == dev && d.bytes >= bytes) return 0;
  DD_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done.push_back({func, dev, bytes});
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn enc = get_encode();
  DD_CHECK(enc != nullptr, -3, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DD_CHECK(r == CUDA_SUCCESS, -3, "cuTensorMapEncodeTiled(2d) failed: %d (rows=%llu cols=%llu ld=%llu box=%ux%u)",
           (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
  return 0;
}

int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                      uint64_t s1, uint64_t s2, uint32_t b0, uint32_t b1, uint32_t b2) {
  EncodeTiledFn enc = get_encode();
  DD_CHECK(enc != nullptr, -3, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1 * 2, s2 * 2};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DD_CHECK(r == CUDA_SUCCESS, -3, "cuTensorMapEncodeTiled(3d) failed: %d", (int)r);
  return 0;
}

}  // namespace dd

extern "C" {
int dd_version(void) { return DD_VERSION; }
const char* dd_last_error(void) { return dd::g_err; }
long long dd_launch_count(void) { return dd::g_launches.load(); }
int dd_gemm(const dd_gemm_args* args, void* stream) {
  int rc = dd::gemm_run(args, reinterpret_cast<cudaStream_t>(stream));
  if (rc == 0) dd::count_launch();
  return rc;
}
int dd_groupnorm(const dd_groupnorm_args* args, void* stream) {
  return dd::groupnorm_run(args, reinterpret_cast<cudaStream_t>(stream));
}
long long dd_groupnorm_scratch_floats(int n_img, int c, int hw) { return dd::groupnorm_scratch_floats(n_img, c, hw); }
int dd_layernorm(const dd_layernorm_args* args, void* stream) {
  return dd::layernorm_run(args, reinterpret_cast<cudaStream_t>(stream));
}
int dd_attention(const dd_attention_args* args, void* stream) {
  int rc = dd::attention_run(args, reinterpret_cast<cudaStream_t>(stream));
  if (rc == 0) dd::count_launch();
  return rc;
}
int dd_ors_project(const float* origins, const float* dirs, const unsigned char* sem, unsigned char* ids, void* rows,
                   long long n_pix, int sample_point, float sample_step, int D, int H, int W, int keep_fg, int keep_bg,
                   void* stream) {
  int rc = dd::ors_project_run(origins, dirs, sem, ids, rows, n_pix, sample_point, sample_step, D, H, W, keep_fg, keep_bg,
                               reinterpret_cast<cudaStream_t>(stream));
  if (rc == 0) dd::count_launch();
  return rc;
}
int dd_temporal_attention(const dd_temporal_attention_args* args, void* stream) {
  int rc = dd::temporal_attention_run(args, reinterpret_cast<cudaStream_t>(stream));
  if (rc == 0) dd::count_launch();
  return rc;
}
}
