// Does a SWIZZLE_128B K-major UMMA operand descriptor accept a start address that is shifted by whole 128-byte rows
// (not 1024-byte aligned)?  If so, the three kw taps of a 3x3 implicit-GEMM conv can share ONE staged activation box
// (rows m0-1 .. m0+128): tap kw reads rows [kw, kw+128) of it.  Tests shift d = 0..7 with the descriptor's
// base_offset field = 0 and = d ((start >> 7) & 7), M = 128, N = 64, K = 64 (4 UMMAs), against a host reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../dualdiff_b200/csrc -o desc_shift_test desc_shift_test.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>
#include "dd_common.cuh"

namespace dd { void set_error(const char*, ...) {} }
using namespace dd;

constexpr int ROWS = 144, N = 64, K = 64, NMODE = 17;   // mode 16: A = fp16, B = bf16 (mixed operand formats)

__global__ void __launch_bounds__(128) k(const bf16* A, const bf16* B, float* out, const __half* A16) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base, sB = base + ROWS * 128;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tptr;
  // swizzled fill (what TMA SWIZZLE_128B produces for a 64-column bf16 box at a 1024-aligned address)
  for (int i = threadIdx.x; i < ROWS * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(gen + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * K + c * 8);
  }
  for (int i = threadIdx.x; i < N * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(gen + ROWS * 128 + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * K + c * 8);
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(smem_u32(&tptr), 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tptr;
  constexpr uint32_t IDESC = umma_idesc_bf16(128, N, 0, 0);
  for (int mode = 0; mode < NMODE; ++mode) {
    const int d = mode == 16 ? 0 : mode >> 1;
    const uint64_t bo = (mode & 1) ? (uint64_t)d : 0ull;
    if (mode == 16) {   // refill A with fp16 data
      for (int i = threadIdx.x; i < ROWS * 8; i += blockDim.x) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(gen + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A16 + r * K + c * 8);
      }
      fence_proxy_async_smem();
      __syncthreads();
    }
    const uint32_t idesc = mode == 16 ? (IDESC & ~(7u << 7)) : IDESC;   // a_format: 1 = bf16, 0 = f16
    if (threadIdx.x == 0) {
      for (int kk = 0; kk < K / 16; ++kk) {
        const uint64_t dA = umma_smem_desc(sA + d * 128, 16, 1024, 2) | (bo << 49);
        const uint64_t dB = umma_smem_desc(sB, 16, 1024, 2);
        umma_bf16(tmem, dA + 2 * kk, dB + 2 * kk, idesc, kk != 0 ? 1u : 0u);
      }
      umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), mode & 1);
    tc_fence_after();
    uint32_t v[32];
    const uint32_t lane_sel = (uint32_t)((threadIdx.x >> 5) * 32) << 16;
    for (int c = 0; c < N; c += 32) {
      tmem_ld_32x32(tmem + lane_sel + c, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out[((size_t)mode * 128 + threadIdx.x) * N + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
  }
  if (threadIdx.x < 32) tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<bf16> hA(ROWS * K), hB(N * K);
  std::vector<float> fA(ROWS * K), fB(N * K);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)(rand() % 17 - 8); hA[i] = __float2bfloat16(fA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)(rand() % 9 - 4); hB[i] = __float2bfloat16(fB[i]); }
  bf16 *dA, *dB;
  std::vector<__half> hA16(ROWS * K);
  for (size_t i = 0; i < hA16.size(); ++i) hA16[i] = __float2half(fA[i]);
  __half* dA16;
  cudaMalloc(&dA16, hA16.size() * 2);
  cudaMemcpy(dA16, hA16.data(), hA16.size() * 2, cudaMemcpyHostToDevice);
  float* dO;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, (size_t)NMODE * 128 * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  const int smem = (ROWS + N) * 128 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<1, 128, smem>>>(dA, dB, dO, dA16);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> hO((size_t)NMODE * 128 * N);
  cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
  for (int mode = 0; mode < NMODE; ++mode) {
    const int d = mode == 16 ? 0 : mode >> 1;
    int bad = 0;
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < N; ++n) {
        float ref = 0.f;
        for (int kk = 0; kk < K; ++kk) ref += fA[(r + d) * K + kk] * fB[n * K + kk];
        if (ref != hO[((size_t)mode * 128 + r) * N + n]) ++bad;
      }
    if (mode == 16) printf("A = fp16, B = bf16 (mixed formats): %s (%d / %d mismatches)\n", bad ? "WRONG" : "exact", bad, 128 * N);
    else printf("row shift %d, base_offset %d: %s (%d / %d mismatches)\n", d, (mode & 1) ? d : 0, bad ? "WRONG" : "exact", bad, 128 * N);
  }
  return 0;
}
