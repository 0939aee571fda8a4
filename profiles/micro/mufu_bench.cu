// Issue-rate probe for the softmax inner loop of dd_attention.cu (sm_100a): thread-level instructions per clock per SM
// of the candidate instructions.  8 independent chains per thread, 1024 threads per SM, 148 CTAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu && ./mufu_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

enum { EX2_F32, EX2_BF16X2, EX2_F16X2, FMA_F32, FMA_F32X2, MAX_F32, MAX3_F32, CVT_BF16X2, ADD_F32X2, MIX_OLD, MIX_NEW, N_MODES };
static const char* names[] = {"ex2.approx.ftz.f32", "ex2.approx.ftz.bf16x2", "ex2.approx.f16x2", "fma.rn.f32", "fma.rn.f32x2",
                              "max.f32", "max.f32 (3-input)", "cvt.rn.bf16x2.f32", "add.f32x2",
                              "softmax mix OLD (per 2 elems: 2 ffma 2 ex2.f32 2 fadd 2 fmax 1 cvt)",
                              "softmax mix NEW (per 2 elems: 1 ffma2 1 cvt 1 ex2.bf16x2 1 max3)"};
static const int elems[] = {1, 2, 2, 1, 2, 1, 2, 2, 2, 2, 2};

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, int iters, float seed) {
  float a[8], c[8];
  uint32_t b[8];
  unsigned long long d[8];
  for (int i = 0; i < 8; ++i) {
    a[i] = -0.001f * (threadIdx.x + i) * seed;
    c[i] = a[i] * 0.5f;
    b[i] = 0xBC00BC00u + threadIdx.x + i;
    d[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(c[i]);
  }
  const float s1 = 0.999f * seed, s2 = -0.0001f * seed;
  const unsigned long long s1x2 = ((unsigned long long)__float_as_uint(s1) << 32) | __float_as_uint(s1);
  const unsigned long long s2x2 = ((unsigned long long)__float_as_uint(s2) << 32) | __float_as_uint(s2);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == EX2_F32) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == EX2_BF16X2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b[i]));
      if (MODE == EX2_F16X2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b[i]));
      if (MODE == FMA_F32) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(s1), "f"(s2));
      if (MODE == FMA_F32X2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(s1x2), "l"(s2x2));
      if (MODE == ADD_F32X2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(s2x2));
      if (MODE == MAX_F32) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c[(i + 1) & 7]));
      if (MODE == MAX3_F32) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c[(i + 1) & 7]), "f"(c[(i + 2) & 7]));
      if (MODE == CVT_BF16X2) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b[i]) : "f"(a[i]), "f"(__uint_as_float(b[(i + 1) & 7])));
      if (MODE == MIX_OLD) {
        float x0, x1;
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(x0) : "f"(a[i]), "f"(s1), "f"(s2));
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(x1) : "f"(c[i]), "f"(s1), "f"(s2));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
        asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(x0));
        asm volatile("max.f32 %0, %0, %1;" : "+f"(c[i]) : "f"(x1));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(a[(i + 1) & 7]) : "f"(x0));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(c[(i + 1) & 7]) : "f"(x1));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b[i]) : "f"(x0), "f"(x1));
      }
      if (MODE == MIX_NEW) {
        unsigned long long x;
        uint32_t pk;
        float lo, hi;
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x) : "l"(d[i]), "l"(s1x2), "l"(s2x2));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(hi), "f"(lo));
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(pk));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(lo), "f"(hi));
        b[i] ^= pk;
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + c[i] + __uint_as_float(b[i]) + __uint_as_float((uint32_t)d[i]) + __uint_as_float((uint32_t)(d[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(float* out) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 4000;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<148, 1024>>>(out, iters, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
  }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  const double instr = 148.0 * 1024 * 8.0 * iters;  // thread-level (groups of) instructions
  const double clk_khz = 1965e3;                      // B200 max SM clock; nvidia-smi sampled by the caller
  const double cycles = ms * 1e-3 * clk_khz * 1e3;
  printf("%-72s %8.3f ms  %7.2f thread-instr/clk/SM  %7.2f elements/clk/SM  %s\n", names[MODE], ms, instr / cycles / 148.0,
         instr / cycles / 148.0 * elems[MODE], err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4);
  run<EX2_F32>(out);
  run<EX2_BF16X2>(out);
  run<EX2_F16X2>(out);
  run<FMA_F32>(out);
  run<FMA_F32X2>(out);
  run<ADD_F32X2>(out);
  run<MAX_F32>(out);
  run<MAX3_F32>(out);
  run<CVT_BF16X2>(out);
  run<MIX_OLD>(out);
  run<MIX_NEW>(out);
  return 0;
}
