// TMEM read/write bandwidth probe (sm_100a): bytes per clock per SM of tcgen05.ld / tcgen05.st for the shapes an
// attention / GEMM epilogue can use, with 4 or 8 warps per CTA and 1 or 2 CTAs per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bench tmem_bench.cu && ./tmem_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define R32(r) "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), \
  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), \
  "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), \
  "=r"(r[30]), "=r"(r[31])
#define I32(r) "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), \
  "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), \
  "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), \
  "r"(r[30]), "r"(r[31])
#define L32 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}"
#define S32 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32}"

enum { LD_32x32_x32, ST_32x32_x32, LD_32x32_x32_NOWAIT, LD_32x32_x32_DEEP, N_MODES };
static const char* names[] = {"tcgen05.ld.32x32b.x32 (wait each)", "tcgen05.st.32x32b.x32 (wait each)",
                              "tcgen05.ld.32x32b.x32 (4 in flight)", "tcgen05.ld.32x32b.x32 (16 in flight)"};

template <int MODE>
__global__ void k(uint32_t* out, int iters, int cols) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tptr)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t r[32], acc = 0;
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
    const uint32_t a = base + (uint32_t)((it * 32) & (cols - 32));
    if (MODE == LD_32x32_x32) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " L32 ", [%32];" : R32(r) : "r"(a) : "memory");
    if (MODE == ST_32x32_x32) asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], " S32 ";" ::"r"(a), I32(r) : "memory");
    if (MODE == LD_32x32_x32_NOWAIT || MODE == LD_32x32_x32_DEEP) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " L32 ", [%32];" : R32(r) : "r"(a) : "memory");
      if ((it & (MODE == LD_32x32_x32_DEEP ? 15 : 3)) == 3) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else if (MODE == ST_32x32_x32) {
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    acc ^= r[0] ^ r[11] ^ r[22] ^ r[31];   // fixed indices: r[] must stay in registers (a dynamic index would spill it)
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "r"(cols) : "memory");
}

template <int MODE>
static void run(uint32_t* out, int threads, int ctas_per_sm) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  const int cols = ctas_per_sm == 1 ? 512 : 256;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<148 * ctas_per_sm, threads>>>(out, iters, cols);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
  }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  // bytes per instruction per warp: 32 regs x 32 lanes x 4 B = 4096 (the 16-lane shapes move the same 4 KB per warp)
  const double bytes_per_sm = 4096.0 * (threads / 32) * ctas_per_sm * iters;
  const double cycles = ms * 1e-3 * 1965e6;
  printf("%-38s warps/CTA %d CTAs/SM %d  %8.3f ms  %7.1f B/clk/SM (at 1965 MHz) %s\n", names[MODE], threads / 32, ctas_per_sm, ms,
         bytes_per_sm / cycles, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
  uint32_t* out;
  cudaMalloc(&out, 148 * 2 * 256 * 4);
  for (int c = 1; c <= 2; ++c)
    for (int t = 128; t <= 256; t += 128) {
      run<LD_32x32_x32>(out, t, c);
      run<LD_32x32_x32_NOWAIT>(out, t, c);
      run<LD_32x32_x32_DEEP>(out, t, c);
      run<ST_32x32_x32>(out, t, c);
    }
  // one warp alone: per-SMSP / per-lane-quarter rate
  run<LD_32x32_x32_NOWAIT>(out, 32, 1);
  return 0;
}
