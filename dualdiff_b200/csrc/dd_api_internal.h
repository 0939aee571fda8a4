// internal declarations shared by the .cu translation units
#pragma once
#include <cuda_runtime.h>

#include "../../include/dualdiff_b200.h"

namespace dd {
int gemm_run(const dd_gemm_args* a, cudaStream_t stream);
int groupnorm_run(const dd_groupnorm_args* a, cudaStream_t stream);
long long groupnorm_scratch_floats(int n_img, int C, int HW);
int layernorm_run(const dd_layernorm_args* a, cudaStream_t stream);
int attention_run(const dd_attention_args* a, cudaStream_t stream);
int temporal_attention_run(const dd_temporal_attention_args* a, cudaStream_t stream);
int ors_project_run(const float* origins, const float* dirs, const unsigned char* sem, unsigned char* ids, void* rows,
                    long long n_pix, int sample_point, float sample_step, int D, int H, int W, int keep_fg, int keep_bg,
                    cudaStream_t stream);
int seq_attention_run(const dd_seq_attention_args* a, cudaStream_t stream);
void count_launch(int n = 1);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device); thread-safe (function attributes are
// per device context, and the C entry points may be called from several host threads)
int ensure_dyn_smem(const void* func, int bytes);
}  // namespace dd
