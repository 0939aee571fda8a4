/* dualdiff_b200 — C ABI of libdualdiff_sm100.so
 *
 * The drop-in boundary for the DualDiff denoising-step hot path.  The reference is
 * pure Python and reaches its kernels through torch/diffusers/xformers; it has no FFI of its own, so
 * each entry point below names the reference call site (file:line under
 * /root/reference/MD_txt_con_fusion/magicdrive) whose arithmetic it replaces.  The Python modules in
 * dualdiff_b200/ (same class names / forward signatures / state-dict keys as the reference) bind these
 * with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer is a raw CUDA device pointer (tensor.data_ptr()); PyTorch owns all memory;
 *   - activations are bf16, channels-last: "compact" = [rows = img*H*W, C]; "padded" = the zero-haloed
 *     pixel layout [img][H+1][W+1][C] consumed by the 3x3 implicit-GEMM convolution;
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*), performs no
 *     allocation and no host synchronisation, and is safe under CUDA-graph capture;
 *   - return 0 on success, negative on error; dd_last_error() returns a thread-local message.
 *     There is no CPU fallback: without an sm_100 device the calls fail.
 */
#ifndef DUALDIFF_B200_H_
#define DUALDIFF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DD_VERSION 100
#if defined(__GNUC__)
#define DD_API __attribute__((visibility("default")))
#else
#define DD_API
#endif

DD_API int dd_version(void);
DD_API const char* dd_last_error(void);
/* number of kernels launched by this library since load (bench.py `gpu_launches`) */
DD_API long long dd_launch_count(void);

/* ---- GEMM / implicit-GEMM convolution on tcgen05 tensor cores ------------------------------------
 * out = epilogue( A[M, K(*taps)] x W[N, taps*K]^T )
 * replaces: torch.nn.Linear / Conv2d(1x1, 3x3 s1 p1) inside diffusers ResnetBlock2D, Transformer2DModel,
 * Attention.to_q/k/v/out, FeedForward (GEGLU), Upsample2D.conv, zero convs
 * (networks/unet_addon_rawbox.py:965,1031-1039; networks/unet_2d_condition_multiview.py:443,522;
 *  networks/blocks.py:72-83).
 */
typedef struct dd_gemm_args {
  const void* a;      /* bf16 [M, k1 or K] row-major, leading dim a_ld (elements)               */
  const void* a2;     /* optional second K-source (skip-connection concat), cols [k1, K)        */
  const void* w;      /* bf16 [N, taps*K] row-major (tap-major, channel-minor), leading dim w_ld */
  void* out;          /* bf16 or fp32 [rows, n_store], leading dim out_ld                        */
  const float* bias;  /* fp32 [N] or NULL                                                        */
  const float* rowvec;/* fp32 [n_img, rowvec_ld]: per-image vector added to every row, or NULL   */
  const void* res1;   /* bf16 residual [rows, >=n_store] or NULL                                 */
  const void* res2;   /* second bf16 residual or NULL                                            */
  int M, N, K;        /* K = channels per tap                                                    */
  int k1;             /* split point when a2 != NULL (multiple of 64)                            */
  int taps;           /* 1 = linear / 1x1 conv; 9 = 3x3 stride-1 pad-1 conv over the padded layout */
  int conv_h, conv_w; /* taps == 9: image height/width (padded layout has (H+1) x (W+1) rows)    */
  long long a_ld, a2_ld, w_ld, out_ld, res1_ld, res2_ld;
  int rowvec_ld, rows_per_img;
  int out_f32;        /* 1: fp32 output                                                          */
  int geglu;          /* 1: out[:, j] = v[:, j] * gelu(g[:, j]); W packed as 128-value / 128-gate
                         column groups per 256-wide tile (dd pack_geglu in dualdiff_b200/packing.py) */
  int force_bn;       /* 0 = auto tile width; testing hook                                        */
} dd_gemm_args;
DD_API int dd_gemm(const dd_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DUALDIFF_B200_H_ */
