// dualdiff_b200 — Occupancy Ray-shape Sampling projector.
// Replaces networks/occ3d_proj.py:50-113 (OccupancyRay.project): the reference builds an 18-channel one-hot volume
// (1 x 18 x 200 x 200 x 16 floats) and runs F.grid_sample(mode='nearest') + argmax over 6 x h x w x sample_point ray
// samples on the CPU inside collate_fn.  Here one thread per sample looks the voxel label up directly (640 KB of uint8
// labels, L2-resident) -- gather-bound, no one-hot, no argmax -- and can emit, besides the class ids, the normalised
// bf16 rows [6*h*w, sample_point] the foreground ControlNet branch consumes (dataset/utils.py:412-420: fg/bg class
// filter, /17) so that no NCHW tensor is ever materialised.
// The arithmetic mirrors the reference operation by operation (no FMA contraction) so that, given the same ray origins
// and directions, every sample lands in the same voxel: points = o + step*d; /40; z' = z*40/3.2 - 2.2/3.2;
// index = nearbyint(((g + 1) * size - 1) / 2)   (grid_sample, align_corners=False).
#include "dd_api_internal.h"
#include "dd_common.cuh"

namespace dd {

__global__ void __launch_bounds__(256)
ors_project_kernel(const float* __restrict__ origins, const float* __restrict__ dirs,
                   const unsigned char* __restrict__ sem, unsigned char* __restrict__ ids, bf16* __restrict__ rows,
                   long long n_pix, int S, float step, int D, int H, int W, int keep_fg, int keep_bg) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pix * S) return;
  const long long pix = i / S;
  const int s = (int)(i - pix * S);
  const float t = __fmul_rn((float)s, step);
  const float off = (float)(2.2 / 3.2);
  float g[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float p = __fadd_rn(origins[pix * 3 + a], __fmul_rn(t, dirs[pix * 3 + a]));   // metres
    g[a] = __fdiv_rn(p, 40.f);
  }
  g[2] = __fsub_rn(__fdiv_rn(__fmul_rn(g[2], 40.f), 3.2f), off);
  // grid_sample axes: x -> last volume axis (height, W entries), y -> middle axis, z -> first axis (occ3d_proj.py:93-94)
  const float fx = rintf(__fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(g[2], 1.f), (float)W), 1.f), 2.f));
  const float fy = rintf(__fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(g[1], 1.f), (float)H), 1.f), 2.f));
  const float fz = rintf(__fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(g[0], 1.f), (float)D), 1.f), 2.f));
  int lab = 17;   // outside the volume (zeros padding -> class 17, occ3d_proj.py:103-104)
  if (fx >= 0.f && fx < (float)W && fy >= 0.f && fy < (float)H && fz >= 0.f && fz < (float)D)
    lab = sem[((long long)(int)fz * H + (int)fy) * W + (int)fx];
  if (ids) ids[i] = (unsigned char)lab;
  if (rows) {
    int v = lab;
    if (!keep_fg && v <= 10) v = 17;   // dataset/utils.py:414-417
    if (!keep_bg && v >= 11) v = 17;
    rows[i] = __float2bfloat16(__fdiv_rn((float)v, 17.f));
  }
}

int ors_project_run(const float* origins, const float* dirs, const unsigned char* sem, unsigned char* ids, void* rows,
                    long long n_pix, int sample_point, float sample_step, int D, int H, int W, int keep_fg, int keep_bg,
                    cudaStream_t stream) {
  DD_CHECK(origins && dirs && sem && (ids || rows), -1, "dd_ors_project: null pointer");
  DD_CHECK(n_pix > 0 && sample_point > 0 && D > 0 && H > 0 && W > 0, -1, "dd_ors_project: bad shape");
  const long long total = n_pix * sample_point;
  ors_project_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(origins, dirs, sem, ids, reinterpret_cast<bf16*>(rows),
                                                                          n_pix, sample_point, sample_step, D, H, W, keep_fg, keep_bg);
  DD_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace dd
