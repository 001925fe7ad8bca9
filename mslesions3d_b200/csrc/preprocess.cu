// Intensity normalisation of the data module on the device (datasets.py:403, MONAI NormalizeIntensity(nonzero=True)):
// per (volume, channel) z-score over the NON-ZERO voxels, zeros stay zero; the result is written in the stem's
// input format (NCDHW, fp32 or bf16), so a raw batch needs one pass here instead of a host-side numpy pass.
//   stage 1: per block partial {count, sum, sum of squares} in fp64          (grid: blocks_per_item x items)
//   stage 2: mean / population std per item (std == 0 -> 1, as MONAI divides only when std != 0)
//   stage 3: y = x != 0 ? (x - mean) / std : 0
#include "common.cuh"
#include "../../include/ssd3d_b200.h"

namespace ssd3d {

constexpr int NORM_BLOCKS = 64;     // partial blocks per (volume, channel)

__global__ void __launch_bounds__(256) norm_partial_kernel(const float* __restrict__ x, long long vox,
                                                           double* __restrict__ partial) {
  __shared__ double red[3][8];
  const int item = blockIdx.y;
  const float* src = x + (long long)item * vox;
  double cnt = 0.0, s = 0.0, q = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < vox; i += (long long)gridDim.x * blockDim.x) {
    const float v = __ldg(src + i);
    if (v != 0.f) { cnt += 1.0; s += (double)v; q += (double)v * (double)v; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    s += __shfl_down_sync(0xffffffffu, s, o);
    q += __shfl_down_sync(0xffffffffu, q, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = cnt; red[1][warp] = s; red[2][warp] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; c += red[2][w]; }
    double* dst = partial + ((size_t)item * gridDim.x + blockIdx.x) * 3;
    dst[0] = a; dst[1] = b; dst[2] = c;
  }
}

__global__ void norm_finalize_kernel(const double* __restrict__ partial, int blocks, int items,
                                     float* __restrict__ mean_std) {
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= items) return;
  double cnt = 0.0, s = 0.0, q = 0.0;
  for (int b = 0; b < blocks; ++b) {
    const double* p = partial + ((size_t)item * blocks + b) * 3;
    cnt += p[0]; s += p[1]; q += p[2];
  }
  double mean = 0.0, sd = 1.0;
  if (cnt > 0.0) {
    mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    sd = sqrt(var);
    if (sd == 0.0) sd = 1.0;
  }
  mean_std[2 * item] = (float)mean;
  mean_std[2 * item + 1] = (float)sd;
}

template <typename TOut>
__global__ void __launch_bounds__(256) norm_apply_kernel(const float* __restrict__ x, long long vox,
                                                         const float* __restrict__ mean_std, TOut* __restrict__ y) {
  const int item = blockIdx.y;
  const float mean = mean_std[2 * item], sd = mean_std[2 * item + 1];
  const float* src = x + (long long)item * vox;
  TOut* dst = y + (long long)item * vox;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < vox; i += (long long)gridDim.x * blockDim.x) {
    const float v = __ldg(src + i);
    const float r = (v != 0.f) ? __fdiv_rn(__fsub_rn(v, mean), sd) : 0.f;
    if constexpr (sizeof(TOut) == 2) dst[i] = __float2bfloat16_rn(r);
    else dst[i] = r;
  }
}

}  // namespace ssd3d

using namespace ssd3d;

extern "C" int64_t ssd3d_normalize_workspace_bytes(int items) {
  return (int64_t)items * NORM_BLOCKS * 3 * 8 + (int64_t)items * 2 * 4 + 256;
}

extern "C" int ssd3d_normalize_intensity_nonzero(const float* x, int items, int64_t voxels, void* y, int y_is_bf16,
                                                 void* workspace, int64_t workspace_bytes, void* stream) {
  if (!x || !y || !workspace || items <= 0 || voxels <= 0) return SSD3D_ERR_ARG;
  if (workspace_bytes < ssd3d_normalize_workspace_bytes(items)) return SSD3D_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(workspace);
  float* mean_std = reinterpret_cast<float*>(partial + (size_t)items * NORM_BLOCKS * 3);
  dim3 grid(NORM_BLOCKS, (unsigned)items);
  norm_partial_kernel<<<grid, 256, 0, st>>>(x, (long long)voxels, partial);
  SSD3D_CHECK_LAUNCH();
  norm_finalize_kernel<<<(items + 127) / 128, 128, 0, st>>>(partial, NORM_BLOCKS, items, mean_std);
  SSD3D_CHECK_LAUNCH();
  if (y_is_bf16)
    norm_apply_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(x, (long long)voxels, mean_std, static_cast<__nv_bfloat16*>(y));
  else
    norm_apply_kernel<float><<<grid, 256, 0, st>>>(x, (long long)voxels, mean_std, static_cast<float*>(y));
  SSD3D_CHECK_LAUNCH();
  return SSD3D_OK;
}
