// The six public box functions of the reference's utils.py:42-149 as device kernels: four coordinate
// transforms and the all-pairs intersection / Jaccard overlap.  One thread per box (transforms) or per
// pair (overlaps, set-2 boxes staged in shared memory); all arithmetic through boxes.cuh so that results
// are bit-identical with the separately-rounded fp32 torch ops the reference executes.
#include "boxes.cuh"

namespace ssd3d {

__global__ void __launch_bounds__(256) box_transform_kernel(int mode, const float* __restrict__ in,
                                                            const float* __restrict__ priors, float* __restrict__ out,
                                                            long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Box6 a = load_box(in + i * 6);
  Box6 r;
  switch (mode) {
    case SSD3D_BOX_CXCYCZ_TO_XYZ: r = cxcycz_to_xyz(a); break;
    case SSD3D_BOX_XYZ_TO_CXCYCZ: r = xyz_to_cxcycz(a); break;
    case SSD3D_BOX_GCXGCYGCZ_TO_CXCYCZ: r = gcxgcygcz_to_cxcycz(a, load_box(priors + i * 6)); break;
    default: r = cxcycz_to_gcxgcygcz(a, load_box(priors + i * 6)); break;
  }
  store_box(out + i * 6, r);
}

// grid: (ceil(n2/128), ceil(n1/8)); block 128 threads = 128 columns; each block walks 8 rows
__global__ void __launch_bounds__(128) iou_pairwise_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                           float* __restrict__ out, long long n1, long long n2,
                                                           int want_iou) {
  const long long j = (long long)blockIdx.x * 128 + threadIdx.x;
  const long long i0 = (long long)blockIdx.y * 8;
  Box6 bj;
  float vb = 0.f;
  if (j < n2) {
    bj = load_box(b + j * 6);
    vb = box_volume(bj);
  }
  __shared__ float ra[8][6];
  if (threadIdx.x < 48) {
    const int r = threadIdx.x / 6, k = threadIdx.x % 6;
    if (i0 + r < n1) ra[r][k] = a[(i0 + r) * 6 + k];
  }
  __syncthreads();
  if (j >= n2) return;
  for (int r = 0; r < 8 && i0 + r < n1; ++r) {
    Box6 ai;
#pragma unroll
    for (int k = 0; k < 6; ++k) ai.v[k] = ra[r][k];
    const float v = want_iou ? box_iou(ai, box_volume(ai), bj, vb) : box_intersection(ai, bj);
    out[(i0 + r) * n2 + j] = v;
  }
}

}  // namespace ssd3d

using namespace ssd3d;

extern "C" int ssd3d_box_transform(int mode, const float* in, const float* priors, float* out, int64_t n,
                                   void* stream) {
  if (!in || !out || n < 0 || mode < 0 || mode > 3) return SSD3D_ERR_ARG;
  if (mode >= 2 && !priors) return SSD3D_ERR_ARG;
  if (n == 0) return SSD3D_OK;
  box_transform_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(mode, in, priors,
                                                                                                  out, n);
  SSD3D_CHECK_LAUNCH();
  return SSD3D_OK;
}

extern "C" int ssd3d_iou3d_pairwise(const float* a, const float* b, float* out, int64_t n1, int64_t n2, int want_iou,
                                    void* stream) {
  if (!a || !b || !out || n1 < 0 || n2 < 0) return SSD3D_ERR_ARG;
  if (n1 == 0 || n2 == 0) return SSD3D_OK;
  const long long gy = (n1 + 7) / 8;
  if (gy > 65535) return SSD3D_ERR_UNSUPPORTED;
  dim3 grid((unsigned)((n2 + 127) / 128), (unsigned)gy);
  iou_pairwise_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(a, b, out, n1, n2, want_iou);
  SSD3D_CHECK_LAUNCH();
  return SSD3D_OK;
}
