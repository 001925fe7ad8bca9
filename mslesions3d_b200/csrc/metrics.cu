// Detection metrics of one class (utils.py:155-230 compute_metrics_per_class, and the cumulative precision /
// recall / 11-point average precision of utils.py:296-318): the reference walks the score-sorted detections in
// a Python loop, one IoU call and several host syncs per detection.  The only sequential dependency is "has
// this ground-truth object been claimed by an earlier detection", i.e. per object the FIRST detection (in
// sorted order) whose best overlap is that object -- an atomicMin over sorted ranks.  Everything is integer
// exact: IoU with the reference's separately rounded fp32 steps (boxes.cuh), first-maximum tie rule of
// torch.max, strict `> min_overlap`, and a stable descending order for equal scores (the reference's sort
// leaves ties unspecified).
//   map_best_gt   : per detection, first-max IoU over the class's objects of the same image
//   map_rank      : stable descending rank by counting (n^2 compares: n <= a few 10^4 detections)
//   map_claim     : atomicMin(first_rank[object], rank) for detections above the overlap threshold
//   map_flags     : TP / FP per sorted position, detected flag and volume per object
//   map_ap        : one block: inclusive scans of TP / FP, precision at the 11 recall thresholds, AP,
//                   recall / precision / F1 of the class
#include "boxes.cuh"
#include "../../include/ssd3d_b200.h"

namespace ssd3d {

__global__ void __launch_bounds__(256) map_best_gt_kernel(const float* __restrict__ det_boxes,
                                                          const int* __restrict__ det_images, int nd,
                                                          const float* __restrict__ true_boxes,
                                                          const int* __restrict__ true_images, int nt,
                                                          int* __restrict__ best_gt, float* __restrict__ best_iou) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nd) return;
  const Box6 db = load_box(det_boxes + (size_t)d * 6);
  const float vd = box_volume(db);
  const int img = det_images[d];
  int best = -1;
  float bi = 0.f;
  for (int t = 0; t < nt; ++t) {
    if (true_images[t] != img) continue;
    const Box6 tb = load_box(true_boxes + (size_t)t * 6);
    const float iou = box_iou(db, vd, tb, box_volume(tb));
    // torch.max over the image's objects in their original order: first maximum; NaN wins (propagates)
    if (best < 0 || iou > bi || (iou != iou && bi == bi)) { best = t; bi = iou; }
  }
  best_gt[d] = best;
  best_iou[d] = bi;
}

__global__ void __launch_bounds__(256) map_rank_kernel(const float* __restrict__ scores, int nd,
                                                       int* __restrict__ rank, float* __restrict__ sorted_scores,
                                                       int* __restrict__ sort_index) {
  __shared__ float tile[256];
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const float s = d < nd ? scores[d] : 0.f;
  int r = 0;
  for (int j0 = 0; j0 < nd; j0 += 256) {
    __syncthreads();
    if (j0 + threadIdx.x < nd) tile[threadIdx.x] = scores[j0 + threadIdx.x];
    __syncthreads();
    const int lim = min(256, nd - j0);
    for (int j = 0; j < lim; ++j) {
      const float o = tile[j];
      r += (o > s) || (o == s && (j0 + j) < d);
    }
  }
  if (d < nd) {
    rank[d] = r;
    sorted_scores[r] = s;
    sort_index[r] = d;
  }
}

__global__ void __launch_bounds__(256) map_claim_kernel(const int* __restrict__ best_gt,
                                                        const float* __restrict__ best_iou,
                                                        const int* __restrict__ rank,
                                                        const unsigned char* __restrict__ difficult, int nd,
                                                        float min_overlap, unsigned int* __restrict__ first_rank) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nd) return;
  const int g = best_gt[d];
  if (g >= 0 && best_iou[d] > min_overlap && !difficult[g]) atomicMin(first_rank + g, (unsigned int)rank[d]);
}

__global__ void __launch_bounds__(256) map_flags_kernel(const int* __restrict__ best_gt,
                                                        const float* __restrict__ best_iou,
                                                        const int* __restrict__ rank,
                                                        const unsigned char* __restrict__ difficult, int nd,
                                                        float min_overlap, const unsigned int* __restrict__ first_rank,
                                                        const float* __restrict__ true_boxes, int nt,
                                                        float* __restrict__ tp, float* __restrict__ fp,
                                                        unsigned char* __restrict__ detected,
                                                        float* __restrict__ volumes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nd) {
    const int g = best_gt[i], r = rank[i];
    float t = 0.f, f = 0.f;
    if (g < 0) f = 1.f;                                  // no object of this class in the image
    else if (best_iou[i] > min_overlap) {
      if (!difficult[g]) { if (first_rank[g] == (unsigned int)r) t = 1.f; else f = 1.f; }
    } else f = 1.f;
    tp[r] = t;
    fp[r] = f;
  }
  if (i < nt) {
    detected[i] = first_rank[i] != 0xffffffffu;
    volumes[i] = box_volume(load_box(true_boxes + (size_t)i * 6));    // utils.py:150-152
  }
}

// block-wide inclusive scan of (a, b) over n elements in chunks of 1024 (exact: the values are small integers)
__global__ void __launch_bounds__(1024) map_ap_kernel(const float* __restrict__ tp, const float* __restrict__ fp,
                                                      int nd, int nt, const unsigned char* __restrict__ difficult,
                                                      const float* __restrict__ thresholds,
                                                      int n_thr, float* __restrict__ cum_precision,
                                                      float* __restrict__ cum_recall,
                                                      const unsigned char* __restrict__ detected,
                                                      float* __restrict__ out /* AP, recall, precision, F1, 11 precisions */) {
  __shared__ float wa[32], wb[32];
  __shared__ float carry_a, carry_b;
  __shared__ unsigned int best[16];
  __shared__ int n_det, n_easy_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { carry_a = 0.f; carry_b = 0.f; n_det = 0; n_easy_s = 0; }
  if (tid < 16) best[tid] = 0u;
  __syncthreads();
  {
    int easy = 0;
    for (int t = tid; t < nt; t += 1024) easy += difficult[t] ? 0 : 1;     // utils.py:270
    if (easy) atomicAdd(&n_easy_s, easy);
  }
  __syncthreads();
  const int n_easy = n_easy_s;
  for (int i0 = 0; i0 < nd; i0 += 1024) {
    const int i = i0 + tid;
    float a = i < nd ? tp[i] : 0.f, b = i < nd ? fp[i] : 0.f;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const float ua = __shfl_up_sync(0xffffffffu, a, s), ub = __shfl_up_sync(0xffffffffu, b, s);
      if (lane >= s) { a += ua; b += ub; }
    }
    if (lane == 31) { wa[warp] = a; wb[warp] = b; }
    __syncthreads();
    if (warp == 0) {
      float x = wa[lane], y = wb[lane];
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        const float ux = __shfl_up_sync(0xffffffffu, x, s), uy = __shfl_up_sync(0xffffffffu, y, s);
        if (lane >= s) { x += ux; y += uy; }
      }
      wa[lane] = x; wb[lane] = y;
    }
    __syncthreads();
    const float ca = carry_a + (warp ? wa[warp - 1] : 0.f) + a;
    const float cb = carry_b + (warp ? wb[warp - 1] : 0.f) + b;
    if (i < nd) {
      // utils.py:299-300
      const float prec = __fdiv_rn(ca, __fadd_rn(__fadd_rn(ca, cb), 1e-10f));
      const float rec = __fdiv_rn(ca, (float)n_easy);
      cum_precision[i] = prec;
      cum_recall[i] = rec;
      for (int k = 0; k < n_thr; ++k)
        if (rec >= thresholds[k]) atomicMax(&best[k], __float_as_uint(prec));   // prec >= 0: uint order == float order
    }
    __syncthreads();
    if (tid == 1023) { carry_a = ca; carry_b = cb; }
    __syncthreads();
  }
  int local = 0;
  for (int t = tid; t < nt; t += 1024) local += detected[t] ? 1 : 0;
  if (local) atomicAdd(&n_det, local);
  __syncthreads();
  if (tid == 0) {
    float sum = 0.f;
    for (int k = 0; k < n_thr; ++k) {
      const float pk = __uint_as_float(best[k]);
      out[4 + k] = pk;
      sum = __fadd_rn(sum, pk);
    }
    out[0] = __fdiv_rn(sum, (float)n_thr);
    const float tps = carry_a, fps = carry_b, fns = (float)(nt - n_det);
    const float rec = __fdiv_rn(tps, __fadd_rn(tps, fns));
    const float pre = __fdiv_rn(tps, __fadd_rn(tps, fps));
    out[1] = rec;
    out[2] = pre;
    out[3] = __fdiv_rn(__fmul_rn(__fmul_rn(2.f, pre), rec), __fadd_rn(pre, rec));
  }
}

}  // namespace ssd3d

using namespace ssd3d;

extern "C" int64_t ssd3d_map_workspace_bytes(int64_t nd, int64_t nt) {
  return (int64_t)(nd * (4 + 4 + 4) + nt * 4 + 256);
}

extern "C" int ssd3d_map_class(const float* det_boxes, const float* det_scores, const int32_t* det_images, int64_t nd,
                               const float* true_boxes, const uint8_t* true_difficulties, const int32_t* true_images,
                               int64_t nt, float min_overlap, const float* recall_thresholds, int n_thresholds,
                               float* sorted_scores, int32_t* sort_index, float* true_positives,
                               float* false_positives, uint8_t* detected, float* true_volumes, float* cum_precision,
                               float* cum_recall, float* out_stats, void* workspace, int64_t workspace_bytes,
                               void* stream) {
  if (nd <= 0 || nt < 0 || nd > 0x3fffffff || nt > 0x3fffffff || n_thresholds < 1 || n_thresholds > 16) return SSD3D_ERR_ARG;
  if (!det_boxes || !det_scores || !det_images || !recall_thresholds || !sorted_scores || !sort_index ||
      !true_positives || !false_positives || !cum_precision || !cum_recall || !out_stats || !workspace)
    return SSD3D_ERR_ARG;
  if (nt > 0 && (!true_boxes || !true_difficulties || !true_images || !detected || !true_volumes)) return SSD3D_ERR_ARG;
  if (workspace_bytes < ssd3d_map_workspace_bytes(nd, nt)) return SSD3D_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int* best_gt = static_cast<int*>(workspace);
  float* best_iou = reinterpret_cast<float*>(best_gt + nd);
  int* rank = reinterpret_cast<int*>(best_iou + nd);
  unsigned int* first_rank = reinterpret_cast<unsigned int*>(rank + nd);
  const int n = (int)nd, m = (int)nt;
  const unsigned gd = (unsigned)((n + 255) / 256);
  if (m > 0) {
    cudaError_t e = cudaMemsetAsync(first_rank, 0xff, (size_t)m * 4, st);     // 0xffffffff = not claimed
    if (e != cudaSuccess) return (int)e;
  }
  map_best_gt_kernel<<<gd, 256, 0, st>>>(det_boxes, det_images, n, true_boxes, true_images, m, best_gt, best_iou);
  SSD3D_CHECK_LAUNCH();
  map_rank_kernel<<<gd, 256, 0, st>>>(det_scores, n, rank, sorted_scores, sort_index);
  SSD3D_CHECK_LAUNCH();
  if (m > 0) {
    map_claim_kernel<<<gd, 256, 0, st>>>(best_gt, best_iou, rank, true_difficulties, n, min_overlap, first_rank);
    SSD3D_CHECK_LAUNCH();
  }
  const int nmax = n > m ? n : m;
  map_flags_kernel<<<(unsigned)((nmax + 255) / 256), 256, 0, st>>>(best_gt, best_iou, rank, true_difficulties, n,
                                                                   min_overlap, first_rank, true_boxes, m,
                                                                   true_positives, false_positives, detected,
                                                                   true_volumes);
  SSD3D_CHECK_LAUNCH();
  map_ap_kernel<<<1, 1024, 0, st>>>(true_positives, false_positives, n, m, true_difficulties, recall_thresholds,
                                    n_thresholds, cum_precision, cum_recall, detected, out_stats);
  SSD3D_CHECK_LAUNCH();
  return SSD3D_OK;
}
