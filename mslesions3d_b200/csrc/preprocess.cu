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

// ------------------------------------------------------------------------------------------------
// Ground-truth boxes from a segmentation volume (utils.py:438-513, BoundingBoxesGeneratord "binary" / "classes"):
// face-connected components (scipy.ndimage.label, default structure) of each class, one box
// [min index, max index] / image size per component, ordered by (class, first voxel in C order) = scipy's label
// order; components that are one voxel thick along an axis have zero volume and are dropped (utils.py:475-480).
//   init     L[v] = v for class voxels, -1 otherwise                      (one volume-local int per voxel)
//   merge    lock-free union-find with atomicMin towards the smaller index over the -w/-h/-d neighbours
//   compress L[v] = root(v); roots append their key class*V + v to the volume's list
//   rank     rank by counting over the (distinct) keys -> R[root] = position in the reference's order
//   bbox     per voxel min/max into box[rank], one atomic set per (warp, component) via match_any + redux
//   finalize divide by the image size (fp32, round to nearest as torch), zero-volume filter, ordered compaction
// ------------------------------------------------------------------------------------------------
namespace ssd3d {

template <typename T>
__device__ __forceinline__ int seg_class(const T* __restrict__ seg, long long i, int n_classes) {
  const T v = seg[i];
  if (n_classes <= 0) return v != (T)0 ? 1 : 0;           // binary
  const int c = (int)v;
  return ((T)c == v && c >= 1 && c <= n_classes) ? c : 0;
}

__device__ __forceinline__ int cc_find(const int* L, int x) {
  int p = L[x];
  while (p != x) { x = p; p = L[x]; }
  return x;
}

__device__ __forceinline__ void cc_unite(int* L, int a, int b) {
  bool done = false;
  while (!done) {
    a = cc_find(L, a);
    b = cc_find(L, b);
    if (a < b) { const int old = atomicMin(&L[b], a); done = (old == b); b = old; }
    else if (b < a) { const int old = atomicMin(&L[a], b); done = (old == a); a = old; }
    else done = true;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) cc_init_kernel(const T* __restrict__ seg, long long total, long long V,
                                                      int n_classes, int* __restrict__ L) {
  const long long g = (long long)blockIdx.x * 256 + threadIdx.x;
  if (g >= total) return;
  L[g] = seg_class(seg, g, n_classes) ? (int)(g % V) : -1;
}

template <typename T>
__global__ void __launch_bounds__(256) cc_merge_kernel(const T* __restrict__ seg, long long total, int D, int H, int W,
                                                       int n_classes, int* __restrict__ L) {
  const long long g = (long long)blockIdx.x * 256 + threadIdx.x;
  if (g >= total) return;
  const int c = seg_class(seg, g, n_classes);
  if (!c) return;
  const long long V = (long long)D * H * W;
  const int v = (int)(g % V);
  int* Lv = L + (g - v);
  const T* sv = seg + (g - v);
  const int w = v % W, h = (v / W) % H, d = v / (W * H);
  if (w > 0 && seg_class(sv, v - 1, n_classes) == c) cc_unite(Lv, v, v - 1);
  if (h > 0 && seg_class(sv, v - W, n_classes) == c) cc_unite(Lv, v, v - W);
  if (d > 0 && seg_class(sv, v - W * H, n_classes) == c) cc_unite(Lv, v, v - W * H);
}

template <typename T>
__global__ void __launch_bounds__(256) cc_compress_kernel(const T* __restrict__ seg, long long total, long long V,
                                                          int n_classes, int max_boxes, int* __restrict__ L,
                                                          int* __restrict__ R, long long* __restrict__ root_key,
                                                          int* __restrict__ root_count) {
  const long long g = (long long)blockIdx.x * 256 + threadIdx.x;
  if (g >= total) return;
  const int c = seg_class(seg, g, n_classes);
  if (!c) return;
  const int v = (int)(g % V);
  const int n = (int)(g / V);
  int* Lv = L + (g - v);
  const int r = cc_find(Lv, v);
  Lv[v] = r;
  if (r == v) {
    R[g] = -1;
    const int slot = atomicAdd(&root_count[n], 1);
    if (slot < max_boxes) root_key[(long long)n * max_boxes + slot] = (long long)c * V + v;
  }
}

__global__ void __launch_bounds__(256) cc_rank_kernel(const long long* __restrict__ root_key,
                                                      const int* __restrict__ root_count, long long V, int max_boxes,
                                                      int* __restrict__ R, int* __restrict__ ibox,
                                                      int* __restrict__ cls_of_rank) {
  const int n = blockIdx.y;
  const int cnt = min(root_count[n], max_boxes);
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= cnt) return;
  const long long* keys = root_key + (long long)n * max_boxes;
  const long long k = keys[i];
  int rank = 0;
  for (int j = 0; j < cnt; ++j) rank += (keys[j] < k) ? 1 : 0;
  R[(long long)n * V + (k % V)] = rank;
  cls_of_rank[n * max_boxes + rank] = (int)(k / V);
  int* b = ibox + ((long long)n * max_boxes + rank) * 6;
  b[0] = b[1] = b[2] = 0x7fffffff;
  b[3] = b[4] = b[5] = -1;
}

template <typename T>
__global__ void __launch_bounds__(256) cc_bbox_kernel(const T* __restrict__ seg, long long total, int D, int H, int W,
                                                      int n_classes, int max_boxes, const int* __restrict__ L,
                                                      const int* __restrict__ R, int* __restrict__ ibox) {
  const long long g = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long V = (long long)D * H * W;
  int key = -1, d = 0, h = 0, w = 0;
  if (g < total && seg_class(seg, g, n_classes)) {
    const int v = (int)(g % V);
    const int n = (int)(g / V);
    const int rank = R[(g - v) + L[g]];
    if (rank >= 0) key = n * max_boxes + rank;
    w = v % W; h = (v / W) % H; d = v / (W * H);
  }
  const unsigned active = __ballot_sync(0xffffffffu, key >= 0);
  if (key < 0) return;
  const unsigned grp = __match_any_sync(active, key);
  const int lo_d = __reduce_min_sync(grp, d), lo_h = __reduce_min_sync(grp, h), lo_w = __reduce_min_sync(grp, w);
  const int hi_d = __reduce_max_sync(grp, d), hi_h = __reduce_max_sync(grp, h), hi_w = __reduce_max_sync(grp, w);
  if ((int)(threadIdx.x & 31) == __ffs(grp) - 1) {
    int* b = ibox + (long long)key * 6;
    atomicMin(b + 0, lo_d); atomicMin(b + 1, lo_h); atomicMin(b + 2, lo_w);
    atomicMax(b + 3, hi_d); atomicMax(b + 4, hi_h); atomicMax(b + 5, hi_w);
  }
}

__global__ void __launch_bounds__(256) cc_finalize_kernel(const int* __restrict__ ibox,
                                                          const int* __restrict__ cls_of_rank,
                                                          const int* __restrict__ root_count, int D, int H, int W,
                                                          int max_boxes, float* __restrict__ boxes,
                                                          long long* __restrict__ labels, int* __restrict__ counts,
                                                          int* __restrict__ n_components) {
  __shared__ int warp_tot[8];
  __shared__ int base;
  const int n = blockIdx.x;
  const int total = root_count[n];
  const int cnt = min(total, max_boxes);
  if (threadIdx.x == 0) { base = 0; n_components[n] = total; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r0 = 0; r0 < cnt; r0 += 256) {
    const int r = r0 + threadIdx.x;
    int b[6] = {0, 0, 0, 0, 0, 0};
    bool keep = false;
    if (r < cnt) {
#pragma unroll
      for (int j = 0; j < 6; ++j) b[j] = ibox[((long long)n * max_boxes + r) * 6 + j];
      keep = b[3] > b[0] && b[4] > b[1] && b[5] > b[2];      // zero volume <=> one voxel thick along an axis
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    int off = base;
    for (int q = 0; q < warp; ++q) off += warp_tot[q];
    off += __popc(m & ((1u << lane) - 1u));
    if (keep) {
      float* o = boxes + ((long long)n * max_boxes + off) * 6;
      const float dims[3] = {(float)D, (float)H, (float)W};
#pragma unroll
      for (int j = 0; j < 6; ++j) o[j] = __fdiv_rn((float)b[j], dims[j % 3]);
      labels[(long long)n * max_boxes + off] = cls_of_rank[n * max_boxes + r];
    }
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int q = 0; q < 8; ++q) t += warp_tot[q]; base += t; }
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[n] = base;
}

// ------------------------------------------------------------------------------------------------
// "instances" mode (utils.py:439-441,483-513): the volume already holds one integer id per object, and
// `thresholds` [(min_c, max_c)] map id ranges to classes 1, 2, ...  One box per id, [min index, max index] of
// ALL its voxels (an id may be spread over several blobs), ordered by class, then by ascending id (np.unique).
//   mark   table[id] = present                       (table of K = max id + 1 entries per volume)
//   rank   per class, ids in ascending order get consecutive ranks (block scan over the table)
//   bbox   per voxel min/max into box[rank] (same warp aggregation as above), then cc_finalize_kernel
// Ids outside every range are ignored; ranges must not overlap (the host checks).
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ int inst_id(const T* __restrict__ seg, long long i, int K) {
  const T v = seg[i];
  const int c = (int)v;
  return ((T)c == v && c >= 1 && c < K) ? c : 0;
}

template <typename T>
__global__ void __launch_bounds__(256) inst_mark_kernel(const T* __restrict__ seg, long long total, long long V, int K,
                                                        int* __restrict__ table) {
  const long long g = (long long)blockIdx.x * 256 + threadIdx.x;
  if (g >= total) return;
  const int id = inst_id(seg, g, K);
  if (id) table[(g / V) * K + id] = -2;                 // present, not ranked yet (the table starts at -1)
}

__global__ void __launch_bounds__(256) inst_rank_kernel(const int* __restrict__ thr, int n_thr, int K, int max_boxes,
                                                        int* __restrict__ table, int* __restrict__ ibox,
                                                        int* __restrict__ cls_of_rank, int* __restrict__ root_count) {
  __shared__ int warp_tot[8];
  __shared__ int base;
  const int n = blockIdx.x;
  int* tab = table + (long long)n * K;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = 0; c < n_thr; ++c) {
    const int lo = max(thr[2 * c], 1), hi = min(thr[2 * c + 1], K);
    for (int id0 = lo; id0 < hi; id0 += 256) {
      const int id = id0 + threadIdx.x;
      const bool flag = id < hi && tab[id] == -2;
      const unsigned m = __ballot_sync(0xffffffffu, flag);
      if (lane == 0) warp_tot[warp] = __popc(m);
      __syncthreads();
      int off = base;
      for (int q = 0; q < warp; ++q) off += warp_tot[q];
      off += __popc(m & ((1u << lane) - 1u));
      if (flag) {
        if (off < max_boxes) {
          tab[id] = off;
          cls_of_rank[n * max_boxes + off] = c + 1;
          int* b = ibox + ((long long)n * max_boxes + off) * 6;
          b[0] = b[1] = b[2] = 0x7fffffff;
          b[3] = b[4] = b[5] = -1;
        } else {
          tab[id] = -1;
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) { int t = 0; for (int q = 0; q < 8; ++q) t += warp_tot[q]; base += t; }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) root_count[n] = base;
}

template <typename T>
__global__ void __launch_bounds__(256) inst_bbox_kernel(const T* __restrict__ seg, long long total, int D, int H, int W,
                                                        int K, int max_boxes, const int* __restrict__ table,
                                                        int* __restrict__ ibox) {
  const long long g = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long V = (long long)D * H * W;
  int key = -1, d = 0, h = 0, w = 0;
  if (g < total) {
    const int id = inst_id(seg, g, K);
    if (id) {
      const int v = (int)(g % V);
      const int n = (int)(g / V);
      const int rank = table[(long long)n * K + id];
      if (rank >= 0) key = n * max_boxes + rank;
      w = v % W; h = (v / W) % H; d = v / (W * H);
    }
  }
  const unsigned active = __ballot_sync(0xffffffffu, key >= 0);
  if (key < 0) return;
  const unsigned grp = __match_any_sync(active, key);
  const int lo_d = __reduce_min_sync(grp, d), lo_h = __reduce_min_sync(grp, h), lo_w = __reduce_min_sync(grp, w);
  const int hi_d = __reduce_max_sync(grp, d), hi_h = __reduce_max_sync(grp, h), hi_w = __reduce_max_sync(grp, w);
  if ((int)(threadIdx.x & 31) == __ffs(grp) - 1) {
    int* b = ibox + (long long)key * 6;
    atomicMin(b + 0, lo_d); atomicMin(b + 1, lo_h); atomicMin(b + 2, lo_w);
    atomicMax(b + 3, hi_d); atomicMax(b + 4, hi_h); atomicMax(b + 5, hi_w);
  }
}

struct CcWs {
  int* L; int* R; long long* root_key; int* root_count; int* ibox; int* cls_of_rank;
};

static inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

static CcWs cc_carve(void* ws, int N, long long V, int max_boxes) {
  uint8_t* p = static_cast<uint8_t*>(ws);
  CcWs c;
  c.L = reinterpret_cast<int*>(p); p += al256((size_t)N * V * 4);
  c.R = reinterpret_cast<int*>(p); p += al256((size_t)N * V * 4);
  c.root_key = reinterpret_cast<long long*>(p); p += al256((size_t)N * max_boxes * 8);
  c.root_count = reinterpret_cast<int*>(p); p += al256((size_t)N * 4);
  c.ibox = reinterpret_cast<int*>(p); p += al256((size_t)N * max_boxes * 6 * 4);
  c.cls_of_rank = reinterpret_cast<int*>(p);
  return c;
}

template <typename T>
static int run_gt_boxes(const T* seg, int N, int D, int H, int W, int n_classes, int max_boxes, float* boxes,
                        long long* labels, int* counts, int* n_components, void* ws, cudaStream_t st) {
  const long long V = (long long)D * H * W, total = V * N;
  CcWs c = cc_carve(ws, N, V, max_boxes);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  cudaError_t e = cudaMemsetAsync(c.root_count, 0, (size_t)N * 4, st);
  if (e != cudaSuccess) return (int)e;
  cc_init_kernel<T><<<blocks, 256, 0, st>>>(seg, total, V, n_classes, c.L);
  SSD3D_CHECK_LAUNCH();
  cc_merge_kernel<T><<<blocks, 256, 0, st>>>(seg, total, D, H, W, n_classes, c.L);
  SSD3D_CHECK_LAUNCH();
  cc_compress_kernel<T><<<blocks, 256, 0, st>>>(seg, total, V, n_classes, max_boxes, c.L, c.R, c.root_key, c.root_count);
  SSD3D_CHECK_LAUNCH();
  cc_rank_kernel<<<dim3((unsigned)((max_boxes + 255) / 256), (unsigned)N), 256, 0, st>>>(c.root_key, c.root_count, V,
                                                                                         max_boxes, c.R, c.ibox,
                                                                                         c.cls_of_rank);
  SSD3D_CHECK_LAUNCH();
  cc_bbox_kernel<T><<<blocks, 256, 0, st>>>(seg, total, D, H, W, n_classes, max_boxes, c.L, c.R, c.ibox);
  SSD3D_CHECK_LAUNCH();
  cc_finalize_kernel<<<(unsigned)N, 256, 0, st>>>(c.ibox, c.cls_of_rank, c.root_count, D, H, W, max_boxes, boxes, labels,
                                                  counts, n_components);
  SSD3D_CHECK_LAUNCH();
  return SSD3D_OK;
}

struct InstWs {
  int* table; int* root_count; int* ibox; int* cls_of_rank;
};
static InstWs inst_carve(void* ws, int N, int K, int max_boxes) {
  uint8_t* p = static_cast<uint8_t*>(ws);
  InstWs c;
  c.table = reinterpret_cast<int*>(p); p += al256((size_t)N * K * 4);
  c.root_count = reinterpret_cast<int*>(p); p += al256((size_t)N * 4);
  c.ibox = reinterpret_cast<int*>(p); p += al256((size_t)N * max_boxes * 6 * 4);
  c.cls_of_rank = reinterpret_cast<int*>(p);
  return c;
}

template <typename T>
static int run_gt_boxes_instances(const T* seg, int N, int D, int H, int W, const int* thr, int n_thr, int K,
                                  int max_boxes, float* boxes, long long* labels, int* counts, int* n_components,
                                  void* ws, cudaStream_t st) {
  const long long V = (long long)D * H * W, total = V * N;
  InstWs c = inst_carve(ws, N, K, max_boxes);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  cudaError_t e = cudaMemsetAsync(c.table, 0xff, (size_t)N * K * 4, st);
  if (e != cudaSuccess) return (int)e;
  inst_mark_kernel<T><<<blocks, 256, 0, st>>>(seg, total, V, K, c.table);
  SSD3D_CHECK_LAUNCH();
  inst_rank_kernel<<<(unsigned)N, 256, 0, st>>>(thr, n_thr, K, max_boxes, c.table, c.ibox, c.cls_of_rank, c.root_count);
  SSD3D_CHECK_LAUNCH();
  inst_bbox_kernel<T><<<blocks, 256, 0, st>>>(seg, total, D, H, W, K, max_boxes, c.table, c.ibox);
  SSD3D_CHECK_LAUNCH();
  cc_finalize_kernel<<<(unsigned)N, 256, 0, st>>>(c.ibox, c.cls_of_rank, c.root_count, D, H, W, max_boxes, boxes, labels,
                                                  counts, n_components);
  SSD3D_CHECK_LAUNCH();
  return SSD3D_OK;
}

}  // namespace ssd3d

extern "C" int64_t ssd3d_gt_boxes_instances_workspace_bytes(int N, int id_limit, int max_boxes) {
  if (N <= 0 || id_limit <= 0 || max_boxes <= 0) return 0;
  return (int64_t)(ssd3d::al256((size_t)N * id_limit * 4) + ssd3d::al256((size_t)N * 4) +
                   ssd3d::al256((size_t)N * max_boxes * 24) + ssd3d::al256((size_t)N * max_boxes * 4));
}

extern "C" int ssd3d_gt_boxes_from_instances(const void* seg, int seg_dtype, int N, int D, int H, int W,
                                             const int32_t* thresholds, int n_thresholds, int id_limit, int max_boxes,
                                             float* boxes, int64_t* labels, int32_t* counts, int32_t* n_components,
                                             void* workspace, int64_t workspace_bytes, void* stream) {
  if (!seg || !thresholds || !boxes || !labels || !counts || !n_components || !workspace) return SSD3D_ERR_ARG;
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0 || max_boxes <= 0 || n_thresholds <= 0 || id_limit <= 1) return SSD3D_ERR_ARG;
  const long long V = (long long)D * H * W;
  if (V >= (1ll << 31) || (long long)N * max_boxes >= (1ll << 31) || (long long)N * id_limit >= (1ll << 31))
    return SSD3D_ERR_UNSUPPORTED;
  if (workspace_bytes < ssd3d_gt_boxes_instances_workspace_bytes(N, id_limit, max_boxes)) return SSD3D_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long* lab = reinterpret_cast<long long*>(labels);
  if (seg_dtype == 0)
    return ssd3d::run_gt_boxes_instances<uint8_t>(static_cast<const uint8_t*>(seg), N, D, H, W, thresholds, n_thresholds,
                                                  id_limit, max_boxes, boxes, lab, counts, n_components, workspace, st);
  if (seg_dtype == 1)
    return ssd3d::run_gt_boxes_instances<float>(static_cast<const float*>(seg), N, D, H, W, thresholds, n_thresholds,
                                                id_limit, max_boxes, boxes, lab, counts, n_components, workspace, st);
  return SSD3D_ERR_ARG;
}

extern "C" int64_t ssd3d_gt_boxes_workspace_bytes(int N, int D, int H, int W, int max_boxes) {
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0 || max_boxes <= 0) return 0;
  const long long V = (long long)D * H * W;
  return (int64_t)(2 * ssd3d::al256((size_t)N * V * 4) + ssd3d::al256((size_t)N * max_boxes * 8) +
                   ssd3d::al256((size_t)N * 4) + ssd3d::al256((size_t)N * max_boxes * 24) +
                   ssd3d::al256((size_t)N * max_boxes * 4));
}

extern "C" int ssd3d_gt_boxes_from_segmentation(const void* seg, int seg_dtype, int N, int D, int H, int W,
                                                int n_classes, int max_boxes, float* boxes, int64_t* labels,
                                                int32_t* counts, int32_t* n_components, void* workspace,
                                                int64_t workspace_bytes, void* stream) {
  if (!seg || !boxes || !labels || !counts || !n_components || !workspace) return SSD3D_ERR_ARG;
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0 || max_boxes <= 0 || n_classes < 0) return SSD3D_ERR_ARG;
  const long long V = (long long)D * H * W;
  if (V >= (1ll << 31)) return SSD3D_ERR_UNSUPPORTED;      // volume-local voxel indices are 32-bit
  if ((long long)N * max_boxes >= (1ll << 31)) return SSD3D_ERR_UNSUPPORTED;
  if (workspace_bytes < ssd3d_gt_boxes_workspace_bytes(N, D, H, W, max_boxes)) return SSD3D_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long* lab = reinterpret_cast<long long*>(labels);
  if (seg_dtype == 0)
    return ssd3d::run_gt_boxes<uint8_t>(static_cast<const uint8_t*>(seg), N, D, H, W, n_classes, max_boxes, boxes, lab,
                                        counts, n_components, workspace, st);
  if (seg_dtype == 1)
    return ssd3d::run_gt_boxes<float>(static_cast<const float*>(seg), N, D, H, W, n_classes, max_boxes, boxes, lab, counts,
                                      n_components, workspace, st);
  return SSD3D_ERR_ARG;
}
