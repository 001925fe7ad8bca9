// Training-side kernels of the SSD3D path (MultiBoxLoss, ssd3d.py:741-941):
//   match_iou     : IoU of every (object, prior) pair computed on the fly; per prior the first-max object
//                   (ssd3d.py:798-803,829-836), per object the first-max prior through a warp-shuffle
//                   reduction + 64-bit atomicMax on {iou bits, ~prior index} (ssd3d.py:811)
//   match_assign  : force-match with last-writer-wins (ssd3d.py:865-868), labels + hard/soft threshold
//                   (ssd3d.py:871-881), target encoding for every prior (ssd3d.py:887)
//   multibox_*    : cross-entropy over all non-ignored priors + L1 over positives, normalised by the
//                   number of positives (ssd3d.py:891-933), with the analytic gradient w.r.t. the head
//                   outputs; optional hard-negative mining (the commented ssd3d.py:926-932) by radix select.
#include "boxes.cuh"

namespace ssd3d {

constexpr int OBJ_CHUNK = 64;

__global__ void __launch_bounds__(256) match_iou_kernel(const float* __restrict__ gt_boxes,
                                                        const int* __restrict__ gt_offsets,
                                                        const PriorSrc priors, long long P,
                                                        float* __restrict__ overlap, int* __restrict__ obj_for_prior,
                                                        unsigned long long* __restrict__ best_key) {
  const int img = blockIdx.y;
  const int g0 = gt_offsets[img], g1 = gt_offsets[img + 1];
  const int n_obj = g1 - g0;
  if (n_obj <= 0) return;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = p < P;
  const int lane = threadIdx.x & 31;
  Box6 pr;
  float vp = 0.f;
  if (live) {
    pr = cxcycz_to_xyz(load_prior(priors, p));   // MultiBoxLoss.priors_xyz, ssd3d.py:753
    vp = box_volume(pr);
  }
  __shared__ float ob[OBJ_CHUNK][6];
  __shared__ float ov[OBJ_CHUNK];
  float best = 0.f;
  int best_o = 0;
  for (int c0 = 0; c0 < n_obj; c0 += OBJ_CHUNK) {
    const int cn = min(OBJ_CHUNK, n_obj - c0);
    __syncthreads();
    if (threadIdx.x < cn) {
      const Box6 b = load_box(gt_boxes + (long long)(g0 + c0 + threadIdx.x) * 6);
#pragma unroll
      for (int k = 0; k < 6; ++k) ob[threadIdx.x][k] = b.v[k];
      ov[threadIdx.x] = box_volume(b);
    }
    __syncthreads();
    for (int o = 0; o < cn; ++o) {
      Box6 b;
#pragma unroll
      for (int k = 0; k < 6; ++k) b.v[k] = ob[o][k];
      float iou = 0.f;
      unsigned long long key = 0ull;
      if (live) {
        iou = box_iou(b, ov[o], pr, vp);
        if (c0 + o == 0 || iou > best) { best = iou; best_o = c0 + o; }   // first maximum wins
        key = ((unsigned long long)__float_as_uint(iou) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)p);
      }
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, s);
        key = other > key ? other : key;
      }
      if (lane == 0) atomicMax(&best_key[g0 + c0 + o], key);
    }
  }
  if (live) {
    overlap[(long long)img * P + p] = best;
    obj_for_prior[(long long)img * P + p] = best_o;
  }
}

__global__ void __launch_bounds__(256) match_assign_kernel(const float* __restrict__ gt_boxes,
                                                           const long long* __restrict__ gt_labels,
                                                           const int* __restrict__ gt_offsets,
                                                           const PriorSrc priors, long long P, float t0,
                                                           float t1, const unsigned long long* __restrict__ best_key,
                                                           float* __restrict__ overlap, int* __restrict__ obj_for_prior,
                                                           int* __restrict__ prior_for_obj,
                                                           long long* __restrict__ true_classes,
                                                           float* __restrict__ true_locs) {
  const int img = blockIdx.y;
  const int g0 = gt_offsets[img], g1 = gt_offsets[img + 1];
  const int n_obj = g1 - g0;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = p < P;
  const long long ip = (long long)img * P + (live ? p : 0);
  if (n_obj <= 0) {   // ssd3d.py:854-855: the image keeps its zero-initialised targets
    if (live) {
      true_classes[ip] = 0;
#pragma unroll
      for (int k = 0; k < 6; ++k) true_locs[ip * 6 + k] = 0.f;
      overlap[ip] = 0.f;
      obj_for_prior[ip] = 0;
    }
    return;
  }
  int o = 0;
  float ovl = 0.f;
  if (live) {
    o = obj_for_prior[ip];
    ovl = overlap[ip];
  }
  __shared__ uint32_t pfo[OBJ_CHUNK];
  for (int c0 = 0; c0 < n_obj; c0 += OBJ_CHUNK) {
    const int cn = min(OBJ_CHUNK, n_obj - c0);
    __syncthreads();
    if (threadIdx.x < cn) {
      const uint32_t pp = 0xffffffffu - (uint32_t)(best_key[g0 + c0 + threadIdx.x] & 0xffffffffull);
      pfo[threadIdx.x] = pp;
      if (blockIdx.x == 0) prior_for_obj[g0 + c0 + threadIdx.x] = (int)pp;
    }
    __syncthreads();
    if (live) {
      for (int k = 0; k < cn; ++k)
        if (pfo[k] == (uint32_t)p) { o = c0 + k; ovl = 1.0f; }   // ascending objects: the last one wins
    }
  }
  if (!live) return;
  long long label = gt_labels[g0 + o];
  if (ovl < t0) label = 0;
  else if (ovl < t1) label = -1;
  true_classes[ip] = label;
  overlap[ip] = ovl;
  obj_for_prior[ip] = o;
  const Box6 enc = cxcycz_to_gcxgcygcz(xyz_to_cxcycz(load_box(gt_boxes + (long long)(g0 + o) * 6)),
                                       load_prior(priors, p));
  store_box(true_locs + ip * 6, enc);
}

// ------------------------------------------------------------------------------------------------
// MultiBox loss
// ------------------------------------------------------------------------------------------------
struct LossAcc {        // lives at the start of the workspace
  double sum_ce_neg;    // CE over non-positive, non-ignored priors
  double sum_ce_pos;    // CE over positives
  double sum_l1;        // sum |pred - target| over positives (6 coords each)
  double sum_hard;      // HNM: CE over the selected hard negatives
  int n_pos;
  int pad;
};

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0.0;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) t += __shfl_xor_sync(0xffffffffu, t, s);
  }
  return t;  // valid in thread 0
}

// per-prior cross entropy -log_softmax(scores)[target] in torch's formulation
__device__ __forceinline__ float cross_entropy_row(const float* s, int C, int target) {
  float m = s[0];
  for (int k = 1; k < C; ++k) m = fmaxf(m, s[k]);
  float sum = 0.f;
  for (int k = 0; k < C; ++k) sum = __fadd_rn(sum, expf(__fsub_rn(s[k], m)));
  return -__fsub_rn(__fsub_rn(s[target], m), logf(sum));
}

__global__ void __launch_bounds__(256) multibox_reduce_kernel(const float* __restrict__ locs,
                                                              const float* __restrict__ scores,
                                                              const long long* __restrict__ tc,
                                                              const float* __restrict__ tl, long long P, int C,
                                                              LossAcc* __restrict__ acc, int* __restrict__ n_pos_img,
                                                              float* __restrict__ ce_neg) {
  __shared__ double sh[8];
  const int img = blockIdx.y;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double ce_n = 0.0, ce_p = 0.0, l1 = 0.0;
  int pos = 0;
  if (p < P) {
    const long long ip = (long long)img * P + p;
    const long long t = tc[ip];
    float ce = cross_entropy_row(scores + ip * C, C, (int)(t > 0 ? t : 0));
    if (t < 0) ce = 0.f;
    if (t > 0) {
      pos = 1;
      ce_p = ce;
#pragma unroll
      for (int k = 0; k < 6; ++k) l1 += (double)fabsf(__fsub_rn(locs[ip * 6 + k], tl[ip * 6 + k]));
    } else {
      ce_n = ce;
    }
    if (ce_neg) ce_neg[ip] = (t > 0) ? 0.f : ce;
  }
  const double a = block_sum(ce_n, sh);
  const double b = block_sum(ce_p, sh);
  const double c = block_sum(l1, sh);
  const int np = __syncthreads_count(pos);
  if (threadIdx.x == 0) {
    if (a != 0.0) atomicAdd(&acc->sum_ce_neg, a);
    if (b != 0.0) atomicAdd(&acc->sum_ce_pos, b);
    if (c != 0.0) atomicAdd(&acc->sum_l1, c);
    if (np) {
      atomicAdd(&acc->n_pos, np);
      atomicAdd(&n_pos_img[img], np);
    }
  }
}

// Hard-negative mining for one image: the k = ratio * n_pos largest entries of ce_neg (all >= 0, so the
// raw float bits order them).  MSB-first radix select finds the k-th largest key; `sel` marks the chosen
// priors (ties at the threshold broken by ascending prior index).
__global__ void __launch_bounds__(1024) hnm_select_kernel(const float* __restrict__ ce_neg,
                                                          const int* __restrict__ n_pos_img, long long P, int ratio,
                                                          uint8_t* __restrict__ sel, LossAcc* __restrict__ acc) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_remaining;
  __shared__ double sh[32];
  __shared__ int s_run;
  const int img = blockIdx.x;
  const float* v = ce_neg + (long long)img * P;
  uint8_t* out = sel + (long long)img * P;
  long long k = (long long)ratio * n_pos_img[img];
  if (k > P) k = P;
  if (k <= 0) {
    for (long long i = threadIdx.x; i < P; i += blockDim.x) out[i] = 0;
    return;
  }
  if (threadIdx.x == 0) { s_prefix = 0u; s_remaining = (unsigned int)k; }
  __syncthreads();
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    const unsigned int prefix = s_prefix;
    const unsigned int himask = (shift == 24) ? 0u : (0xffffffffu << (shift + 8));
    for (long long i = threadIdx.x; i < P; i += blockDim.x) {
      const unsigned int key = __float_as_uint(v[i]);
      if ((key & himask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int rem = s_remaining;
      int b = 255;
      for (; b > 0; --b) {
        if (hist[b] >= rem) break;
        rem -= hist[b];
      }
      s_prefix = prefix | ((unsigned int)b << shift);
      s_remaining = rem;   // how many entries equal to the final key are still needed
    }
    __syncthreads();
  }
  const unsigned int thr = s_prefix;
  const unsigned int ties_needed = s_remaining;
  if (threadIdx.x == 0) s_run = 0;
  __syncthreads();
  double local = 0.0;
  for (long long base = 0; base < P; base += blockDim.x) {
    const long long i = base + threadIdx.x;
    const unsigned int key = (i < P) ? __float_as_uint(v[i]) : 0u;
    const bool tie = (i < P) && key == thr;
    // ordered rank of the ties inside this 1024-wide slab
    const unsigned ballot = __ballot_sync(0xffffffffu, tie);
    __shared__ int wcount[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) wcount[warp] = __popc(ballot);
    __syncthreads();
    int before = s_run;
    for (int w = 0; w < warp; ++w) before += wcount[w];
    const int rank = before + __popc(ballot & ((1u << lane) - 1u));
    bool chosen = false;
    if (i < P) {
      chosen = (key > thr) || (tie && (unsigned)rank < ties_needed);
      out[i] = chosen ? 1 : 0;
      if (chosen) local += (double)v[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += wcount[w];
      s_run += t;
    }
    __syncthreads();
  }
  const double tot = block_sum(local, sh);
  if (threadIdx.x == 0 && tot != 0.0) atomicAdd(&acc->sum_hard, tot);
}

__global__ void multibox_finalize_kernel(const LossAcc* __restrict__ acc, int hnm, float* __restrict__ out_loss,
                                         int* __restrict__ n_pos_out) {
  const float n_pos = (float)acc->n_pos;
  const float neg = (float)(hnm ? acc->sum_hard : acc->sum_ce_neg);
  const float pos = (float)acc->sum_ce_pos;
  out_loss[0] = __fdiv_rn(__fadd_rn(neg, pos), n_pos);                          // ssd3d.py:933
  out_loss[1] = (float)(acc->sum_l1 / ((double)acc->n_pos * 6.0));              // nn.L1Loss mean, ssd3d.py:896
  if (n_pos_out) *n_pos_out = acc->n_pos;
}

__global__ void __launch_bounds__(256) multibox_grad_kernel(const float* __restrict__ locs,
                                                            const float* __restrict__ scores,
                                                            const long long* __restrict__ tc,
                                                            const float* __restrict__ tl, long long P, int C,
                                                            float alpha, const LossAcc* __restrict__ acc,
                                                            const uint8_t* __restrict__ sel,
                                                            float* __restrict__ grad_locs,
                                                            float* __restrict__ grad_scores) {
  const int img = blockIdx.y;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const long long ip = (long long)img * P + p;
  const long long t = tc[ip];
  const float n_pos = (float)acc->n_pos;
  if (grad_scores) {
    const float* s = scores + ip * C;
    float w = (t < 0) ? 0.f : 1.f;
    if (sel && t == 0) w = sel[ip] ? 1.f : 0.f;
    const int target = (int)(t > 0 ? t : 0);
    float m = s[0];
    for (int k = 1; k < C; ++k) m = fmaxf(m, s[k]);
    float sum = 0.f;
    for (int k = 0; k < C; ++k) sum += expf(s[k] - m);
    const float scale = w / n_pos;
    for (int k = 0; k < C; ++k) {
      const float sm = expf(s[k] - m) / sum;
      grad_scores[ip * C + k] = (sm - (k == target ? 1.f : 0.f)) * scale;
    }
  }
  if (grad_locs) {
    const float g = alpha / (n_pos * 6.0f);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      float v = 0.f;
      if (t > 0) {
        const float d = locs[ip * 6 + k] - tl[ip * 6 + k];
        v = d > 0.f ? g : (d < 0.f ? -g : 0.f);
      }
      grad_locs[ip * 6 + k] = v;
    }
  }
}

static inline long long align256(long long v) { return (v + 255) & ~255ll; }

}  // namespace ssd3d

using namespace ssd3d;

static int match_priors_impl(const float* gt_boxes, const int64_t* gt_labels, const int32_t* gt_offsets, int N,
                             int64_t T, PriorSrc priors_cxcycz, int64_t P, float t0, float t1,
                             int64_t* true_classes, float* true_locs, float* overlap, int32_t* object_for_prior,
                             int32_t* prior_for_object, void* best_key_ws, void* stream) {
  if (!gt_offsets || (!priors_cxcycz.ptr && !priors_cxcycz.tbl) || !true_classes || !true_locs || !overlap || !object_for_prior) return SSD3D_ERR_ARG;
  if (N <= 0 || P <= 0 || T < 0 || P > 0x7fffffffll) return SSD3D_ERR_ARG;
  if (T > 0 && (!gt_boxes || !gt_labels || !prior_for_object || !best_key_ws)) return SSD3D_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)((P + 255) / 256), (unsigned)N);
  if (T > 0) {
    cudaError_t e = cudaMemsetAsync(best_key_ws, 0, (size_t)T * 8, st);
    if (e != cudaSuccess) return (int)e;
    match_iou_kernel<<<grid, 256, 0, st>>>(gt_boxes, gt_offsets, priors_cxcycz, P, overlap, object_for_prior,
                                           static_cast<unsigned long long*>(best_key_ws));
    SSD3D_CHECK_LAUNCH();
  }
  match_assign_kernel<<<grid, 256, 0, st>>>(gt_boxes, reinterpret_cast<const long long*>(gt_labels), gt_offsets,
                                            priors_cxcycz, P, t0, t1,
                                            static_cast<const unsigned long long*>(best_key_ws), overlap,
                                            object_for_prior, prior_for_object,
                                            reinterpret_cast<long long*>(true_classes), true_locs);
  SSD3D_CHECK_LAUNCH();
  return SSD3D_OK;
}

extern "C" int ssd3d_match_priors(const float* gt_boxes, const int64_t* gt_labels, const int32_t* gt_offsets, int N,
                                  int64_t T, const float* priors_cxcycz, int64_t P, float t0, float t1,
                                  int64_t* true_classes, float* true_locs, float* overlap, int32_t* object_for_prior,
                                  int32_t* prior_for_object, void* best_key_ws, void* stream) {
  return match_priors_impl(gt_boxes, gt_labels, gt_offsets, N, T, PriorSrc{priors_cxcycz, nullptr}, P, t0, t1,
                           true_classes, true_locs, overlap, object_for_prior, prior_for_object, best_key_ws, stream);
}
extern "C" int ssd3d_match_priors_analytic(const float* gt_boxes, const int64_t* gt_labels, const int32_t* gt_offsets,
                                           int N, int64_t T, const ssd3d_prior_table* table, int64_t P, float t0,
                                           float t1, int64_t* true_classes, float* true_locs, float* overlap,
                                           int32_t* object_for_prior, int32_t* prior_for_object, void* best_key_ws,
                                           void* stream) {
  return match_priors_impl(gt_boxes, gt_labels, gt_offsets, N, T, PriorSrc{nullptr, table}, P, t0, t1, true_classes,
                           true_locs, overlap, object_for_prior, prior_for_object, best_key_ws, stream);
}

extern "C" int64_t ssd3d_multibox_workspace_bytes(int N, int64_t P) {
  if (N <= 0 || P <= 0) return 0;
  return align256(sizeof(LossAcc)) + align256(4ll * N) + align256(4ll * N * P) + align256(1ll * N * P);
}

extern "C" int ssd3d_multibox_loss(const float* locs, const float* scores, const int64_t* true_classes,
                                   const float* true_locs, int N, int64_t P, int n_classes, float alpha,
                                   int hard_negative_mining, int neg_pos_ratio, float* out_loss, int32_t* n_pos_out,
                                   float* grad_locs, float* grad_scores, void* workspace, int64_t workspace_bytes,
                                   void* stream) {
  if (!locs || !scores || !true_classes || !true_locs || !out_loss || !workspace) return SSD3D_ERR_ARG;
  if (N <= 0 || P <= 0 || n_classes < 2) return SSD3D_ERR_ARG;
  if (workspace_bytes < ssd3d_multibox_workspace_bytes(N, P)) return SSD3D_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  LossAcc* acc = reinterpret_cast<LossAcc*>(ws);
  int* n_pos_img = reinterpret_cast<int*>(ws + align256(sizeof(LossAcc)));
  float* ce_neg = reinterpret_cast<float*>(ws + align256(sizeof(LossAcc)) + align256(4ll * N));
  uint8_t* sel = ws + align256(sizeof(LossAcc)) + align256(4ll * N) + align256(4ll * N * P);
  cudaError_t e = cudaMemsetAsync(ws, 0, (size_t)(align256(sizeof(LossAcc)) + align256(4ll * N)), st);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((unsigned)((P + 255) / 256), (unsigned)N);
  const long long* tc = reinterpret_cast<const long long*>(true_classes);
  multibox_reduce_kernel<<<grid, 256, 0, st>>>(locs, scores, tc, true_locs, P, n_classes, acc, n_pos_img,
                                               hard_negative_mining ? ce_neg : nullptr);
  SSD3D_CHECK_LAUNCH();
  if (hard_negative_mining) {
    hnm_select_kernel<<<N, 1024, 0, st>>>(ce_neg, n_pos_img, P, neg_pos_ratio, sel, acc);
    SSD3D_CHECK_LAUNCH();
  }
  multibox_finalize_kernel<<<1, 1, 0, st>>>(acc, hard_negative_mining, out_loss, n_pos_out);
  SSD3D_CHECK_LAUNCH();
  if (grad_locs || grad_scores) {
    multibox_grad_kernel<<<grid, 256, 0, st>>>(locs, scores, tc, true_locs, P, n_classes, alpha, acc,
                                               hard_negative_mining ? sel : nullptr, grad_locs, grad_scores);
    SSD3D_CHECK_LAUNCH();
  }
  return SSD3D_OK;
}
