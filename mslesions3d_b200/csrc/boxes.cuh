// Box arithmetic shared by the detect / match / box-utility kernels.  Every step is one separately
// rounded fp32 operation in the reference's order (utils.py:42-149): the intrinsics below forbid the
// compiler from contracting a*b+c into an FMA, which would change low-order bits and with them the
// integer outputs (NMS keep lists, matched prior indices) that must be exact.
#pragma once
#include "common.cuh"

namespace ssd3d {

struct Box6 {
  float v[6];
};

__device__ __forceinline__ Box6 load_box(const float* p) {
  Box6 b;
  const float2 a = *reinterpret_cast<const float2*>(p);
  const float2 c = *reinterpret_cast<const float2*>(p + 2);
  const float2 d = *reinterpret_cast<const float2*>(p + 4);
  b.v[0] = a.x; b.v[1] = a.y; b.v[2] = c.x; b.v[3] = c.y; b.v[4] = d.x; b.v[5] = d.y;
  return b;
}
__device__ __forceinline__ void store_box(float* p, const Box6& b) {
  *reinterpret_cast<float2*>(p) = make_float2(b.v[0], b.v[1]);
  *reinterpret_cast<float2*>(p + 2) = make_float2(b.v[2], b.v[3]);
  *reinterpret_cast<float2*>(p + 4) = make_float2(b.v[4], b.v[5]);
}

// torch.clamp(x, min=0): NaN propagates
__device__ __forceinline__ float clamp_min0(float x) { return (x < 0.0f) ? 0.0f : x; }

// utils.py:142-147
__device__ __forceinline__ float box_volume(const Box6& a) {
  return __fmul_rn(__fmul_rn(__fsub_rn(a.v[3], a.v[0]), __fsub_rn(a.v[4], a.v[1])), __fsub_rn(a.v[5], a.v[2]));
}
// utils.py:119-122
__device__ __forceinline__ float box_intersection(const Box6& a, const Box6& b) {
  const float d0 = clamp_min0(__fsub_rn(fminf(a.v[3], b.v[3]), fmaxf(a.v[0], b.v[0])));
  const float d1 = clamp_min0(__fsub_rn(fminf(a.v[4], b.v[4]), fmaxf(a.v[1], b.v[1])));
  const float d2 = clamp_min0(__fsub_rn(fminf(a.v[5], b.v[5]), fmaxf(a.v[2], b.v[2])));
  return __fmul_rn(__fmul_rn(d0, d1), d2);
}
// utils.py:135-149: inter / ((vol_a + vol_b) - inter)
__device__ __forceinline__ float box_iou(const Box6& a, float vol_a, const Box6& b, float vol_b) {
  const float inter = box_intersection(a, b);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(vol_a, vol_b), inter));
}

// utils.py:50-51
__device__ __forceinline__ Box6 cxcycz_to_xyz(const Box6& c) {
  Box6 r;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float half = __fdiv_rn(c.v[3 + k], 2.0f);
    r.v[k] = __fsub_rn(c.v[k], half);
    r.v[3 + k] = __fadd_rn(c.v[k], half);
  }
  return r;
}
// utils.py:101-102
__device__ __forceinline__ Box6 xyz_to_cxcycz(const Box6& b) {
  Box6 r;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    r.v[k] = __fdiv_rn(__fadd_rn(b.v[3 + k], b.v[k]), 2.0f);
    r.v[3 + k] = __fsub_rn(b.v[3 + k], b.v[k]);
  }
  return r;
}
// utils.py:67-68
__device__ __forceinline__ Box6 gcxgcygcz_to_cxcycz(const Box6& g, const Box6& p) {
  Box6 r;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    r.v[k] = __fadd_rn(__fdiv_rn(__fmul_rn(g.v[k], p.v[3 + k]), 10.0f), p.v[k]);
    r.v[3 + k] = __fmul_rn(expf(__fdiv_rn(g.v[3 + k], 5.0f)), p.v[3 + k]);
  }
  return r;
}
// utils.py:88-89
__device__ __forceinline__ Box6 cxcycz_to_gcxgcygcz(const Box6& c, const Box6& p) {
  Box6 r;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    r.v[k] = __fdiv_rn(__fsub_rn(c.v[k], p.v[k]), __fdiv_rn(p.v[3 + k], 10.0f));
    r.v[3 + k] = __fmul_rn(logf(__fdiv_rn(c.v[3 + k], p.v[3 + k])), 5.0f);
  }
  return r;
}

// ------------------------------------------------------------------------------------------------
// Prior source: the (P, 6) tensor, or the closed form of ssd3d.py:300-330 evaluated from the prior index
// (float64 arithmetic rounded once to fp32, like the reference's Python doubles -> FloatTensor).
// ------------------------------------------------------------------------------------------------
static_assert(sizeof(ssd3d_prior_table) == 336, "ssd3d_prior_table layout is part of the ABI (_lib.PriorTable mirrors it)");

struct PriorSrc {
  const float* ptr;                  // (P, 6) centre-size priors, or nullptr
  const ssd3d_prior_table* tbl;      // device copy of the table (used when ptr == nullptr)
};

__device__ __forceinline__ Box6 prior_from_table(const ssd3d_prior_table* __restrict__ t, long long p) {
  int l = 0;
  const int nl = t->n_layers;
  while (l + 1 < nl && p >= t->start[l + 1]) ++l;
  const long long local = p - t->start[l];
  const int nb = t->n_boxes[l];
  const int b = (int)(local % nb);
  long long v = local / nb;
  const int d0 = t->d0[l], d1 = t->d1[l], d2 = t->d2[l];
  const int k = (int)(v % d2); v /= d2;
  const int j = (int)(v % d1);
  const int i = (int)(v / d1);
  Box6 r;
  r.v[0] = __double2float_rn(__ddiv_rn((double)j + 0.5, (double)d1));     // cx <- array axis 1 (ssd3d.py:305)
  r.v[1] = __double2float_rn(__ddiv_rn((double)i + 0.5, (double)d0));     // cy <- array axis 0 (ssd3d.py:306)
  r.v[2] = __double2float_rn(__ddiv_rn((double)k + 0.5, (double)d2));     // cz <- array axis 2 (ssd3d.py:304)
  const float s = t->size[l][b];
  r.v[3] = s; r.v[4] = s; r.v[5] = s;
  return r;
}

__device__ __forceinline__ Box6 load_prior(const PriorSrc& src, long long p) {
  return src.ptr ? load_box(src.ptr + p * 6) : prior_from_table(src.tbl, p);
}

// order-preserving map float -> uint32 (ascending)
__device__ __forceinline__ uint32_t float_orderable(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_from_orderable(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

}  // namespace ssd3d
