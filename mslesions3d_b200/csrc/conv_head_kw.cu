// SSD head for LARGE feature maps as "kw-GEMM + (kd,kh) stencil" (ssd3d.py:131-167).
//
// With N = 16 output channels a UMMA costs about as much as one with N = 144 -- not in the tensor pipe (its time is
// max(M,128)*N/256 cycles: scripts/ubench/umma_rate.cu measures 48 / 64 / 128 cycles at N = 64 / 128 / 256) but in the
// single issuing thread, whose barrier waits, descriptor arithmetic and commits per UMMA exceed the pipe time of a
// narrow instruction (the issue queue is shallow) -- and the per-tap implicit
// GEMM of gemm_tc.cu re-reads every activation 27 times from L2 (221 MB for the 16^3 x 128-channel map of the
// benchmark).  Here the 3x3x3 conv is split:
//   1. head_kw_gemm_kernel: Y[v][(kd,kh), n] = sum_{kw,c} x[v + (0,0,kw-1)][c] * w[n][kd,kh,kw][c]
//      -- an implicit GEMM over the 3 W-taps only (K = 3*C, three W-shifted 5-D TMA boxes per 64-channel chunk,
//      zero fill = padding) with N = 9*16 = 144 columns: 9x more work per UMMA, 3x instead of 27x activation reads;
//   2. head_stencil_kernel: out[d,h,w][n] = bias[n] + sum_{kd,kh} Y[d+kd-1, h+kh-1, w][(kd,kh), n]
//      -- nine shifted 16-byte reads per output quad (Y is L2 resident), fixed summation order, written straight
//      into the concatenated (N,P,6) / (N,P,n_classes) outputs with the NaN flags.
// The weight matrix is re-tiled from the standard packed head weight (NPAD=16, 27*C) by a tiny kernel per call.
#include "common.cuh"
#include "tma_host.h"
#include "../../include/ssd3d_b200.h"

namespace ssd3d {

constexpr int KW_N = 144;      // 9 (kd,kh) groups x 16 output columns

struct HeadKwParams {
  int C, D, H, W, N;
  int TW, TH, TD, TN;
  int tiles_w, tiles_h, tiles_d;
  int num_kb, stages;
  int BN, tmem_cols;           // columns per CTA (144, or 48 with three column tiles on small maps)
  float* Y;                    // (N, 144, D*H*W) fp32: column-major per image, so that the row-per-thread epilogue
                               // of the GEMM and the voxel-per-thread stencil both access it coalesced
  // stencil
  int n_loc, n_cls, bpl, n_classes;
  long long P, prior_off;
  float* locs;
  float* scores;
  const float* bias;
  int* nan_flag;
};

__device__ __forceinline__ uint64_t kw_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
  return d;
}

// w2[(g*16 + n)][kw*C + c] = w[n][((g*3 + kw))*C + c],  g = kd*3 + kh
__global__ void __launch_bounds__(256) head_kw_repack_kernel(const __nv_bfloat16* __restrict__ w,
                                                             __nv_bfloat16* __restrict__ w2, int C) {
  pdl_wait();
  pdl_launch_dependents();
  const int total = KW_N * 3 * C / 8;          // 16-byte chunks
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int per_row = 3 * C / 8;
  const int row = i / per_row, ch = i % per_row;
  const int g = row >> 4, n = row & 15;
  const int col = ch * 8, kw = col / C, c = col - kw * C;
  *reinterpret_cast<uint4*>(w2 + (size_t)row * 3 * C + col) =
      __ldg(reinterpret_cast<const uint4*>(w + (size_t)n * 27 * C + (size_t)(g * 3 + kw) * C + c));
}

__global__ void __launch_bounds__(192, 2) head_kw_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB,
                                                              const HeadKwParams p) {
  extern __shared__ uint8_t kw_raw[];
  const uint32_t raw = smem_u32(kw_raw);
  uint8_t* smem = kw_raw + ((1024u - (raw & 1023u)) & 1023u);
  constexpr int A_BYTES = 128 * 64 * 2;
  const int B_BYTES = p.BN * 64 * 2;
  const int STAGE_BYTES = A_BYTES + B_BYTES;
  const int n0 = blockIdx.y * p.BN;            // first Y column of this CTA
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * STAGE_BYTES);
  uint64_t* empty = full + p.stages;
  uint64_t* tmem_full = empty + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  int t = blockIdx.x;
  const int tw0 = (t % p.tiles_w) * p.TW; t /= p.tiles_w;
  const int th0 = (t % p.tiles_h) * p.TH; t /= p.tiles_h;
  const int td0 = (t % p.tiles_d) * p.TD; t /= p.tiles_d;
  const int tn0 = t * p.TN;
  const int KC = p.C / 64;

  if (warp == 4) {
    if (lane == 0) {
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % p.stages, use = kb / p.stages;
        if (use > 0) mbar_wait(&empty[s], (uint32_t)((use - 1) & 1));
        uint8_t* a_dst = smem + (size_t)s * STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], (uint32_t)STAGE_BYTES);
        const int kw = kb / KC, kc = kb - kw * KC;
        tma_load_5d(a_dst, &tmA, &full[s], kc * 64, tw0 + kw - 1, th0, td0, tn0);
        tma_load_2d(a_dst + A_BYTES, &tmB, &full[s], kw * p.C + kc * 64, n0);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, p.BN);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % p.stages;
        mbar_wait(&full[s], (uint32_t)((kb / p.stages) & 1));
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint64_t da = kw_desc(a_addr), db = kw_desc(a_addr + A_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
    }
    __syncwarp();
  } else {
    mbar_wait(tmem_full, 0);
    __syncwarp();
    tc_fence_after();
    const int row = warp * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    int r = row;
    const int w = tw0 + r % p.TW; r /= p.TW;
    const int h = th0 + r % p.TH; r /= p.TH;
    const int d = td0 + r % p.TD; r /= p.TD;
    const int n = tn0 + r;
    const bool valid = (w < p.W) && (h < p.H) && (d < p.D) && (n < p.N);
    // Y is stored column-major per image: the 32 lanes of a warp are consecutive voxels of the tile (8 along W,
    // then H), so each scalar store instruction writes whole 32-byte sectors (4 per warp) instead of 32 half-used
    // lines with the row-major float4 stores this replaces
    const long long V = (long long)p.D * p.H * p.W;
    float* ycol = p.Y + ((long long)n * KW_N + n0) * V + ((long long)d * p.H + h) * p.W + w;
    for (int c = 0; c < p.BN; c += 16) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(taddr + (uint32_t)c, v);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int q = 0; q < 16; ++q) ycol[(long long)(c + q) * V] = __uint_as_float(v[q]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// out[d,h,w][c] = bias[c] + sum_{g = kd*3+kh ascending, in bounds} Y[g*16 + c][d+kd-1, h+kh-1, w]
// thread = one voxel x 4 of the 16 columns (warp q of the block owns columns 4q .. 4q+3 of the block's 32 voxels):
// every load is coalesced across the warp (consecutive voxels of a column), and the map's N*D*H*W voxels give four
// times the threads of a voxel-per-thread mapping -- at 32 768 voxels that one left the machine at 11 % occupancy
// with 144 dependent L2 loads per thread (14.7 us per launch in profiles/r02_launches_final.csv).
__global__ void __launch_bounds__(128) head_stencil_kernel(const HeadKwParams p) {
  pdl_wait();
  pdl_launch_dependents();
  const long long V = (long long)p.D * p.H * p.W;
  const long long total = (long long)p.N * V;
  const int cq = threadIdx.x >> 5;                                  // column quad
  const long long gid = (long long)blockIdx.x * 32 + (threadIdx.x & 31);
  if (gid >= total) return;
  const int n = (int)(gid / V);
  long long r = gid - (long long)n * V;
  const int w = (int)(r % p.W); r /= p.W;
  const int h = (int)(r % p.H);
  const int d = (int)(r / p.H);
  const int ncol = p.n_loc + p.n_cls;
  const int c0 = 4 * cq;
  if (c0 >= ncol) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int g = 0; g < 9; ++g) {
    const int dd = d + g / 3 - 1, hh = h + g % 3 - 1;
    if ((unsigned)dd >= (unsigned)p.D || (unsigned)hh >= (unsigned)p.H) continue;
    const float* src = p.Y + ((long long)n * KW_N + g * 16 + c0) * V + ((long long)dd * p.H + hh) * p.W + w;
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = __ldg(src + (long long)c * V);
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = __fadd_rn(acc[c], v[c]);
  }
  const long long prior = p.prior_off + (((long long)d * p.H + h) * p.W + w) * p.bpl;
  float* lp = p.locs + ((long long)n * p.P + prior) * 6;
  float* sp = p.scores + ((long long)n * p.P + prior) * p.n_classes;
  bool bad_l = false, bad_s = false;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int c = c0 + q;
    if (c >= ncol) break;
    const float val = __fadd_rn(acc[q], __ldg(p.bias + c));
    if (c < p.n_loc) { lp[c] = val; bad_l |= (val != val); }
    else { sp[c - p.n_loc] = val; bad_s |= (val != val); }
  }
  if (p.nan_flag) {
    if (bad_l) atomicOr(p.nan_flag, SSD3D_NAN_LOCS);
    if (bad_s) atomicOr(p.nan_flag, SSD3D_NAN_SCORES);
  }
}

static inline int kw_p2ceil(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

}  // namespace ssd3d

using namespace ssd3d;

int64_t ssd3d_head_kw_workspace_bytes(int N, int C, int D, int H, int W) {
  const long long M = (long long)N * D * H * W;
  return (int64_t)(M * KW_N * 4 + (long long)KW_N * 3 * C * 2 + 1024);
}

bool ssd3d_head_kw_applicable(int N, int C, int D, int H, int W, int NPAD) {
  return NPAD == 16 && C % 64 == 0 && (long long)N * D * H * W >= 256;
}

// w_is_kw != 0: `w` already is the (144, 3*C) tiling (ssd3d_head_weight_kw); otherwise it is re-tiled here per call
int ssd3d_head_conv_kw(const void* x, const void* w, int w_is_kw, const float* bias, float* locs, float* scores, int N,
                       int C, int D, int H, int W, int bpl, int n_classes, int NPAD, int64_t P, int64_t prior_offset,
                       int* nan_flag, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  if (!ssd3d_head_kw_applicable(N, C, D, H, W, NPAD)) return SSD3D_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < ssd3d_head_kw_workspace_bytes(N, C, D, H, W)) return SSD3D_ERR_ARG;
  const long long M = (long long)N * D * H * W;
  float* Y = static_cast<float*>(workspace);
  const __nv_bfloat16* w2 = static_cast<const __nv_bfloat16*>(w);
  if (!w_is_kw) {
    __nv_bfloat16* tmp = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(workspace) +
                                                          ((M * KW_N * 4 + 1023) & ~1023ll));
    const int total = KW_N * 3 * C / 8;
    SSD3D_LAUNCH_PDL(head_kw_repack_kernel, dim3((total + 255) / 256), dim3(256), 0, st,
                     static_cast<const __nv_bfloat16*>(w), tmp, C);
    w2 = tmp;
  }
  HeadKwParams p{};
  p.C = C; p.D = D; p.H = H; p.W = W; p.N = N;
  p.TW = kw_p2ceil(W) < 8 ? kw_p2ceil(W) : 8;
  p.TH = kw_p2ceil(H) < 4 ? kw_p2ceil(H) : 4;
  {
    const int rest = 128 / (p.TW * p.TH);
    p.TD = kw_p2ceil(D) < rest ? kw_p2ceil(D) : rest;
  }
  p.TN = 128 / (p.TW * p.TH * p.TD);
  p.tiles_w = (W + p.TW - 1) / p.TW;
  p.tiles_h = (H + p.TH - 1) / p.TH;
  p.tiles_d = (D + p.TD - 1) / p.TD;
  const int tiles_n = (N + p.TN - 1) / p.TN;
  p.num_kb = 3 * (C / 64);
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_d * tiles_n;
  p.BN = (m_tiles >= 148) ? KW_N : 48;        // small maps: three column tiles so that more SMs take part
  p.tmem_cols = (p.BN == KW_N) ? 256 : 64;
  // 3 x 34 KB stages: two CTAs per SM on large maps; deeper ring for the long K loops of the small ones
  p.stages = (p.BN == KW_N) ? 3 : 6;
  if (p.stages > p.num_kb) p.stages = p.num_kb;
  p.Y = Y;
  p.n_loc = bpl * 6; p.n_cls = bpl * n_classes; p.bpl = bpl; p.n_classes = n_classes;
  p.P = P; p.prior_off = prior_offset; p.locs = locs; p.scores = scores; p.bias = bias; p.nan_flag = nan_flag;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
    const uint64_t strides[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2,
                                 (uint64_t)D * H * W * C * 2};
    const uint32_t box[5] = {64u, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TD, (uint32_t)p.TN};
    if (make_tma_bf16(&tmA, x, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return SSD3D_ERR_TMA;
  }
  {
    const uint64_t dims[2] = {(uint64_t)3 * C, (uint64_t)KW_N};
    const uint64_t strides[1] = {(uint64_t)3 * C * 2};
    const uint32_t box[2] = {64u, (uint32_t)p.BN};
    if (make_tma_bf16(&tmB, w2, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return SSD3D_ERR_TMA;
  }
  const size_t smem = (size_t)p.stages * (128 * 64 * 2 + p.BN * 64 * 2) + (2 * p.stages + 1) * 8 + 16 + 1024;
  cudaError_t e = cudaFuncSetAttribute(head_kw_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((unsigned)m_tiles, (unsigned)(KW_N / p.BN));
  SSD3D_LAUNCH_PDL(head_kw_gemm_kernel, grid, dim3(192), smem, st, tmA, tmB, p);
  SSD3D_LAUNCH_PDL(head_stencil_kernel, dim3((unsigned)((M + 31) / 32)), dim3(128), 0, st, p);
  return SSD3D_OK;
}

// (NPAD = 16, 27*C) packed head weight -> (144, 3*C) tiling of the kw-GEMM, once per weight version
extern "C" int ssd3d_head_weight_kw(const void* w, int C, void* w_kw, void* stream) {
  if (!w || !w_kw || C <= 0 || (C % 64)) return SSD3D_ERR_ARG;
  const int total = KW_N * 3 * C / 8;
  SSD3D_LAUNCH_PDL(head_kw_repack_kernel, dim3((total + 255) / 256), dim3(256), 0, static_cast<cudaStream_t>(stream),
                   static_cast<const __nv_bfloat16*>(w), static_cast<__nv_bfloat16*>(w_kw), C);
  return SSD3D_OK;
}

extern "C" int ssd3d_head_kw_supported(int N, int C, int D, int H, int W, int NPAD) {
  return ssd3d_head_kw_applicable(N, C, D, H, W, NPAD) ? 1 : 0;
}
