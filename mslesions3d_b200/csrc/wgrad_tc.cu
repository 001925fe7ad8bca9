// Backward-pass contractions of the training step on tcgen05 (three kernels: pointwise weight gradient, SSD head
// weight gradient, SSD head data gradient), operands straight from TMA in the layouts the channels-last tensors
// already have.  First: the pointwise-conv weight gradient (autograd of nn.Conv3d(kernel_size=1), mobilenet.py:40
// under LSSD3D.training_step, ssd3d.py:467-531):
//
//      dW[n][k] = sum_m dz[m][n] * x[m][k]          dz (M, Cout) bf16, x (M, Cin) bf16, dW (Cout, Cin) fp32
//
// Both operands are channels-last, so the reduction index m (the voxel) is the SLOW index of both: in UMMA terms
// A = dz^T and B = x^T are "MN-major" operands.  That is exactly what a plain TMA box of [64 voxels][64 channels]
// with the 128-byte swizzle puts into shared memory -- rows of 128 B = 64 channels (the M / N index, contiguous),
// one row per voxel (the K index), 16-byte chunks XOR-ed with the row index modulo 8 -- i.e. the canonical MN-major
// SWIZZLE_128B layout  ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO))  in elements (cute::UMMA::make_umma_desc<Major::MN>):
// SBO = 1024 B between 8-voxel groups, LBO = the box size between 64-channel blocks.  No transposition pass, no
// ldmatrix.trans: the instruction descriptor's a_major / b_major bits tell the tensor core to read them that way.
// Cin = 32 (the first block) has 64-byte rows: same scheme with the 64-byte swizzle.
//
// Work item = (128 output channels, <= 128 input channels, a contiguous range of 64-voxel chunks); CTA = 192 threads:
// warp 0 TMA producer, warp 1 TMEM owner + single-lane UMMA issue (4 x K16 per chunk), warps 2-5 epilogue
// (TMEM lane quarter = warp % 4) writing the fp32 tile of this voxel range to its slab; the slabs are added in a
// fixed order by sum_partials (train.cu), or the tile goes straight to dW when there is a single range.
// Out-of-range voxels and output channels >= Cout are zero-filled by the TMA unit.
#include "common.cuh"
#include "tma_host.h"

namespace ssd3d {

typedef __nv_bfloat16 bf16;

namespace {

constexpr int WT_KT = 64;            // voxels per stage
constexpr int WT_STAGES = 4;
constexpr int WT_THREADS = 192;
constexpr int WT_A_BYTES = 2 * WT_KT * 128;        // two 64-channel boxes of dz
constexpr int WT_B_BYTES = 2 * WT_KT * 128;        // up to two 64-channel boxes of x
constexpr int WT_STAGE_BYTES = WT_A_BYTES + WT_B_BYTES;
constexpr int WT_SMEM = WT_STAGES * WT_STAGE_BYTES + 1024 /*alignment*/ + 256 /*barriers*/;

struct WgradTcParams {
  long long M;
  int Cin, Cout;
  int NT;                  // input channels per tile: 32, 64 or 128
  int chunks_per_split;
  int n_chunks;            // ceil(M / 64)
  int ld;                  // row pitch of the output (floats)
  long long slab;          // floats between two voxel ranges' outputs
  float* out;
};

// MN-major operand: `row_bytes` = 128 (SWIZZLE_128B, 64 channels per row) or 64 (SWIZZLE_64B, 32 channels)
__device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr, uint32_t row_bytes, uint32_t block_stride) {
  const uint32_t lbo = block_stride, sbo = 8 * row_bytes;
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(row_bytes == 128 ? 2u : 4u) << 61;
  return d;
}

__global__ void __launch_bounds__(WT_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmZ,
                                                                 const __grid_constant__ CUtensorMap tmX,
                                                                 const WgradTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)WT_STAGES * WT_STAGE_BYTES);
  uint64_t* empty = full + WT_STAGES;
  uint64_t* acc_full = empty + WT_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NT = p.NT;
  const uint32_t tmem_cols = NT < 32 ? 32u : (uint32_t)NT;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmZ);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < WT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  const int k0 = blockIdx.x * NT;            // first input channel of the tile
  const int n0 = blockIdx.y * 128;           // first output channel
  const int c_begin = blockIdx.z * p.chunks_per_split;
  int c_end = c_begin + p.chunks_per_split;
  if (c_end > p.n_chunks) c_end = p.n_chunks;
  const int n_iter = c_end - c_begin;
  const uint32_t x_row = NT >= 64 ? 128u : 64u;                    // bytes per voxel row of one x box
  const int x_boxes = NT >= 64 ? NT / 64 : 1;
  const uint32_t x_box_bytes = WT_KT * x_row;
  const uint32_t stage_tx = WT_A_BYTES + (uint32_t)x_boxes * x_box_bytes;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % WT_STAGES;
        if (it >= WT_STAGES) mbar_wait(&empty[s], ((it / WT_STAGES) - 1) & 1);
        uint8_t* a = smem + (size_t)s * WT_STAGE_BYTES;
        uint8_t* b = a + WT_A_BYTES;
        const int m0 = (c_begin + it) * WT_KT;
        mbar_arrive_expect_tx(&full[s], stage_tx);
        tma_load_2d(a, &tmZ, &full[s], n0, m0);
        tma_load_2d(a + WT_KT * 128, &tmZ, &full[s], n0 + 64, m0);      // all zero when n0 + 64 >= Cout
        for (int j = 0; j < x_boxes; ++j) tma_load_2d(b + (size_t)j * x_box_bytes, &tmX, &full[s], k0 + 64 * j, m0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // a_major = b_major = MN (bits 15, 16 of the instruction descriptor)
      const uint32_t idesc = umma_idesc_bf16(128, NT) | (1u << 15) | (1u << 16);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % WT_STAGES;
        mbar_wait(&full[s], (it / WT_STAGES) & 1);
        tc_fence_after();
        const uint32_t a = smem_u32(smem + (size_t)s * WT_STAGE_BYTES);
        const uint32_t b = a + WT_A_BYTES;
#pragma unroll
        for (int ks = 0; ks < WT_KT / 16; ++ks) {
          const uint64_t da = desc_mn(a + ks * 16 * 128, 128, WT_KT * 128);
          const uint64_t db = desc_mn(b + ks * 16 * x_row, x_row, x_box_bytes);
          umma_bf16_ss(tmem_base, da, db, idesc, (it > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);               // the stage is free once these UMMAs have read it
      }
      umma_commit(acc_full);
    }
  } else {
    // ===================== epilogue: TMEM -> fp32 tile of this voxel range =====================
    const int q = warp & 3;                   // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;            // output channel within the tile
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const bool valid = (n0 + row) < p.Cout && n_iter > 0;
    float* dst = p.out + (size_t)blockIdx.z * p.slab + (size_t)(n0 + row) * p.ld + k0;
    for (int c = 0; c < NT; c += 16) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(dst + c + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}


// ------------------------------------------------------------------------------------------------
// SSD head (3x3x3, pad 1) weight gradient on tcgen05, one 16-column group of the packed [loc | class] gradient rows
// (autograd of the loc_convs / cl_convs of PredictionConvolutions, ssd3d.py:131-167):
//
//      dW[c][tap][n] = sum_u x[u][c] * dO[u - off(tap)][n]           u = voxel, off(tap) = (kd-1, kh-1, kw-1)
//
// i.e. for each kd ONE GEMM  D[c][j*16 + n]  (M = 128 channels, N = 9 (kh,kw) taps x 16 columns = 144, K = voxels)
// whose B operand is nine shifted views of the (voxel, 16) gradient rows.  K runs over 64-voxel BOXES
// (bw x bh x bd x bn of the map, chosen per map): a 5-D TMA box of x gives the A rows, the same box shifted by the tap
// gives that tap's 64 x 16 B block, and voxels outside the map -- the conv's zero padding, and the ragged edge of the
// tiling -- are zero-filled by the TMA unit in both.  Both operands are MN-major: A as in the pointwise kernel
// (128-byte rows, SWIZZLE_128B), B with 32-byte rows (16 columns) under SWIZZLE_32B, canonical layout
// ((8,2,m),(8,k)) : ((1,8,LBO),(16,SBO)) -- SBO = 256 B between 8-voxel groups, LBO = 2 KB between taps.
// CTA = (128-channel tile, kd, range of boxes); same warp roles as above; the fp32 tile goes to the slab
// [range][kd][C][144], and head_wsum_kernel adds the ranges in order and scatters to dw_loc / dw_cls (Cout, C, 27).
// ------------------------------------------------------------------------------------------------
constexpr int HT_STAGES = 4;
constexpr int HT_A_BYTES = 2 * 64 * 128;                 // 128 channels x 64 voxels
constexpr int HT_B_BYTES = 9 * 64 * 32;                  // 9 taps x 64 voxels x 16 columns
constexpr int HT_STAGE_BYTES = 35 * 1024;                // A + B = 34 KB, stages 1 KB aligned
constexpr int HT_SMEM = HT_STAGES * HT_STAGE_BYTES + 1024 + 256;

struct HeadWgradTcParams {
  int C;
  int bw, bh, bd, bn;          // box extent (product 64)
  int tw, th, td;              // boxes along W, H, D (tn follows from the total)
  int n_boxes, boxes_per_split;
  float* slab;                 // [split][3][C][144]
};

__device__ __forceinline__ uint64_t desc_mn_sw32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(2048u >> 4) << 16;      // LBO: next 16-column block = next tap
  d |= (uint64_t)(256u >> 4) << 32;       // SBO: next group of 8 voxels
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;                 // SWIZZLE_32B
  return d;
}

__global__ void __launch_bounds__(WT_THREADS, 1) head_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                      const __grid_constant__ CUtensorMap tmG,
                                                                      const HeadWgradTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)HT_STAGES * HT_STAGE_BYTES);
  uint64_t* empty = full + HT_STAGES;
  uint64_t* acc_full = empty + HT_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 256;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmG);
    for (int s = 0; s < HT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  const int c0 = blockIdx.x * 128;
  const int kd = blockIdx.y;
  const int b_begin = blockIdx.z * p.boxes_per_split;
  int b_end = b_begin + p.boxes_per_split;
  if (b_end > p.n_boxes) b_end = p.n_boxes;
  const int n_iter = b_end - b_begin;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % HT_STAGES;
        if (it >= HT_STAGES) mbar_wait(&empty[s], ((it / HT_STAGES) - 1) & 1);
        uint8_t* a = smem + (size_t)s * HT_STAGE_BYTES;
        uint8_t* b = a + HT_A_BYTES;
        int t = b_begin + it;
        const int w0 = (t % p.tw) * p.bw; t /= p.tw;
        const int h0 = (t % p.th) * p.bh; t /= p.th;
        const int d0 = (t % p.td) * p.bd; t /= p.td;
        const int n0 = t * p.bn;
        mbar_arrive_expect_tx(&full[s], HT_A_BYTES + HT_B_BYTES);
        tma_load_5d(a, &tmX, &full[s], c0, w0, h0, d0, n0);
        tma_load_5d(a + 64 * 128, &tmX, &full[s], c0 + 64, w0, h0, d0, n0);       // zeros when c0 + 64 >= C
#pragma unroll
        for (int j = 0; j < 9; ++j)     // dO[u - off]: the box moves by -(k - 1) along each axis
          tma_load_5d(b + j * 2048, &tmG, &full[s], 0, w0 - (j % 3 - 1), h0 - (j / 3 - 1), d0 - (kd - 1), n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, 144) | (1u << 15) | (1u << 16);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % HT_STAGES;
        mbar_wait(&full[s], (it / HT_STAGES) & 1);
        tc_fence_after();
        const uint32_t a = smem_u32(smem + (size_t)s * HT_STAGE_BYTES);
        const uint32_t b = a + HT_A_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_bf16_ss(tmem_base, desc_mn(a + ks * 16 * 128, 128, 64 * 128), desc_mn_sw32(b + ks * 16 * 32), idesc,
                       (it > 0 || ks > 0) ? 1u : 0u);
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const bool valid = (c0 + row) < p.C && n_iter > 0;
    float* dst = p.slab + (((size_t)blockIdx.z * 3 + kd) * p.C + (c0 + row)) * 144;
#pragma unroll 1
    for (int c = 0; c < 144; c += 16) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(dst + c + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// dw_{loc,cls}[row][c][tap] = sum over ranges of slab[s][kd][c][j*16 + n];  row = 16*g + n of the packed
// [loc rows | class rows | padding] gradient columns.  Thread = (tap, c, n), n fastest: 64-byte reads.
__global__ void __launch_bounds__(256) head_wsum_kernel(const float* __restrict__ slab, int S, int C, int g, int n_loc,
                                                        int n_cls, float* __restrict__ dw_loc,
                                                        float* __restrict__ dw_cls) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 27ll * C * 16) return;
  const int n = (int)(i & 15);
  const int c = (int)((i >> 4) % C);
  const int tap = (int)((i >> 4) / C);
  const int kd = tap / 9, j = tap - kd * 9;
  const float* src = slab + ((size_t)kd * C + c) * 144 + j * 16 + n;
  const size_t stride = (size_t)3 * C * 144;
  float acc = 0.f;
  for (int s = 0; s < S; ++s) acc += src[(size_t)s * stride];
  const int r = g * 16 + n;
  if (r < n_loc) dw_loc[((size_t)r * C + c) * 27 + tap] = acc;
  else if (r - n_loc < n_cls) dw_cls[((size_t)(r - n_loc) * C + c) * 27 + tap] = acc;
}


// ------------------------------------------------------------------------------------------------
// SSD head data gradient on tcgen05 (transposed 3x3x3 conv, 16 gradient columns per group -> C channels; autograd
// of ssd3d.py:131-167 w.r.t. the feature map):
//
//      dx[u][c] = sum_g sum_tap sum_n dO_g[u - off(tap)][n] * w[16g + n][tap*C + c]      (+ addend[u][c])
//
// One UMMA per (group, tap): M = 128 voxels (a 5-D TMA box bw x bh x bd x bn of the map, shifted by the tap, zero
// filled outside the map = the conv's padding), K = the 16 gradient columns -- a 32-byte row, K-major under
// SWIZZLE_32B -- and N = up to 128 channels of the weight rows w[16g .. 16g+15][tap*C + c0 ...], which are MN-major
// (channel contiguous) exactly as TMA delivers them (SWIZZLE_128B, 16 rows = 2 K groups).  27 x groups steps through a
// ring of 8 KB stages, fp32 accumulator in TMEM, epilogue = + addend, bf16, 32-byte stores to the voxel's row.
// ------------------------------------------------------------------------------------------------
constexpr int HD_STAGES = 8;
constexpr int HD_A_BYTES = 128 * 32;                  // 128 voxels x 16 columns
constexpr int HD_B_BYTES = 2 * 16 * 128;              // 16 weight rows x (up to) 2 x 64 channels
constexpr int HD_STAGE_BYTES = HD_A_BYTES + HD_B_BYTES;
constexpr int HD_SMEM = HD_STAGES * HD_STAGE_BYTES + 1024 + 256;

struct HeadDgradTcParams {
  int N, D, H, W, C;
  int NT;                      // channels per tile: 64 or 128
  int groups;
  int lw, lh, ld;              // log2 of the box extents bw, bh, bd (bn = 128 >> (lw + lh + ld))
  int tw, th, td;              // boxes along W, H, D
  const bf16* addend;          // (M, C) or null
  bf16* dx;                    // (M, C)
};

__device__ __forceinline__ uint64_t desc_k_sw32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                 // LBO: unused (K = one 32-byte row)
  d |= (uint64_t)(256u >> 4) << 32;       // SBO: next group of 8 voxels
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;                 // SWIZZLE_32B
  return d;
}

__global__ void __launch_bounds__(WT_THREADS, 1) head_dgrad_tc_kernel(const __grid_constant__ CUtensorMap tmG,
                                                                      const __grid_constant__ CUtensorMap tmW,
                                                                      const HeadDgradTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)HD_STAGES * HD_STAGE_BYTES);
  uint64_t* empty = full + HD_STAGES;
  uint64_t* acc_full = empty + HD_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NT = p.NT;
  const uint32_t tmem_cols = (uint32_t)NT;              // 64 or 128
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmG);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < HD_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  int t = blockIdx.x;
  const int w0 = (t % p.tw) << p.lw; t /= p.tw;
  const int h0 = (t % p.th) << p.lh; t /= p.th;
  const int d0 = (t % p.td) << p.ld; t /= p.td;
  const int lb = 7 - p.lw - p.lh - p.ld;               // log2(bn)
  const int n0 = t << lb;
  const int c0 = blockIdx.y * NT;
  const int n_iter = 27 * p.groups;
  const int w_boxes = NT / 64;
  const uint32_t stage_tx = HD_A_BYTES + (uint32_t)w_boxes * 2048u;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % HD_STAGES;
        if (it >= HD_STAGES) mbar_wait(&empty[s], ((it / HD_STAGES) - 1) & 1);
        uint8_t* a = smem + (size_t)s * HD_STAGE_BYTES;
        uint8_t* b = a + HD_A_BYTES;
        const int g = it / 27, tap = it - g * 27;
        const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
        mbar_arrive_expect_tx(&full[s], stage_tx);
        tma_load_5d(a, &tmG, &full[s], 0, w0 - (kw - 1), h0 - (kh - 1), d0 - (kd - 1), g * p.N + n0);
        for (int j = 0; j < w_boxes; ++j) tma_load_2d(b + j * 2048, &tmW, &full[s], tap * p.C + c0 + 64 * j, g * 16);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, NT) | (1u << 16);      // A K-major, B MN-major
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % HD_STAGES;
        mbar_wait(&full[s], (it / HD_STAGES) & 1);
        tc_fence_after();
        const uint32_t a = smem_u32(smem + (size_t)s * HD_STAGE_BYTES);
        umma_bf16_ss(tmem_base, desc_k_sw32(a), desc_mn(a + HD_A_BYTES, 128, 2048), idesc, it > 0 ? 1u : 0u);
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;                     // voxel of the box, W fastest
    const int iw = row & ((1 << p.lw) - 1);
    const int ih = (row >> p.lw) & ((1 << p.lh) - 1);
    const int id = (row >> (p.lw + p.lh)) & ((1 << p.ld) - 1);
    const int in = row >> (p.lw + p.lh + p.ld);
    const int w = w0 + iw, h = h0 + ih, d = d0 + id, n = n0 + in;
    const bool valid = w < p.W && h < p.H && d < p.D && n < p.N;
    const long long m = (((long long)n * p.D + d) * p.H + h) * p.W + w;
    mbar_wait(acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < NT; c += 16) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      tmem_ld_wait();
      if (valid) {
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
        bf16* dst = p.dx + m * p.C + c0 + c;
        if (p.addend) {
          const uint4 a0 = *reinterpret_cast<const uint4*>(p.addend + m * p.C + c0 + c);
          const uint4 a1 = *reinterpret_cast<const uint4*>(p.addend + m * p.C + c0 + c + 8);
          const uint32_t au[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) { f[2 * j] += bf16_lo(au[j]); f[2 * j + 1] += bf16_hi(au[j]); }
        }
        uint4 o0, o1;
        o0.x = pack_bf16x2(f[0], f[1]);   o0.y = pack_bf16x2(f[2], f[3]);
        o0.z = pack_bf16x2(f[4], f[5]);   o0.w = pack_bf16x2(f[6], f[7]);
        o1.x = pack_bf16x2(f[8], f[9]);   o1.y = pack_bf16x2(f[10], f[11]);
        o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
        *reinterpret_cast<uint4*>(dst) = o0;
        *reinterpret_cast<uint4*>(dst + 8) = o1;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace

// plan shared with the workspace query: tiles, voxel ranges
static void wgrad_tc_plan(long long M, int Cin, int Cout, int* NT, int* tiles, int* splits, int* cps) {
  *NT = Cin >= 128 ? 128 : Cin;                       // 32, 64, 128
  const int kt = Cin / *NT, nt = (Cout + 127) / 128;
  *tiles = kt * nt;
  const int chunks = (int)((M + WT_KT - 1) / WT_KT);
  int s = persistent_sms() / *tiles;
  if (s < 1) s = 1;
  const int by_work = (chunks + 3) / 4;               // at least four chunks per range: the pipeline has to fill
  if (s > by_work) s = by_work;
  if (s < 1) s = 1;
  *cps = (chunks + s - 1) / s;
  *splits = (chunks + *cps - 1) / *cps;
}

int64_t wgrad_tc_workspace_bytes(long long M, int Cin, int Cout) {
  int NT, tiles, S, cps;
  wgrad_tc_plan(M, Cin, Cout, &NT, &tiles, &S, &cps);
  return S > 1 ? (int64_t)S * Cout * Cin * 4 : 0;
}

// -> number of voxel ranges S written ([S][Cout][Cin] in `partial`; S == 1: written to dw directly), < 0: error
int wgrad_tc_launch(const void* dz, const void* x, long long M, int Cin, int Cout, float* dw, float* partial,
                    cudaStream_t st) {
  if (M <= 0 || M >= (1ll << 31) - 64 || (Cin != 32 && Cin % 64) || Cin <= 0 || Cout <= 0 || (Cout % 64)) return -1;
  if (Cin > 128 && Cin % 128) return -1;
  WgradTcParams p{};
  int tiles, S, cps;
  wgrad_tc_plan(M, Cin, Cout, &p.NT, &tiles, &S, &cps);
  p.M = M; p.Cin = Cin; p.Cout = Cout;
  p.chunks_per_split = cps;
  p.n_chunks = (int)((M + WT_KT - 1) / WT_KT);
  p.ld = Cin;
  p.slab = (long long)Cout * Cin;
  p.out = S > 1 ? partial : dw;
  CUtensorMap tmZ, tmX;
  {
    const uint64_t dims[2] = {(uint64_t)Cout, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)Cout * 2};
    const uint32_t box[2] = {64u, (uint32_t)WT_KT};
    if (make_tma_bf16(&tmZ, dz, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return -2;
  }
  {
    const uint64_t dims[2] = {(uint64_t)Cin, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)Cin * 2};
    const uint32_t box[2] = {Cin >= 64 ? 64u : 32u, (uint32_t)WT_KT};
    if (make_tma_bf16(&tmX, x, 2, dims, strides, box, Cin >= 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B))
      return -2;
  }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM) != cudaSuccess)
      return -3;
    attr_set = true;
  }
  const dim3 grid((unsigned)(Cin / p.NT), (unsigned)((Cout + 127) / 128), (unsigned)S);
  if (launch_pdl(wgrad_tc_kernel, grid, dim3(WT_THREADS), (size_t)WT_SMEM, st, tmZ, tmX, p) != cudaSuccess) return -4;
  return S;
}


// 64-voxel box of the (W, H, D, N) map: powers of two with product 64, none larger than its axis; fewest boxes wins.
static bool head_tc_box(int N, int D, int H, int W, int* box, int* cnt) {
  long long best = -1;
  for (int bw = 1; bw <= 64; bw <<= 1)
    for (int bh = 1; bw * bh <= 64; bh <<= 1)
      for (int bd = 1; bw * bh * bd <= 64; bd <<= 1) {
        const int bn = 64 / (bw * bh * bd);
        if (bw > W || bh > H || bd > D || bn > N) continue;
        const long long n = (long long)((W + bw - 1) / bw) * ((H + bh - 1) / bh) * ((D + bd - 1) / bd) * ((N + bn - 1) / bn);
        // ties: the wider box along W (neighbouring voxels are neighbours in memory)
        if (best < 0 || n < best || (n == best && bw > box[0])) { best = n; box[0] = bw; box[1] = bh; box[2] = bd; box[3] = bn; }
      }
  if (best < 0 || best >= (1ll << 30)) return false;
  *cnt = (int)best;
  return true;
}

static void head_tc_plan(int C, int n_boxes, int* S, int* bps) {
  const int tiles = (C + 127) / 128;
  int s = persistent_sms() / (tiles * 3);
  const int by_work = (n_boxes + 7) / 8;
  if (s > by_work) s = by_work;
  if (s < 1) s = 1;
  *bps = (n_boxes + s - 1) / s;
  *S = (n_boxes + *bps - 1) / *bps;
}

// one 16-column group.  0: done (dw_loc / dw_cls rows of this group written), -1: shape not taken, < -1: error
int head_wgrad_tc_launch(const void* dO16, const void* x, int N, int C, int D, int H, int W, int g, int n_loc, int n_cls,
                         float* dw_loc, float* dw_cls, float* workspace, int64_t workspace_bytes, cudaStream_t st) {
  if (C <= 0 || (C % 64)) return -1;
  HeadWgradTcParams p{};
  int box[4];
  if (!head_tc_box(N, D, H, W, box, &p.n_boxes)) return -1;
  int S;
  head_tc_plan(C, p.n_boxes, &S, &p.boxes_per_split);
  if (workspace_bytes < (int64_t)S * 3 * C * 144 * 4) return -1;
  p.C = C;
  p.bw = box[0]; p.bh = box[1]; p.bd = box[2]; p.bn = box[3];
  p.tw = (W + p.bw - 1) / p.bw; p.th = (H + p.bh - 1) / p.bh; p.td = (D + p.bd - 1) / p.bd;
  p.slab = workspace;
  CUtensorMap tmX, tmG;
  {
    const uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
    const uint64_t strides[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2,
                                 (uint64_t)D * H * W * C * 2};
    const uint32_t bx[5] = {64u, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bd, (uint32_t)p.bn};
    if (make_tma_bf16(&tmX, x, 5, dims, strides, bx, CU_TENSOR_MAP_SWIZZLE_128B)) return -2;
  }
  {
    const uint64_t dims[5] = {16u, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
    const uint64_t strides[4] = {32u, (uint64_t)W * 32, (uint64_t)H * W * 32, (uint64_t)D * H * W * 32};
    const uint32_t bx[5] = {16u, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bd, (uint32_t)p.bn};
    if (make_tma_bf16(&tmG, dO16, 5, dims, strides, bx, CU_TENSOR_MAP_SWIZZLE_32B)) return -2;
  }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(head_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_SMEM) != cudaSuccess)
      return -3;
    attr_set = true;
  }
  const dim3 grid((unsigned)((C + 127) / 128), 3u, (unsigned)S);
  if (launch_pdl(head_wgrad_tc_kernel, grid, dim3(WT_THREADS), (size_t)HT_SMEM, st, tmX, tmG, p) != cudaSuccess) return -4;
  const long long total = 27ll * C * 16;
  if (launch_pdl(head_wsum_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, (const float*)workspace, S,
                 C, g, n_loc, n_cls, dw_loc, dw_cls) != cudaSuccess)
    return -4;
  return 0;
}


// 128-voxel box for the data gradient (same rule as head_tc_box); log2 extents
static bool head_dgrad_box(int N, int D, int H, int W, int* lg, long long* cnt) {
  long long best = -1;
  for (int lw = 0; lw <= 7; ++lw)
    for (int lh = 0; lw + lh <= 7; ++lh)
      for (int ld = 0; lw + lh + ld <= 7; ++ld) {
        const int bw = 1 << lw, bh = 1 << lh, bd = 1 << ld, bn = 128 >> (lw + lh + ld);
        if (bw > W || bh > H || bd > D || bn > N) continue;
        const long long n = (long long)((W + bw - 1) / bw) * ((H + bh - 1) / bh) * ((D + bd - 1) / bd) * ((N + bn - 1) / bn);
        if (best < 0 || n < best || (n == best && lw > lg[0])) { best = n; lg[0] = lw; lg[1] = lh; lg[2] = ld; }
      }
  if (best < 0 || best >= (1ll << 31)) return false;
  *cnt = best;
  return true;
}

// 0: launched, -1: shape not taken (no 128-voxel box / channel count), < -1: error
int head_dgrad_tc_launch(const void* dO, const void* w, const void* addend, void* dx, int N, int C, int D, int H, int W,
                         int groups, cudaStream_t st) {
  if (C <= 0 || (C % 64) || groups <= 0) return -1;
  int lg[3] = {0, 0, 0};
  long long boxes;
  if (!head_dgrad_box(N, D, H, W, lg, &boxes)) return -1;
  HeadDgradTcParams p{};
  p.N = N; p.D = D; p.H = H; p.W = W; p.C = C;
  p.NT = (C % 128 == 0) ? 128 : 64;
  p.groups = groups;
  p.lw = lg[0]; p.lh = lg[1]; p.ld = lg[2];
  p.tw = (W + (1 << p.lw) - 1) >> p.lw; p.th = (H + (1 << p.lh) - 1) >> p.lh; p.td = (D + (1 << p.ld) - 1) >> p.ld;
  p.addend = static_cast<const bf16*>(addend);
  p.dx = static_cast<bf16*>(dx);
  CUtensorMap tmG, tmW;
  {
    const uint64_t dims[5] = {16u, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N * groups};
    const uint64_t strides[4] = {32u, (uint64_t)W * 32, (uint64_t)H * W * 32, (uint64_t)D * H * W * 32};
    const uint32_t bx[5] = {16u, 1u << p.lw, 1u << p.lh, 1u << p.ld, 128u >> (p.lw + p.lh + p.ld)};
    if (make_tma_bf16(&tmG, dO, 5, dims, strides, bx, CU_TENSOR_MAP_SWIZZLE_32B)) return -2;
  }
  {
    const uint64_t dims[2] = {(uint64_t)27 * C, (uint64_t)16 * groups};
    const uint64_t strides[1] = {(uint64_t)27 * C * 2};
    const uint32_t bx[2] = {64u, 16u};
    if (make_tma_bf16(&tmW, w, 2, dims, strides, bx, CU_TENSOR_MAP_SWIZZLE_128B)) return -2;
  }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(head_dgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HD_SMEM) != cudaSuccess)
      return -3;
    attr_set = true;
  }
  const dim3 grid((unsigned)boxes, (unsigned)(C / p.NT));
  if (launch_pdl(head_dgrad_tc_kernel, grid, dim3(WT_THREADS), (size_t)HD_SMEM, st, tmG, tmW, p) != cudaSuccess) return -4;
  return 0;
}

}  // namespace ssd3d
