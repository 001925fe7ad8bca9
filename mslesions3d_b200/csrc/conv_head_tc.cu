// SSD head convolutions (loc + class 3x3x3 convs of one feature map, bias, written straight into the
// concatenated (N,P,6)/(N,P,n_classes) outputs; ssd3d.py:131-167) as a tcgen05 implicit GEMM that loads every
// input voxel ONCE per 64-channel chunk.
//
// One CTA owns an output box of TD x TH x TW voxels.  Its input halo box (TD+2)(TH+2)(TW+2) voxels x 64
// channels is fetched by a single 5-D TMA load into 128B-swizzled shared memory, one 128-byte row per voxel,
// rows in flat (d,h,w) order; voxels outside the volume are zero-filled by TMA (= the conv's padding).  In that
// flat order the input of output row f for tap (kd,kh,kw) is simply row f + delta(tap), so the A operand of
// every tap is the SAME shared-memory tile addressed through a row-shifted UMMA descriptor: 27 taps x 4 k-steps
// x J row blocks of UMMA (M=128, N=NPAD, K=16) per chunk, no im2col and no re-load.  Rows of the flat range that
// fall on halo positions compute garbage that the epilogue never stores.
//
// Warp roles (224 threads): 0-3 epilogue (TMEM lane quarter = warp), 4 activation-TMA producer, 5 weight-TMA
// producer, 6 TMEM allocator + UMMA issuer.  Activation chunks are double buffered, weight tiles (NPAD x 64,
// one per tap) stream through an 8-deep ring.  Small feature maps are split over K (64-channel chunks) across
// CTAs; partial sums go to a workspace and a second kernel reduces them in a fixed order (deterministic).
#include "common.cuh"
#include "tma_host.h"

namespace ssd3d {

constexpr int HB_RING = 8;

struct Head2Params {
  int C, D, H, W, N;
  int TD, TH, TW, HD, HH, HW;
  int tiles_w, tiles_h, tiles_d, tiles_total;
  int J;               // 128-row blocks per tile
  int f0;              // flat halo index of the first interior voxel
  int rows_alloc;      // shared-memory rows per activation buffer (multiple of 8)
  int chunks, S, cps;  // 64-channel chunks, K splits, chunks per split
  int nbuf;            // activation buffers (1 or 2)
  int NPAD, n_loc, n_cls, bpl, n_classes;
  int tmem_cols;
  long long P, prior_off;
  float* locs;
  float* scores;
  const float* bias;
  float* partial;      // (S, tiles_total, J*128, NPAD) fp32 when S > 1
  int* nan_flag;
};

__global__ void __launch_bounds__(224, 1) head2_kernel(const __grid_constant__ CUtensorMap tmX,
                                                       const __grid_constant__ CUtensorMap tmW, const Head2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const int A_BYTES = p.rows_alloc * 128;
  const int B_BYTES = p.NPAD * 128;
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)p.nbuf * A_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + (size_t)HB_RING * B_BYTES);
  uint64_t* a_empty = a_full + 2;
  uint64_t* b_full = a_empty + 2;
  uint64_t* b_empty = b_full + HB_RING;
  uint64_t* acc_full = b_empty + HB_RING;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  int t = blockIdx.x;
  const int tw0 = (t % p.tiles_w) * p.TW; t /= p.tiles_w;
  const int th0 = (t % p.tiles_h) * p.TH; t /= p.tiles_h;
  const int td0 = (t % p.tiles_d) * p.TD; t /= p.tiles_d;
  const int n = t;
  const int split = blockIdx.y;
  const int chunk0 = split * p.cps;
  const int nchunk = min(p.cps, p.chunks - chunk0);

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < HB_RING; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 6) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int R_BYTES = p.HD * p.HH * p.HW * 128;   // bytes one halo box delivers
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 4) {
    // ===================== activation producer: one 5-D box per 64-channel chunk =====================
    if (lane == 0) {
      for (int ci = 0; ci < nchunk; ++ci) {
        const int buf = ci % p.nbuf, use = ci / p.nbuf;
        if (use > 0) mbar_wait(&a_empty[buf], (uint32_t)((use - 1) & 1));
        mbar_arrive_expect_tx(&a_full[buf], (uint32_t)R_BYTES);
        tma_load_5d(sA + (size_t)buf * A_BYTES, &tmX, &a_full[buf], (chunk0 + ci) * 64, tw0 - 1, th0 - 1, td0 - 1, n);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ===================== weight producer: NPAD x 64 tile per (chunk, tap) =====================
    if (lane == 0) {
      int it = 0;
      for (int ci = 0; ci < nchunk; ++ci) {
        for (int tap = 0; tap < 27; ++tap, ++it) {
          const int s = it % HB_RING, use = it / HB_RING;
          if (use > 0) mbar_wait(&b_empty[s], (uint32_t)((use - 1) & 1));
          mbar_arrive_expect_tx(&b_full[s], (uint32_t)B_BYTES);
          tma_load_2d(sB + (size_t)s * B_BYTES, &tmW, &b_full[s], tap * p.C + (chunk0 + ci) * 64, 0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 6) {
    // ===================== UMMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, p.NPAD);
      int it = 0;
      for (int ci = 0; ci < nchunk; ++ci) {
        const int buf = ci % p.nbuf;
        mbar_wait(&a_full[buf], (uint32_t)((ci / p.nbuf) & 1));
        const uint32_t a_base = smem_u32(sA + (size_t)buf * A_BYTES);
        for (int tap = 0; tap < 27; ++tap, ++it) {
          const int s = it % HB_RING;
          mbar_wait(&b_full[s], (uint32_t)((it / HB_RING) & 1));
          tc_fence_after();
          const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
          const int delta = ((kd - 1) * p.HH + (kh - 1)) * p.HW + (kw - 1);
          const uint64_t db = umma_desc_k_sw128(smem_u32(sB + (size_t)s * B_BYTES));
          for (int j = 0; j < p.J; ++j) {
            // rows f0 + 128 j + delta .. +127 of the flat halo tile: a row-shifted view of the same buffer
            const uint64_t da = umma_desc_k_sw128(a_base + (uint32_t)((p.f0 + 128 * j + delta) * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ss(tmem_base + (uint32_t)(j * p.NPAD), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                           (ci | tap | k) != 0 ? 1u : 0u);
          }
          umma_commit(&b_empty[s]);
        }
        umma_commit(&a_empty[buf]);
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else {
    // ===================== epilogue =====================
    mbar_wait(acc_full, 0);
    __syncwarp();
    tc_fence_after();
    bool bad_l = false, bad_s = false;
    for (int j = 0; j < p.J; ++j) {
      const int m = j * 128 + warp * 32 + lane;
      const int f = p.f0 + m;
      const int hw = f % p.HW, hh = (f / p.HW) % p.HH, hd = f / (p.HW * p.HH);
      const int w = tw0 + hw - 1, h = th0 + hh - 1, d = td0 + hd - 1;
      const bool valid = hw >= 1 && hw <= p.TW && hh >= 1 && hh <= p.TH && hd >= 1 && hd <= p.TD && w < p.W &&
                         h < p.H && d < p.D;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * p.NPAD);
      for (int c = 0; c < p.NPAD; c += 16) {
        uint32_t v[16];
        __syncwarp();           // tcgen05.ld is .sync.aligned; lanes diverge on `valid` below
        tmem_ld_32x32b_x16(taddr + (uint32_t)c, v);
        tmem_ld_wait();
        if (!valid) continue;
        if (p.S == 1) {
          const long long prior = p.prior_off + (((long long)d * p.H + h) * p.W + w) * p.bpl;
          float* lp = p.locs + ((long long)n * p.P + prior) * 6;
          float* sp = p.scores + ((long long)n * p.P + prior) * p.n_classes;
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int col = c + q;
            const float val = __fadd_rn(__uint_as_float(v[q]), __ldg(p.bias + col));
            if (col < p.n_loc) {
              lp[col] = val;
              bad_l |= (val != val);
            } else if (col < p.n_loc + p.n_cls) {
              sp[col - p.n_loc] = val;
              bad_s |= (val != val);
            }
          }
        } else {
          float4* dst = reinterpret_cast<float4*>(
              p.partial + ((((long long)split * p.tiles_total + blockIdx.x) * (p.J * 128) + m) * p.NPAD + c));
          dst[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
          dst[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
          dst[2] = make_float4(__uint_as_float(v[8]), __uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
          dst[3] = make_float4(__uint_as_float(v[12]), __uint_as_float(v[13]), __uint_as_float(v[14]), __uint_as_float(v[15]));
        }
      }
    }
    if (p.nan_flag) {
      if (bad_l) atomicOr(p.nan_flag, SSD3D_NAN_LOCS);
      if (bad_s) atomicOr(p.nan_flag, SSD3D_NAN_SCORES);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 6) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// out[voxel][col] = bias[col] + sum over splits (ascending) of partial[split][tile][m][col].
// One thread per (voxel, 4 columns): the S float4 loads are independent, the additions run in split order.
__global__ void __launch_bounds__(256) head2_reduce_kernel(const Head2Params p) {
  pdl_wait();
  pdl_launch_dependents();
  const int groups = p.NPAD >> 2;
  const long long total = (long long)p.N * p.D * p.H * p.W * groups;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int c0 = (int)(gid % groups) * 4;
  long long r = gid / groups;
  const int w = (int)(r % p.W); r /= p.W;
  const int h = (int)(r % p.H); r /= p.H;
  const int d = (int)(r % p.D);
  const int n = (int)(r / p.D);
  const int ncol = p.n_loc + p.n_cls;
  if (c0 >= ncol) return;
  const int tile = ((n * p.tiles_d + d / p.TD) * p.tiles_h + h / p.TH) * p.tiles_w + w / p.TW;
  const int f = ((d % p.TD + 1) * p.HH + (h % p.TH + 1)) * p.HW + (w % p.TW + 1);
  const int m = f - p.f0;
  const long long split_stride = (long long)p.tiles_total * (p.J * 128) * p.NPAD;
  const float* src = p.partial + ((long long)tile * (p.J * 128) + m) * p.NPAD + c0;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int s = 0; s < p.S; ++s) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + s * split_stride));
    acc[0] = __fadd_rn(acc[0], v.x); acc[1] = __fadd_rn(acc[1], v.y);
    acc[2] = __fadd_rn(acc[2], v.z); acc[3] = __fadd_rn(acc[3], v.w);
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

static inline int hp2ceil(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

// tile / split plan shared by the workspace query and the launcher
static bool head2_plan(int N, int C, int D, int H, int W, int NPAD, Head2Params& p) {
  if (C % 64 != 0 || NPAD % 16 != 0 || NPAD > 64) return false;
  p.C = C; p.D = D; p.H = H; p.W = W; p.N = N; p.NPAD = NPAD;
  p.chunks = C / 64;
  p.TW = hp2ceil(W) < 8 ? hp2ceil(W) : 8;
  p.TH = hp2ceil(H) < 8 ? hp2ceil(H) : 8;
  p.TD = hp2ceil(D) < 4 ? hp2ceil(D) : 4;
  auto ntiles = [&]() {
    return (long long)((W + p.TW - 1) / p.TW) * ((H + p.TH - 1) / p.TH) * ((D + p.TD - 1) / p.TD) * N;
  };
  // small maps: shrink the tile until tiles x chunks can occupy the 148 SMs (never below 2 x 2 x TW)
  while (ntiles() * p.chunks < 148) {
    if (p.TD > 2 && p.TD >= p.TH) p.TD >>= 1;
    else if (p.TH > 2) p.TH >>= 1;
    else if (p.TD > 1) p.TD >>= 1;
    else break;
  }
  p.HD = p.TD + 2; p.HH = p.TH + 2; p.HW = p.TW + 2;
  p.tiles_w = (W + p.TW - 1) / p.TW;
  p.tiles_h = (H + p.TH - 1) / p.TH;
  p.tiles_d = (D + p.TD - 1) / p.TD;
  p.tiles_total = (int)ntiles();
  p.f0 = (p.HH + 1) * p.HW + 1;
  const int span = ((p.TD - 1) * p.HH + (p.TH - 1)) * p.HW + p.TW;
  p.J = (span + 127) / 128;
  const int dmax = (p.HH + 1) * p.HW + 1;
  int rows = p.f0 + 128 * p.J + dmax + 1;
  const int R = p.HD * p.HH * p.HW;
  if (rows < R) rows = R;
  p.rows_alloc = (rows + 7) & ~7;
  // K split: enough CTAs to fill the machine, at most one split per chunk
  p.S = 1;
  if (p.tiles_total < 128) {
    long long want = (296 + p.tiles_total - 1) / p.tiles_total;
    p.S = (int)(want < p.chunks ? want : p.chunks);
    if (p.S < 1) p.S = 1;
  }
  p.cps = (p.chunks + p.S - 1) / p.S;
  p.S = (p.chunks + p.cps - 1) / p.cps;
  p.nbuf = p.cps > 1 ? 2 : 1;
  int cols = p.J * NPAD;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols <<= 1;
  if (p.tmem_cols > 512) return false;
  const size_t smem = 1024 + (size_t)p.nbuf * p.rows_alloc * 128 + (size_t)HB_RING * NPAD * 128 + 256;
  if (smem > 225 * 1024) {
    p.nbuf = 1;
    const size_t smem1 = 1024 + (size_t)p.rows_alloc * 128 + (size_t)HB_RING * NPAD * 128 + 256;
    if (smem1 > 225 * 1024) return false;
  }
  return true;
}

}  // namespace ssd3d

using namespace ssd3d;

// conv_head_kw.cu
int64_t ssd3d_head_kw_workspace_bytes(int N, int C, int D, int H, int W);
bool ssd3d_head_kw_applicable(int N, int C, int D, int H, int W, int NPAD);

extern "C" int64_t ssd3d_head_workspace_bytes(int N, int C, int D, int H, int W, int NPAD) {
  Head2Params p{};
  if (N <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
  int64_t need = 0;
  if (head2_plan(N, C, D, H, W, NPAD, p) && p.S > 1) need = (int64_t)p.S * p.tiles_total * p.J * 128 * NPAD * 4;
  if (ssd3d_head_kw_applicable(N, C, D, H, W, NPAD)) {
    const int64_t kw = ssd3d_head_kw_workspace_bytes(N, C, D, H, W);
    if (kw > need) need = kw;
  }
  return need;
}

// returns SSD3D_ERR_UNSUPPORTED when the shape is outside this kernel (caller falls back to the per-tap kernel)
int ssd3d_head_conv_halo(const void* x, const void* w, const float* bias, float* locs, float* scores, int N, int C,
                         int D, int H, int W, int bpl, int n_classes, int NPAD, int64_t P, int64_t prior_offset,
                         int* nan_flag, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  Head2Params p{};
  if (!head2_plan(N, C, D, H, W, NPAD, p)) return SSD3D_ERR_UNSUPPORTED;
  p.bpl = bpl; p.n_classes = n_classes; p.n_loc = bpl * 6; p.n_cls = bpl * n_classes;
  p.P = P; p.prior_off = prior_offset; p.locs = locs; p.scores = scores; p.bias = bias; p.nan_flag = nan_flag;
  if (p.S > 1) {
    const int64_t need = (int64_t)p.S * p.tiles_total * p.J * 128 * NPAD * 4;
    if (!workspace || workspace_bytes < need) return SSD3D_ERR_ARG;
    p.partial = static_cast<float*>(workspace);
  }
  CUtensorMap tmX, tmW;
  {
    const uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
    const uint64_t strides[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2,
                                 (uint64_t)D * H * W * C * 2};
    const uint32_t box[5] = {64u, (uint32_t)p.HW, (uint32_t)p.HH, (uint32_t)p.HD, 1u};
    if (make_tma_bf16(&tmX, x, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return SSD3D_ERR_TMA;
  }
  {
    const uint64_t dims[2] = {(uint64_t)27 * C, (uint64_t)NPAD};
    const uint64_t strides[1] = {(uint64_t)27 * C * 2};
    const uint32_t box[2] = {64u, (uint32_t)NPAD};
    if (make_tma_bf16(&tmW, w, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return SSD3D_ERR_TMA;
  }
  const size_t smem = 1024 + (size_t)p.nbuf * p.rows_alloc * 128 + (size_t)HB_RING * NPAD * 128 + 256;
  cudaError_t e = cudaFuncSetAttribute(head2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((unsigned)p.tiles_total, (unsigned)p.S);
  SSD3D_LAUNCH_PDL(head2_kernel, grid, dim3(224), smem, st, tmX, tmW, p);
  if (p.S > 1) {
    const long long total = (long long)N * D * H * W * (NPAD / 4);
    SSD3D_LAUNCH_PDL(head2_reduce_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, p);
  }
  return SSD3D_OK;
}
