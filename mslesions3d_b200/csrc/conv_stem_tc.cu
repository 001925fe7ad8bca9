// Stem convolution (dense 3x3x3, Cin in {1..4} -> 32, stride (sd,2,2), + BN + ReLU) as a tcgen05 implicit
// GEMM (mobilenet.py:26-31 as instantiated at ssd3d.py:60-61).
//
// On CUDA cores this layer is compute-bound (1728 FMA per output voxel at Cin=2: ~300 us for the 2ch 128^3
// batch-8 workload against a 31 us HBM floor), so the contraction goes to the tensor cores:
//   * one CTA = 128 output voxels (TW x TH x TD box) x 32 channels = one UMMA M=128, N=32 accumulator;
//   * the input halo box of the NCDHW tensor is fetched by ONE 4-D TMA load; voxels outside the volume are
//     zero-filled by the TMA unit, which is the conv's zero padding;
//   * each thread gathers its voxel's 27*Cin taps from the shared-memory halo into one K-major row of the
//     A operand, written in the 128B-swizzle layout the UMMA descriptor expects (K padded to 64/128);
//   * 4 (or 8) UMMAs accumulate into 32 TMEM columns; the epilogue applies scale/shift/ReLU and writes
//     64 contiguous bytes per voxel (channels-last bf16).
// Several CTAs are resident per SM (about 30 KB of shared memory each), so TMA, gather, MMA and store of
// neighbouring tiles overlap without an intra-CTA pipeline.
#include "common.cuh"
#include "tma_host.h"

namespace ssd3d {

struct StemParams {
  int N, D, H, W, Do, Ho, Wo, sd;
  int TW, TH, TD;            // output tile (product 128)
  int TWI, THI, TDI;         // input halo tile (TWI padded so that a row is a multiple of 16 bytes)
  int tiles_w, tiles_h, tiles_d;
  const __nv_bfloat16* wt;   // (32, KPAD) bf16, k = ((ci*3+kd)*3+kh)*3+kw, zero padded
  const float* scale;
  const float* shift;
  __nv_bfloat16* y;          // (N, Do, Ho, Wo, 32)
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

__device__ __forceinline__ uint32_t bf16_bits(float f) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(f));
}

// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a 128B-swizzled K-major tile
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) {
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}

template <typename TIn, int CIN>
__global__ void __launch_bounds__(128) stem_tc_kernel(const __grid_constant__ CUtensorMap tmX, const StemParams p) {
  constexpr int KREAL = 27 * CIN;
  constexpr int KPAD = (KREAL <= 64) ? 64 : 128;
  constexpr int KB = KPAD / 64;
  constexpr int NREG = KPAD / 2;          // packed bf16 pairs per A row
  constexpr int PADL = 16 / (int)sizeof(TIn);   // left halo rounded up to 16 bytes (1 column is needed)

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* sA = smem;                                   // KB x (128 rows x 128 B)
  uint8_t* sB = sA + KB * 16384;                        // KB x (32 rows x 128 B)
  uint8_t* ctl = sB + KB * 4096;                        // mbarriers + TMEM slot (128 B)
  uint8_t* sX = ctl + 128;                              // input halo tile [CIN][TDI][THI][TWI] of TIn
  const int tile_elems = CIN * p.TDI * p.THI * p.TWI;
  uint64_t* bar_in = reinterpret_cast<uint64_t*>(ctl);
  uint64_t* bar_mma = bar_in + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_in + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;

  int t = blockIdx.x;
  const int w0 = (t % p.tiles_w) * p.TW; t /= p.tiles_w;
  const int h0 = (t % p.tiles_h) * p.TH; t /= p.tiles_h;
  const int d0 = (t % p.tiles_d) * p.TD; t /= p.tiles_d;
  const int n = t;

  // barrier init and TMA issue live in warp 1, so that warp 0 reaches the .sync.aligned TMEM
  // allocation fully converged (a lane still inside a divergent branch makes it an illegal instruction)
  if (tid == 32) {
    tma_prefetch_desc(&tmX);
    mbar_init(bar_in, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (tid == 32) {
    mbar_arrive_expect_tx(bar_in, (uint32_t)(tile_elems * sizeof(TIn)));
    // TMA needs the box start to be 16-byte aligned along the innermost dimension (a start at column
    // 2*w0 - 1 is an illegal instruction): fetch from 2*w0 - PADL, PADL = 16 bytes of elements
    tma_load_4d(sX, &tmX, bar_in, 2 * w0 - PADL, 2 * h0 - 1, p.sd * d0 - 1, n * CIN);
  }

  // weights -> swizzled B operand while the TMA is in flight: 32 rows x (KPAD/8) 16-byte chunks
  for (int c = tid; c < 32 * (KPAD / 8); c += 128) {
    const int row = c / (KPAD / 8), ch = c % (KPAD / 8);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.wt + row * KPAD + ch * 8));
    *reinterpret_cast<uint4*>(sB + (ch >> 3) * 4096 + sw128_offset(row, ch & 7)) = v;
  }

  // this thread's output voxel inside the tile (w fastest)
  const int wl = tid % p.TW;
  const int hl = (tid / p.TW) % p.TH;
  const int dl = tid / (p.TW * p.TH);

  mbar_wait(bar_in, 0);

  // ---- gather the 27*CIN taps of this voxel into one K-major row (bf16 pairs in registers) ----
  uint32_t regs[NREG];
#pragma unroll
  for (int i = 0; i < NREG; ++i) regs[i] = 0u;
  const TIn* xt = reinterpret_cast<const TIn*>(sX);
#pragma unroll
  for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int g = (ci * 3 + kd) * 3 + kh;
        // taps kw = 0..2 sit at tile columns 2*wl + PADL - 1 .. 2*wl + PADL + 1
        const TIn* src = xt + ((ci * p.TDI + (p.sd * dl + kd)) * p.THI + (2 * hl + kh)) * p.TWI + 2 * wl + PADL;
        uint32_t e0, e1, e2;
        if constexpr (sizeof(TIn) == 2) {
          const uint32_t a = *reinterpret_cast<const uint32_t*>(src - 2);   // columns -2, -1
          const uint32_t b = *reinterpret_cast<const uint32_t*>(src);       // columns  0, +1
          e0 = a >> 16; e1 = b & 0xffffu; e2 = b >> 16;
        } else {
          const float a = reinterpret_cast<const float*>(src)[-1];
          const float2 b = *reinterpret_cast<const float2*>(src);
          e0 = bf16_bits(a); e1 = bf16_bits(b.x); e2 = bf16_bits(b.y);
        }
        const int k0 = g * 3;
        regs[(k0 + 0) >> 1] |= e0 << (16 * ((k0 + 0) & 1));
        regs[(k0 + 1) >> 1] |= e1 << (16 * ((k0 + 1) & 1));
        regs[(k0 + 2) >> 1] |= e2 << (16 * ((k0 + 2) & 1));
      }
    }
  }
#pragma unroll
  for (int j = 0; j < KPAD / 8; ++j) {
    *reinterpret_cast<uint4*>(sA + (j >> 3) * 16384 + sw128_offset(tid, j & 7)) =
        make_uint4(regs[4 * j], regs[4 * j + 1], regs[4 * j + 2], regs[4 * j + 3]);
  }
  fence_proxy_async_smem();     // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  __syncthreads();

  if (tid == 32) {
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 32);
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
      const uint64_t da = umma_desc_k_sw128(smem_u32(sA + kb * 16384));
      const uint64_t db = umma_desc_k_sw128(smem_u32(sB + kb * 4096));
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_ss(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
    }
    umma_commit(bar_mma);
  }
  __syncwarp();
  mbar_wait(bar_mma, 0);
  __syncwarp();                 // tcgen05.ld is .sync.aligned: reconverge after the per-thread poll loop
  tc_fence_after();

  // ---- epilogue: TMEM lane = tile row = this thread's voxel ----
  uint32_t v0[16], v1[16];
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
  tmem_ld_32x32b_x16(taddr, v0);
  tmem_ld_32x32b_x16(taddr + 16, v1);
  tmem_ld_wait();
  const int wo = w0 + wl, ho = h0 + hl, dz = d0 + dl;
  if (wo < p.Wo && ho < p.Ho && dz < p.Do) {
    uint4* dst = reinterpret_cast<uint4*>(p.y + ((((long long)n * p.Do + dz) * p.Ho + ho) * p.Wo + wo) * 32);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t o[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const int c = q * 8 + h * 2;
        const float a0 = __uint_as_float(c < 16 ? v0[c & 15] : v1[c & 15]);
        const float a1 = __uint_as_float(c + 1 < 16 ? v0[(c + 1) & 15] : v1[(c + 1) & 15]);
        const float2 sc = __ldg(reinterpret_cast<const float2*>(p.scale + c));
        const float2 sh = __ldg(reinterpret_cast<const float2*>(p.shift + c));
        o[h] = pack_bf16x2(relu_nan(__fadd_rn(__fmul_rn(a0, sc.x), sh.x)),
                           relu_nan(__fadd_rn(__fmul_rn(a1, sc.y), sh.y)));
      }
      dst[q] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

static inline int p2ceil(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

template <typename TIn, int CIN>
static int launch_stem_tc(const void* x, const StemParams& p0, CUtensorMapDataType dt, cudaStream_t st) {
  StemParams p = p0;
  constexpr int KPAD = (27 * CIN <= 64) ? 64 : 128;
  constexpr int KB = KPAD / 64;
  CUtensorMap tm;
  {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return SSD3D_ERR_TMA;
    const cuuint64_t es = sizeof(TIn);
    cuuint64_t gdim[4] = {(cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.D, (cuuint64_t)p.N * CIN};
    cuuint64_t gstr[3] = {(cuuint64_t)p.W * es, (cuuint64_t)p.W * p.H * es, (cuuint64_t)p.W * p.H * p.D * es};
    cuuint32_t box[4] = {(cuuint32_t)p.TWI, (cuuint32_t)p.THI, (cuuint32_t)p.TDI, (cuuint32_t)CIN};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, dt, 4, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return SSD3D_ERR_TMA;
  }
  const size_t tile_bytes = ((size_t)CIN * p.TDI * p.THI * p.TWI * sizeof(TIn) + 15) & ~(size_t)15;
  const size_t smem = 1024 + (size_t)KB * (16384 + 4096) + 128 + tile_bytes + 256;
  cudaError_t e = cudaFuncSetAttribute(stem_tc_kernel<TIn, CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const unsigned grid = (unsigned)(p.tiles_w * p.tiles_h * p.tiles_d * p.N);
  stem_tc_kernel<TIn, CIN><<<grid, 128, smem, st>>>(tm, p);
  SSD3D_CHECK_LAUNCH();
  return SSD3D_OK;
}

}  // namespace ssd3d

using namespace ssd3d;

// 1 when the tensor-core stem can take this input (TMA needs 16-byte row strides), else 0
extern "C" int ssd3d_stem_tc_supported(int x_is_bf16, int Cin, int W) {
  if (Cin < 1 || Cin > 4) return 0;
  return ((W * (x_is_bf16 ? 2 : 4)) % 16 == 0) ? 1 : 0;
}

static int stem_tc(const void* x, int x_is_bf16, const void* w_tc, const float* scale, const float* shift, void* y,
                   int N, int Cin, int D, int H, int W, int stride_d, cudaStream_t st) {
  StemParams p{};
  p.N = N; p.D = D; p.H = H; p.W = W; p.sd = stride_d;
  p.Do = (D - 1) / stride_d + 1; p.Ho = (H - 1) / 2 + 1; p.Wo = (W - 1) / 2 + 1;
  // output tile: widest W extent (<= 64) that wastes < 15 % of its columns, then H, then D
  int tw = 8;
  for (int c : {64, 32, 16, 8}) {
    if (c > p2ceil(p.Wo)) continue;
    const int padded = ((p.Wo + c - 1) / c) * c;
    if (padded * 100 <= p.Wo * 115 || c == 8) { tw = c; break; }
  }
  if (p2ceil(p.Wo) < tw) tw = p2ceil(p.Wo);
  p.TW = tw;
  int rest = 128 / p.TW;
  p.TH = p2ceil(p.Ho) < rest ? p2ceil(p.Ho) : rest;
  rest /= p.TH;
  p.TD = rest;                                 // whatever remains goes to D (rows past Do are masked)
  const int padl = x_is_bf16 ? 8 : 4;          // 16 bytes of elements on the left (see the kernel)
  p.TWI = 2 * p.TW + padl;                      // columns 2*w0 - padl .. 2*w0 + 2*TW - 1: a multiple of 16 bytes
  // every tile's first column 2*k*TW - padl must stay 16-byte aligned
  if (p.Wo > p.TW && ((2 * p.TW * (x_is_bf16 ? 2 : 4)) % 16) != 0) return SSD3D_ERR_UNSUPPORTED;
  p.THI = 2 * p.TH + 1;
  p.TDI = stride_d * (p.TD - 1) + 3;
  if (p.TWI > 256 || p.THI > 256 || p.TDI > 256) return SSD3D_ERR_UNSUPPORTED;
  p.tiles_w = (p.Wo + p.TW - 1) / p.TW;
  p.tiles_h = (p.Ho + p.TH - 1) / p.TH;
  p.tiles_d = (p.Do + p.TD - 1) / p.TD;
  p.wt = static_cast<const __nv_bfloat16*>(w_tc);
  p.scale = scale;
  p.shift = shift;
  p.y = static_cast<__nv_bfloat16*>(y);
#define STEM_CASE(T, DT, C) return launch_stem_tc<T, C>(x, p, DT, st)
  if (x_is_bf16) {
    switch (Cin) {
      case 1: STEM_CASE(__nv_bfloat16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 1);
      case 2: STEM_CASE(__nv_bfloat16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
      case 3: STEM_CASE(__nv_bfloat16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3);
      default: STEM_CASE(__nv_bfloat16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4);
    }
  } else {
    switch (Cin) {
      case 1: STEM_CASE(float, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 1);
      case 2: STEM_CASE(float, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2);
      case 3: STEM_CASE(float, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3);
      default: STEM_CASE(float, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4);
    }
  }
#undef STEM_CASE
}

extern "C" int ssd3d_stem_conv_bn_relu(const void* x, int x_is_bf16, const void* w, const float* scale,
                                       const float* shift, void* y, int N, int Cin, int D, int H, int W, int stride_d,
                                       void* stream) {
  if (!x || !w || !scale || !shift || !y || N <= 0 || D <= 0 || H <= 0 || W <= 0) return SSD3D_ERR_ARG;
  if (stride_d != 1 && stride_d != 2) return SSD3D_ERR_ARG;
  if (Cin < 1 || Cin > 4) return SSD3D_ERR_UNSUPPORTED;
  if (ssd3d_stem_tc_supported(x_is_bf16, Cin, W)) {
    const int rc = stem_tc(x, x_is_bf16, w, scale, shift, y, N, Cin, D, H, W, stride_d,
                           static_cast<cudaStream_t>(stream));
    if (rc != SSD3D_ERR_UNSUPPORTED) return rc;
  }
  // rows that TMA cannot address (W * elemsize not a multiple of 16 bytes): CUDA-core kernel
  return ssd3d_stem_conv_bn_relu_simt(x, x_is_bf16, w, scale, shift, y, N, Cin, D, H, W, stride_d, stream);
}
