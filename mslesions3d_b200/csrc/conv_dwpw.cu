// One MobileNet Block (mobilenet.py:34-49, eval mode) as ONE kernel:
//     depthwise 3x3x3 (+BN1+ReLU)  ->  pointwise 1x1x1 on tcgen05 (+BN2+ReLU)  ->  y
// The depthwise result never goes to HBM: a CTA computes a tile of 128 output voxels x ALL Cin channels with the
// TMA-halo / FFMA2 scheme of conv_dw_tma.cu (one 32-channel chunk at a time), rounds it to bf16 exactly like the
// stand-alone kernel stores it, and writes it straight into a K-major, swizzled shared-memory tile that is the A
// operand of the pointwise GEMM.  The pointwise weights (Cout x Cin) sit in shared memory for the whole kernel, the
// accumulator (128 x Cout fp32) in tensor memory, and the epilogue applies BN2 + ReLU + the NaN flag and stores
// bf16 rows.  Versus the two-kernel path this removes the depthwise output's round trip (f1: 2 x 16.8 MB, f2:
// 2 x 4.2 MB, f3: 2 x 8.4 MB at the benchmark size), one launch per block, and the pointwise kernel's own
// load / fill / drain phases: the GEMM of tile j runs under the depthwise FMAs of tile j + 1.
//
// Warp roles (448 threads):
//   warps 0-7    depthwise: thread = 4 channels x WT outputs along W, weights (fp32) from shared memory, BN1 + ReLU,
//                bf16 pack, 8-byte store into A[j & 1]; then arrive on a_full
//   warp 8       TMA producer: 5-D halo boxes (zero fill = padding), double buffered across chunks AND tiles;
//                the pointwise weight tiles once
//   warp 9       TMEM allocation; UMMA issue: D[j & 1] = A[j & 1] . W2^T  (M 128, N Cout, K Cin), commit -> a_empty
//                and tmem_full
//   warps 10-13  epilogue: tcgen05.ld of the row's Cout columns, BN2 scale / shift, ReLU, NaN check, bf16 stores
// A and the accumulator are double buffered, so depthwise(j+1), GEMM(j) and epilogue(j) overlap.
//
// Arithmetic is the stand-alone kernels': fp32 FMAs in (kd, kh, kw) order, separately rounded scale / shift, bf16
// rounding of the intermediate, fp32 accumulation on the tensor pipe.  Roofline: HBM, bytes = 2*N*(Cin*Vin +
// Cout*Vout) + weights.
#include "common.cuh"
#include "tma_host.h"

namespace ssd3d {

namespace dwpw {

constexpr int CB = 32;            // channels per depthwise chunk (64 bytes per voxel)
constexpr int DW_THREADS = 256;
constexpr int THREADS = 448;

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 bf16x2_to_f32x2(uint32_t u) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(u << 16), "r"(u & 0xffff0000u));
  return r;
}
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi)));
  return r;
}
__device__ __forceinline__ void ffma2(f32x2& acc, f32x2 a, f32x2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  uint32_t a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v));
  lo = __uint_as_float(a);
  hi = __uint_as_float(b);
}

struct Params {
  int N, D, H, W, Do, Ho, Wo;
  int tiles_w, tiles_h, tiles_d, tiles;
  const __nv_bfloat16* w1;      // (27, Cin) bf16
  const float* scale1;
  const float* shift1;
  const float* scale2;
  const float* shift2;
  __nv_bfloat16* y;             // (N, Do, Ho, Wo, Cout)
  int* nan_flag;
};

template <int S, int CIN, int COUT>
struct Cfg {
  static constexpr int WT = (S == 2) ? 2 : 4;
  static constexpr int TD = (S == 2) ? 4 : 2;
  static constexpr int TH = (S == 2) ? 4 : 8;
  static constexpr int TW = 8;
  static constexpr int TDI = S * (TD - 1) + 3;
  static constexpr int THI = S * (TH - 1) + 3;
  static constexpr int TWI0 = S * (TW - 1) + 3;
  static constexpr int TWI = (TWI0 & 1) ? TWI0 : TWI0 + 1;          // odd row pitch (in 64-byte voxels)
  static constexpr int HALO_BYTES = TDI * THI * TWI * CB * 2;
  static constexpr int HALO_PITCH = (HALO_BYTES + 127) & ~127;      // TMA destinations: 128-byte aligned
  static constexpr int WQ = TW / WT;
  static constexpr int ITEMS = TD * TH * WQ * 8;
  static constexpr int NCH = CIN / CB;
  static constexpr int BK = (CIN >= 64) ? 64 : 32;
  static constexpr int NKB = CIN / BK;
  static constexpr int A_BYTES = 128 * CIN * 2;
  static constexpr int B_BYTES = COUT * CIN * 2;
  // depthwise weights in shared memory: fp32 (no conversion in the loop) unless that overflows the 227 KB
  static constexpr bool W1F32 = !(S == 2 && CIN == 64);
  static constexpr int W1_BYTES = CIN * 27 * (W1F32 ? 4 : 2);
  static constexpr int TMEM_COLS = 2 * COUT;
  // layout (1024-aligned base): A[2] | B (swizzled operands: 1024-aligned) | halo[2] | w1 | bn1 scale, shift |
  // bn2 scale, shift | barriers
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = OFF_A + 2 * A_BYTES;
  static constexpr int OFF_HALO = OFF_B + B_BYTES;
  static constexpr int OFF_W1 = OFF_HALO + 2 * HALO_PITCH;
  static constexpr int OFF_BN = OFF_W1 + W1_BYTES;
  static constexpr int OFF_BAR = OFF_BN + (2 * CIN + 2 * COUT) * 4;
  static constexpr size_t SMEM = 1024 + (size_t)OFF_BAR + 16 * 8 + 16;
  static_assert(SMEM <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");
  static_assert(TD * TH * TW == 128, "one UMMA M tile per spatial tile");
  static_assert(CIN % 32 == 0 && COUT % 16 == 0 && COUT <= 256 && TMEM_COLS <= 512, "unsupported channel counts");
  static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS >= 32, "TMEM allocation is a power of two >= 32");
};

template <int BK>
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * BK * 2) >> 4) << 32;            // SBO: bytes between 8-row groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)((BK == 64) ? 2u : 4u) << 61;         // SWIZZLE_128B : SWIZZLE_64B
  return d;
}

// byte offset of channel `c` (a multiple of 4) of row `r` inside the swizzled K-major A tile
template <int BK>
__device__ __forceinline__ uint32_t a_offset(int r, int c) {
  const int kb = c / BK, cb = c % BK;
  const int byte = cb * 2, chunk = byte >> 4, in = byte & 15;
  if (BK == 64) return (uint32_t)(kb * (128 * 128) + r * 128 + ((chunk ^ (r & 7)) << 4) + in);
  return (uint32_t)(r * 64 + ((chunk ^ ((r >> 1) & 3)) << 4) + in);
}

template <int S, int CIN, int COUT>
__global__ void __launch_bounds__(THREADS, 1) block_dwpw_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                 const __grid_constant__ CUtensorMap tmW2, const Params p) {
  using C = Cfg<S, CIN, COUT>;
  constexpr int WT = C::WT, TD = C::TD, TH = C::TH, TW = C::TW, BK = C::BK;
  constexpr int NI = (WT - 1) * S + 3;
  extern __shared__ uint8_t dwpw_raw[];
  const uint32_t raw = smem_u32(dwpw_raw);
  uint8_t* smem = dwpw_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* halo = smem + C::OFF_HALO;
  uint8_t* sA = smem + C::OFF_A;
  uint8_t* sB = smem + C::OFF_B;
  float* sW1 = reinterpret_cast<float*>(smem + C::OFF_W1);       // [chunk][tap][32] fp32, or bf16 (W1F32 = false)
  __nv_bfloat16* sW1h = reinterpret_cast<__nv_bfloat16*>(smem + C::OFF_W1);
  float* sBn = reinterpret_cast<float*>(smem + C::OFF_BN);       // scale1[CIN] shift1[CIN] scale2[COUT] shift2[COUT]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* halo_full = bars;          // [2]
  uint64_t* halo_empty = bars + 2;     // [2]
  uint64_t* a_full = bars + 4;         // [2]
  uint64_t* a_empty = bars + 6;        // [2]
  uint64_t* t_full = bars + 8;         // [2]
  uint64_t* t_empty = bars + 10;       // [2]
  uint64_t* b_full = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW2);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&halo_full[b], 1);
      mbar_init(&halo_empty[b], DW_THREADS);
      mbar_init(&a_full[b], DW_THREADS);
      mbar_init(&a_empty[b], 1);
      mbar_init(&t_full[b], 1);
      mbar_init(&t_empty[b], 128);
    }
    mbar_init(b_full, 1);
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, (uint32_t)C::TMEM_COLS);
    tmem_relinquish();
  }
  // constants (not produced by a preceding kernel): depthwise weights as fp32, folded BN vectors
  for (int i = tid; i < 27 * CIN; i += THREADS) {
    const int t = i / CIN, c = i % CIN;
    if (C::W1F32) sW1[((c / CB) * 27 + t) * CB + (c % CB)] = __bfloat162float(p.w1[i]);
    else sW1h[((c / CB) * 27 + t) * CB + (c % CB)] = p.w1[i];
  }
  for (int i = tid; i < CIN; i += THREADS) { sBn[i] = p.scale1[i]; sBn[CIN + i] = p.shift1[i]; }
  for (int i = tid; i < COUT; i += THREADS) { sBn[2 * CIN + i] = p.scale2[i]; sBn[2 * CIN + COUT + i] = p.shift2[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  auto decode = [&](int tile, int& ow0, int& oh0, int& od0, int& n) {
    int t = tile;
    ow0 = (t % p.tiles_w) * TW; t /= p.tiles_w;
    oh0 = (t % p.tiles_h) * TH; t /= p.tiles_h;
    od0 = (t % p.tiles_d) * TD;
    n = t / p.tiles_d;
  };

  if (warp < 8) {
    // ===================================== depthwise =====================================
    const int cv = tid & 7;
    int j = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++j) {
      const int ab = j & 1;
      if (j >= 2) mbar_wait(&a_empty[ab], (uint32_t)(((j >> 1) - 1) & 1));      // the GEMM of tile j-2 has read A[ab]
      int ow0, oh0, od0, n;
      decode(tile, ow0, oh0, od0, n);
      uint8_t* A = sA + (size_t)ab * C::A_BYTES;
#pragma unroll 1
      for (int ch = 0; ch < C::NCH; ++ch) {
        const int hs = j * C::NCH + ch, hb = hs & 1;
        mbar_wait(&halo_full[hb], (uint32_t)((hs >> 1) & 1));
        const uint8_t* in = halo + (size_t)hb * C::HALO_PITCH;
        const float* wch = sW1 + (size_t)ch * 27 * CB + cv * 4;
        const __nv_bfloat16* wchh = sW1h + (size_t)ch * 27 * CB + cv * 4;
        const float4 sc = *reinterpret_cast<const float4*>(sBn + ch * CB + cv * 4);
        const float4 sh = *reinterpret_cast<const float4*>(sBn + CIN + ch * CB + cv * 4);
#pragma unroll 1
        for (int item = tid; item < C::ITEMS; item += DW_THREADS) {
          int r = item >> 3;
          const int h = r % TH; r /= TH;
          const int wq = r % C::WQ;
          const int d = r / C::WQ;
          if (od0 + d >= p.Do || oh0 + h >= p.Ho || ow0 + wq * WT >= p.Wo) continue;   // rows nobody stores
          f32x2 acc[WT][2];
#pragma unroll
          for (int i = 0; i < WT; ++i) { acc[i][0] = 0ull; acc[i][1] = 0ull; }
          const uint8_t* base = in + ((size_t)(((d * S) * C::THI + h * S) * C::TWI + wq * WT * S) * CB + cv * 4) * 2;
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              const uint8_t* row = base + (size_t)((kd * C::THI + kh) * C::TWI) * CB * 2;
              f32x2 x[NI][2];
#pragma unroll
              for (int i = 0; i < NI; ++i) {
                const uint2 u = *reinterpret_cast<const uint2*>(row + i * CB * 2);
                x[i][0] = bf16x2_to_f32x2(u.x);
                x[i][1] = bf16x2_to_f32x2(u.y);
              }
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                f32x2 w0, w1;
                if (C::W1F32) {
                  const float4 wv = *reinterpret_cast<const float4*>(wch + ((kd * 3 + kh) * 3 + kw) * CB);
                  w0 = pack2(wv.x, wv.y);
                  w1 = pack2(wv.z, wv.w);
                } else {
                  const uint2 wu = *reinterpret_cast<const uint2*>(wchh + ((kd * 3 + kh) * 3 + kw) * CB);
                  w0 = bf16x2_to_f32x2(wu.x);
                  w1 = bf16x2_to_f32x2(wu.y);
                }
#pragma unroll
                for (int ow = 0; ow < WT; ++ow) {
                  ffma2(acc[ow][0], x[ow * S + kw][0], w0);
                  ffma2(acc[ow][1], x[ow * S + kw][1], w1);
                }
              }
            }
          }
          const int row0 = (d * TH + h) * TW + wq * WT;
#pragma unroll
          for (int ow = 0; ow < WT; ++ow) {
            float a0, a1, a2, a3;
            unpack2(acc[ow][0], a0, a1);
            unpack2(acc[ow][1], a2, a3);
            a0 = relu_nan(__fadd_rn(__fmul_rn(a0, sc.x), sh.x));
            a1 = relu_nan(__fadd_rn(__fmul_rn(a1, sc.y), sh.y));
            a2 = relu_nan(__fadd_rn(__fmul_rn(a2, sc.z), sh.z));
            a3 = relu_nan(__fadd_rn(__fmul_rn(a3, sc.w), sh.w));
            *reinterpret_cast<uint2*>(A + a_offset<BK>(row0 + ow, ch * CB + cv * 4)) =
                make_uint2(pack_bf16x2(a0, a1), pack_bf16x2(a2, a3));
          }
        }
        mbar_arrive(&halo_empty[hb]);          // this thread is done reading halo[hb]
      }
      fence_proxy_async_smem();                // generic-proxy writes of A -> visible to the tensor core (async proxy)
      mbar_arrive(&a_full[ab]);
    }
  } else if (warp == 8) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      mbar_arrive_expect_tx(b_full, (uint32_t)C::B_BYTES);
      for (int kb = 0; kb < C::NKB; ++kb) tma_load_2d(sB + (size_t)kb * COUT * BK * 2, &tmW2, b_full, kb * BK, 0);
      int hs = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        int ow0, oh0, od0, n;
        decode(tile, ow0, oh0, od0, n);
        for (int ch = 0; ch < C::NCH; ++ch, ++hs) {
          const int hb = hs & 1;
          if (hs >= 2) mbar_wait(&halo_empty[hb], (uint32_t)(((hs >> 1) - 1) & 1));
          mbar_arrive_expect_tx(&halo_full[hb], (uint32_t)C::HALO_BYTES);
          tma_load_5d(halo + (size_t)hb * C::HALO_PITCH, &tmX, &halo_full[hb], ch * CB, ow0 * S - 1, oh0 * S - 1,
                      od0 * S - 1, n);
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ===================================== UMMA issuer =====================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, COUT);
      mbar_wait(b_full, 0u);
      int j = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++j) {
        const int ab = j & 1;
        mbar_wait(&a_full[ab], (uint32_t)((j >> 1) & 1));
        if (j >= 2) mbar_wait(&t_empty[ab], (uint32_t)(((j >> 1) - 1) & 1));
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(ab * COUT);
        const uint32_t a_addr = smem_u32(sA + (size_t)ab * C::A_BYTES), b_addr = smem_u32(sB);
#pragma unroll
        for (int kb = 0; kb < C::NKB; ++kb) {
          const uint64_t da = smem_desc<BK>(a_addr + kb * 128 * BK * 2), db = smem_desc<BK>(b_addr + kb * COUT * BK * 2);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_ss(acc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&a_empty[ab]);
        umma_commit(&t_full[ab]);
      }
    }
    __syncwarp();
  } else {
    // ===================================== epilogue =====================================
    const int q = warp & 3;                    // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int rw = row % TW, rh = (row / TW) % TH, rd = row / (TW * TH);
    const float* sc2 = sBn + 2 * CIN;
    const float* sh2 = sc2 + COUT;
    bool bad = false;
    int j = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++j) {
      const int ab = j & 1;
      int ow0, oh0, od0, n;
      decode(tile, ow0, oh0, od0, n);
      const int od = od0 + rd, oh = oh0 + rh, ow = ow0 + rw;
      const bool valid = od < p.Do && oh < p.Ho && ow < p.Wo;
      __nv_bfloat16* dst = p.y + ((((long long)n * p.Do + od) * p.Ho + oh) * p.Wo + ow) * COUT;
      mbar_wait(&t_full[ab], (uint32_t)((j >> 1) & 1));
      __syncwarp();
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * COUT);
#pragma unroll 1
      for (int c0 = 0; c0 < COUT; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x16(taddr + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        tmem_ld_32x32b_x16(taddr + (uint32_t)(c0 + 16), *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
        tmem_ld_wait();
        if (c0 + 32 >= COUT) {                 // last read of this accumulator: hand the buffer back
          tc_fence_before();
          mbar_arrive(&t_empty[ab]);
        }
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          float r[8];
          const float4 s0 = *reinterpret_cast<const float4*>(sc2 + c0 + c), s1 = *reinterpret_cast<const float4*>(sc2 + c0 + c + 4);
          const float4 h0 = *reinterpret_cast<const float4*>(sh2 + c0 + c), h1 = *reinterpret_cast<const float4*>(sh2 + c0 + c + 4);
          const float scv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
          const float shv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) r[e] = relu_nan(__fadd_rn(__fmul_rn(__uint_as_float(v[c + e]), scv[e]), shv[e]));
          const float chk = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
          bad |= valid && (chk != chk) &&
                 ((r[0] != r[0]) | (r[1] != r[1]) | (r[2] != r[2]) | (r[3] != r[3]) | (r[4] != r[4]) | (r[5] != r[5]) |
                  (r[6] != r[6]) | (r[7] != r[7]));
          if (valid)
            *reinterpret_cast<uint4*>(dst + c0 + c) = make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]),
                                                                 pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7]));
        }
      }
    }
    if (bad && p.nan_flag) atomicOr(p.nan_flag, SSD3D_NAN_BACKBONE);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)C::TMEM_COLS);
  }
}

template <int S, int CIN, int COUT>
static int launch(const void* x, const void* w2, Params& p, cudaStream_t st) {
  using C = Cfg<S, CIN, COUT>;
  p.tiles_w = (p.Wo + C::TW - 1) / C::TW;
  p.tiles_h = (p.Ho + C::TH - 1) / C::TH;
  p.tiles_d = (p.Do + C::TD - 1) / C::TD;
  const long long tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_d * p.N;
  if (tiles > 0x3fffffffll) return SSD3D_ERR_UNSUPPORTED;
  p.tiles = (int)tiles;
  CUtensorMap tmX, tmW2;
  {
    const uint64_t dims[5] = {(uint64_t)CIN, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.D, (uint64_t)p.N};
    const uint64_t strides[4] = {(uint64_t)CIN * 2, (uint64_t)p.W * CIN * 2, (uint64_t)p.H * p.W * CIN * 2,
                                 (uint64_t)p.D * p.H * p.W * CIN * 2};
    const uint32_t box[5] = {(uint32_t)CB, (uint32_t)C::TWI, (uint32_t)C::THI, (uint32_t)C::TDI, 1u};
    if (make_tma_bf16(&tmX, x, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return SSD3D_ERR_TMA;
  }
  {
    const uint64_t dims[2] = {(uint64_t)CIN, (uint64_t)COUT};
    const uint64_t strides[1] = {(uint64_t)CIN * 2};
    const uint32_t box[2] = {(uint32_t)C::BK, (uint32_t)COUT};
    const CUtensorMapSwizzle sw = (C::BK == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    if (make_tma_bf16(&tmW2, w2, 2, dims, strides, box, sw)) return SSD3D_ERR_TMA;
  }
  cudaError_t e = cudaFuncSetAttribute(block_dwpw_kernel<S, CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)C::SMEM);
  if (e != cudaSuccess) return (int)e;
  const int n_sm = persistent_sms();
  const unsigned grid = (unsigned)(tiles < n_sm ? tiles : n_sm);
  SSD3D_LAUNCH_PDL((block_dwpw_kernel<S, CIN, COUT>), dim3(grid), dim3(THREADS), C::SMEM, st, tmX, tmW2, p);
  return SSD3D_OK;
}

}  // namespace dwpw
}  // namespace ssd3d

using namespace ssd3d;

extern "C" int ssd3d_block_fused_supported(int Cin, int Cout, int D, int H, int W, int stride) {
  const int Do = (D - 1) / stride + 1, Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  if (Wo < 8 || Ho < 4 || Do < 4) return 0;                      // small maps: latency bound either way
  if (stride == 2 && Cin == 32 && Cout == 64) return 1;
  if (stride == 2 && Cin == 64 && Cout == 128) return 1;
  if (stride == 1 && Cin == 128 && Cout == 128) return 1;
  return 0;
}

extern "C" int ssd3d_block_dwpw_bn_relu(const void* x, const void* w1, const float* scale1, const float* shift1,
                                        const void* w2, const float* scale2, const float* shift2, void* y, int N,
                                        int Cin, int Cout, int D, int H, int W, int stride, int* nan_flag,
                                        void* stream) {
  if (!x || !w1 || !scale1 || !shift1 || !w2 || !scale2 || !shift2 || !y || N <= 0 || D <= 0 || H <= 0 || W <= 0)
    return SSD3D_ERR_ARG;
  if (!ssd3d_block_fused_supported(Cin, Cout, D, H, W, stride)) return SSD3D_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || (reinterpret_cast<uintptr_t>(y) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(w2) & 15) != 0)
    return SSD3D_ERR_UNSUPPORTED;
  dwpw::Params p{};
  p.N = N; p.D = D; p.H = H; p.W = W;
  p.Do = (D - 1) / stride + 1; p.Ho = (H - 1) / stride + 1; p.Wo = (W - 1) / stride + 1;
  p.w1 = static_cast<const __nv_bfloat16*>(w1);
  p.scale1 = scale1; p.shift1 = shift1; p.scale2 = scale2; p.shift2 = shift2;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.nan_flag = nan_flag;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (stride == 2 && Cin == 32) return dwpw::launch<2, 32, 64>(x, w2, p, st);
  if (stride == 2 && Cin == 64) return dwpw::launch<2, 64, 128>(x, w2, p, st);
  return dwpw::launch<1, 128, 128>(x, w2, p, st);
}
