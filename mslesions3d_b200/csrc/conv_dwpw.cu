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
// Warp roles (512 threads; setmaxnreg moves registers from the service warps to the depthwise warps):
//   warps 0-7    depthwise, 192 registers: thread = 4 channels x {4 outputs along W (stride 1) | 2 outputs along H
//                (stride 2)}, its 27 x 4 weights register-resident as fp32 pairs, BN1 + ReLU, bf16 pack, 8-byte
//                store into the A tile; then arrive on a_full
//   warp 8       TMA producer: 5-D halo boxes (zero fill = padding), double buffered across chunks AND tiles
//                (stride 2: two boxes per chunk, de-interleaved along W, see Cfg); the pointwise weight tiles once
//   warp 9       TMEM allocation; UMMA issue: D[j & 1] = A . W2^T  (M 128, N Cout, K Cin), commit -> a_empty, t_full
//   warps 12-15  epilogue, 64 registers: tcgen05.ld 16 columns at a time, BN2 scale / shift, ReLU, NaN check, stores
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
// 16 warps: 0-7 depthwise (two warpgroups, 192 registers each after setmaxnreg), 8 TMA producer, 9 UMMA issuer,
// 10-11 idle, 12-15 epilogue (one per TMEM lane quarter); warps 8-15 drop to 64 registers
constexpr int DW_WARPS = 8;
constexpr int DW_THREADS = DW_WARPS * 32;
constexpr int THREADS = 512;
constexpr int W_PRODUCER = 8, W_MMA = 9, W_EPI0 = 12;

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 bf16x2_to_f32x2(uint32_t u) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(u << 16), "r"(u & 0xffff0000u));
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

// Tile geometry.  Stride 1: one 5-D halo box per 32-channel chunk, thread = 4 channels x 4 outputs along W (odd
// row pitch: the four 8-lane groups of a warp land in alternating halves of the 32 banks).  Stride 2: every input
// position an instruction touches has the same parity along each axis, so with 64-byte voxels all four lane groups
// would hit the same 16 banks (measured: 4 wavefronts per 8-byte load instead of 2).  The halo therefore arrives
// DE-INTERLEAVED along W -- two boxes with a TMA traversal stride of 2: E = halo columns 0, 2, .., 16 and
// O = columns 1, 3, .., 17 -- and the four lane groups of a warp take four CONSECUTIVE output columns, i.e.
// consecutive E / O entries: alternating bank halves, 2 wavefronts.  A thread then owns 4 channels x 2 outputs
// along H (5 input rows x {E[w], O[w], E[w+1]} per kd = 45 loads for 2 outputs, as before).
template <int S, int CIN, int COUT>
struct Cfg {
  static constexpr int TD = (S == 2) ? 4 : 2;
  static constexpr int TH = (S == 2) ? 4 : 8;
  static constexpr int TW = 8;
  static constexpr int TDI = S * (TD - 1) + 3;
  static constexpr int THI = S * (TH - 1) + 3;
  static constexpr int TWI0 = S * (TW - 1) + 3;
  static constexpr int TWI = (S == 2) ? 9 : ((TWI0 & 1) ? TWI0 : TWI0 + 1);   // columns per box (S=2: per parity)
  static constexpr int BOXES = (S == 2) ? 2 : 1;
  static constexpr int BOX_BYTES = TDI * THI * TWI * CB * 2;
  static constexpr int BOX_PITCH = (BOX_BYTES + 127) & ~127;        // TMA destinations: 128-byte aligned
  static constexpr int HALO_BYTES = BOXES * BOX_BYTES;
  static constexpr int HALO_PITCH = BOXES * BOX_PITCH;
  static constexpr int ITEMS = 128 / ((S == 2) ? 2 : 4) * 8;         // (4-channel vector) x (output group)
  static constexpr int IPT = ITEMS / DW_THREADS;                    // items per thread and chunk
  static constexpr int NCH = CIN / CB;
  static constexpr int BK = (CIN >= 64) ? 64 : 32;
  static constexpr int NKB = CIN / BK;
  static constexpr int A_BUFS = (S == 2 && CIN == 64) ? 1 : 2;      // (2, 64 -> 128) would not fit two
  static constexpr int A_BYTES = 128 * CIN * 2;
  static constexpr int B_BYTES = COUT * CIN * 2;
  static constexpr int W1_BYTES = CIN * 27 * 2;                     // bf16 [chunk][tap][32]
  static constexpr int TMEM_COLS = 2 * COUT;
  // layout (1024-aligned base): A[A_BUFS] | B (swizzled operands: 1024-aligned) | halo[2] | w1 | bn1 scale, shift |
  // bn2 scale, shift | barriers
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = OFF_A + A_BUFS * A_BYTES;
  static constexpr int OFF_HALO = OFF_B + B_BYTES;
  static constexpr int OFF_W1 = OFF_HALO + 2 * HALO_PITCH;
  static constexpr int OFF_BN = OFF_W1 + W1_BYTES;
  static constexpr int OFF_BAR = OFF_BN + (2 * CIN + 2 * COUT) * 4;
  static constexpr size_t SMEM = 1024 + (size_t)OFF_BAR + 16 * 8 + 16;
  static_assert(SMEM <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");
  static_assert(TD * TH * TW == 128, "one UMMA M tile per spatial tile");
  static_assert(ITEMS % DW_THREADS == 0, "whole items per thread");
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

// BN1 + ReLU + bf16 of one output voxel's 4 channels -> 8 bytes of the A tile
template <int BK>
__device__ __forceinline__ void store_a(uint8_t* A, int row, int c, const f32x2 (&acc)[2], const float4& sc,
                                        const float4& sh) {
  float a0, a1, a2, a3;
  unpack2(acc[0], a0, a1);
  unpack2(acc[1], a2, a3);
  a0 = relu_nan(__fadd_rn(__fmul_rn(a0, sc.x), sh.x));
  a1 = relu_nan(__fadd_rn(__fmul_rn(a1, sc.y), sh.y));
  a2 = relu_nan(__fadd_rn(__fmul_rn(a2, sc.z), sh.z));
  a3 = relu_nan(__fadd_rn(__fmul_rn(a3, sc.w), sh.w));
  *reinterpret_cast<uint2*>(A + a_offset<BK>(row, c)) = make_uint2(pack_bf16x2(a0, a1), pack_bf16x2(a2, a3));
}

template <int S, int CIN, int COUT>
__global__ void __launch_bounds__(THREADS, 1) block_dwpw_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                 const __grid_constant__ CUtensorMap tmX2,
                                                                 const __grid_constant__ CUtensorMap tmW2, const Params p) {
  using C = Cfg<S, CIN, COUT>;
  constexpr int TD = C::TD, TH = C::TH, TW = C::TW, BK = C::BK;
  extern __shared__ uint8_t dwpw_raw[];
  const uint32_t raw = smem_u32(dwpw_raw);
  uint8_t* smem = dwpw_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* sA = smem + C::OFF_A;
  uint8_t* sB = smem + C::OFF_B;
  uint8_t* halo = smem + C::OFF_HALO;
  __nv_bfloat16* sW1 = reinterpret_cast<__nv_bfloat16*>(smem + C::OFF_W1);   // [chunk][tap][32]
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
    if (S == 2) tma_prefetch_desc(&tmX2);
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
  if (warp == W_MMA) {
    tmem_alloc(tmem_slot, (uint32_t)C::TMEM_COLS);
    tmem_relinquish();
  }
  // constants (not produced by a preceding kernel): depthwise weights regrouped per chunk, folded BN vectors
  for (int i = tid; i < 27 * CIN; i += THREADS) {
    const int t = i / CIN, c = i % CIN;
    sW1[((c / CB) * 27 + t) * CB + (c % CB)] = p.w1[i];
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

  if (warp < DW_WARPS) {
    // ===================================== depthwise =====================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
    const int cv = tid & 7;
    f32x2 wreg[27][2];                 // this thread's 4 channels x 27 taps as fp32 pairs, per chunk
    auto load_weights = [&](int ch) {
      const __nv_bfloat16* wch = sW1 + (size_t)ch * 27 * CB + cv * 4;
#pragma unroll
      for (int t = 0; t < 27; ++t) {
        const uint2 u = *reinterpret_cast<const uint2*>(wch + t * CB);
        wreg[t][0] = bf16x2_to_f32x2(u.x);
        wreg[t][1] = bf16x2_to_f32x2(u.y);
      }
    };
    if (C::NCH == 1) load_weights(0);
    int j = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++j) {
      const int ab = (C::A_BUFS == 2) ? (j & 1) : 0;
      if (j >= C::A_BUFS)              // the GEMM of tile j - A_BUFS has read A[ab]
        mbar_wait(&a_empty[ab], (uint32_t)((C::A_BUFS == 2 ? ((j >> 1) - 1) : (j - 1)) & 1));
      int ow0, oh0, od0, n;
      decode(tile, ow0, oh0, od0, n);
      uint8_t* A = sA + (size_t)ab * C::A_BYTES;
#pragma unroll 1
      for (int ch = 0; ch < C::NCH; ++ch) {
        const int hs = j * C::NCH + ch, hb = hs & 1;
        if (C::NCH > 1) load_weights(ch);
        const float4 sc = *reinterpret_cast<const float4*>(sBn + ch * CB + cv * 4);
        const float4 sh = *reinterpret_cast<const float4*>(sBn + CIN + ch * CB + cv * 4);
        mbar_wait(&halo_full[hb], (uint32_t)((hs >> 1) & 1));
        const uint8_t* in = halo + (size_t)hb * C::HALO_PITCH;
#pragma unroll 1
        for (int k = 0; k < C::IPT; ++k) {
          const int item = tid + k * DW_THREADS;
          if (S == 2) {
            // thread = 4 channels x outputs (d, 2hp, w) and (d, 2hp + 1, w); lane groups = consecutive w
            const int w = (item >> 3) & 7, hp = (item >> 6) & 1, d = item >> 7;
            if (od0 + d >= p.Do || oh0 + 2 * hp >= p.Ho || ow0 + w >= p.Wo) continue;      // rows nobody stores
            f32x2 acc0[2] = {0ull, 0ull}, acc1[2] = {0ull, 0ull};
            const uint8_t* base = in + ((size_t)(((2 * d) * C::THI + 4 * hp) * C::TWI + w) * CB + cv * 4) * 2;
#pragma unroll
            for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
              for (int r = 0; r < 5; ++r) {
                const uint8_t* row = base + (size_t)((kd * C::THI + r) * C::TWI) * CB * 2;
                const uint2 u0 = *reinterpret_cast<const uint2*>(row);                      // E[w]   : kw = 0
                const uint2 u1 = *reinterpret_cast<const uint2*>(row + C::BOX_PITCH);       // O[w]   : kw = 1
                const uint2 u2 = *reinterpret_cast<const uint2*>(row + CB * 2);             // E[w+1] : kw = 2
                const f32x2 x0[2] = {bf16x2_to_f32x2(u0.x), bf16x2_to_f32x2(u0.y)};
                const f32x2 x1[2] = {bf16x2_to_f32x2(u1.x), bf16x2_to_f32x2(u1.y)};
                const f32x2 x2[2] = {bf16x2_to_f32x2(u2.x), bf16x2_to_f32x2(u2.y)};
                if (r <= 2) {          // output 0: kh = r
                  const int t = (kd * 3 + r) * 3;
                  ffma2(acc0[0], x0[0], wreg[t][0]);     ffma2(acc0[1], x0[1], wreg[t][1]);
                  ffma2(acc0[0], x1[0], wreg[t + 1][0]); ffma2(acc0[1], x1[1], wreg[t + 1][1]);
                  ffma2(acc0[0], x2[0], wreg[t + 2][0]); ffma2(acc0[1], x2[1], wreg[t + 2][1]);
                }
                if (r >= 2) {          // output 1: kh = r - 2
                  const int t = (kd * 3 + (r - 2)) * 3;
                  ffma2(acc1[0], x0[0], wreg[t][0]);     ffma2(acc1[1], x0[1], wreg[t][1]);
                  ffma2(acc1[0], x1[0], wreg[t + 1][0]); ffma2(acc1[1], x1[1], wreg[t + 1][1]);
                  ffma2(acc1[0], x2[0], wreg[t + 2][0]); ffma2(acc1[1], x2[1], wreg[t + 2][1]);
                }
              }
            }
            const int row0 = (d * TH + 2 * hp) * TW + w;
            store_a<BK>(A, row0, ch * CB + cv * 4, acc0, sc, sh);
            store_a<BK>(A, row0 + TW, ch * CB + cv * 4, acc1, sc, sh);
          } else {
            // thread = 4 channels x 4 outputs along W; lane groups = consecutive h (odd row pitch)
            constexpr int WT = 4, NI = WT + 2;
            int r = item >> 3;
            const int h = r % TH; r /= TH;
            const int wq = r % (TW / WT);
            const int d = r / (TW / WT);
            if (od0 + d >= p.Do || oh0 + h >= p.Ho || ow0 + wq * WT >= p.Wo) continue;
            f32x2 acc[WT][2];
#pragma unroll
            for (int i = 0; i < WT; ++i) { acc[i][0] = 0ull; acc[i][1] = 0ull; }
            const uint8_t* base = in + ((size_t)((d * C::THI + h) * C::TWI + wq * WT) * CB + cv * 4) * 2;
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
                  const int t = (kd * 3 + kh) * 3 + kw;
#pragma unroll
                  for (int ow = 0; ow < WT; ++ow) {
                    ffma2(acc[ow][0], x[ow + kw][0], wreg[t][0]);
                    ffma2(acc[ow][1], x[ow + kw][1], wreg[t][1]);
                  }
                }
              }
            }
            const int row0 = (d * TH + h) * TW + wq * WT;
#pragma unroll
            for (int ow = 0; ow < WT; ++ow) store_a<BK>(A, row0 + ow, ch * CB + cv * 4, acc[ow], sc, sh);
          }
        }
        mbar_arrive(&halo_empty[hb]);          // this thread is done reading halo[hb]
      }
      fence_proxy_async_smem();                // generic-proxy writes of A -> visible to the tensor core (async proxy)
      mbar_arrive(&a_full[ab]);
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == W_PRODUCER) {
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
            uint8_t* dst = halo + (size_t)hb * C::HALO_PITCH;
            mbar_arrive_expect_tx(&halo_full[hb], (uint32_t)C::HALO_BYTES);
            tma_load_5d(dst, &tmX, &halo_full[hb], ch * CB, ow0 * S - 1, oh0 * S - 1, od0 * S - 1, n);
            if (S == 2)                 // the odd halo columns: same box, one tensor element further along W
              tma_load_5d(dst + C::BOX_PITCH, &tmX2, &halo_full[hb], ch * CB, ow0 * S, oh0 * S - 1, od0 * S - 1, n);
          }
        }
      }
      __syncwarp();
    } else if (warp == W_MMA) {
      // ===================================== UMMA issuer =====================================
      if (lane == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, COUT);
        mbar_wait(b_full, 0u);
        int j = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++j) {
          const int ab = (C::A_BUFS == 2) ? (j & 1) : 0, tb = j & 1;
          mbar_wait(&a_full[ab], (uint32_t)((C::A_BUFS == 2 ? (j >> 1) : j) & 1));
          if (j >= 2) mbar_wait(&t_empty[tb], (uint32_t)(((j >> 1) - 1) & 1));
          tc_fence_after();
          const uint32_t acc = tmem_base + (uint32_t)(tb * COUT);
          const uint32_t a_addr = smem_u32(sA + (size_t)ab * C::A_BYTES), b_addr = smem_u32(sB);
#pragma unroll
          for (int kb = 0; kb < C::NKB; ++kb) {
            const uint64_t da = smem_desc<BK>(a_addr + kb * 128 * BK * 2), db = smem_desc<BK>(b_addr + kb * COUT * BK * 2);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_ss(acc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&a_empty[ab]);
          umma_commit(&t_full[tb]);
        }
      }
      __syncwarp();
    } else if (warp >= W_EPI0) {
      // ===================================== epilogue =====================================
      const int q = warp & 3;                    // TMEM lane quarter this warp may read
      const int row = q * 32 + lane;
      const int rw = row % TW, rh = (row / TW) % TH, rd = row / (TW * TH);
      const float* sc2 = sBn + 2 * CIN;
      const float* sh2 = sc2 + COUT;
      bool bad = false;
      int j = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++j) {
        const int tb = j & 1;
        int ow0, oh0, od0, n;
        decode(tile, ow0, oh0, od0, n);
        const int od = od0 + rd, oh = oh0 + rh, ow = ow0 + rw;
        const bool valid = od < p.Do && oh < p.Ho && ow < p.Wo;
        __nv_bfloat16* dst = p.y + ((((long long)n * p.Do + od) * p.Ho + oh) * p.Wo + ow) * COUT;
        mbar_wait(&t_full[tb], (uint32_t)((j >> 1) & 1));
        __syncwarp();
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tb * COUT);
#pragma unroll 1
        for (int c0 = 0; c0 < COUT; c0 += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(taddr + (uint32_t)c0, v);
          tmem_ld_wait();
          if (c0 + 16 >= COUT) {                 // last read of this accumulator: hand the buffer back
            tc_fence_before();
            mbar_arrive(&t_empty[tb]);
          }
#pragma unroll
          for (int c = 0; c < 16; c += 8) {
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
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
  CUtensorMap tmX, tmX2, tmW2;
  {
    const uint64_t dims[5] = {(uint64_t)CIN, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.D, (uint64_t)p.N};
    const uint64_t strides[4] = {(uint64_t)CIN * 2, (uint64_t)p.W * CIN * 2, (uint64_t)p.H * p.W * CIN * 2,
                                 (uint64_t)p.D * p.H * p.W * CIN * 2};
    if (S == 2) {
      // de-interleaved along W: the box spans 2*TWI - 1 = 17 columns and every second one is loaded (9 per box)
      const uint32_t box[5] = {(uint32_t)CB, (uint32_t)(2 * C::TWI - 1), (uint32_t)C::THI, (uint32_t)C::TDI, 1u};
      const uint32_t estr[5] = {1u, 2u, 1u, 1u, 1u};
      if (make_tma_bf16(&tmX, x, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, estr)) return SSD3D_ERR_TMA;
      tmX2 = tmX;
    } else {
      const uint32_t box[5] = {(uint32_t)CB, (uint32_t)C::TWI, (uint32_t)C::THI, (uint32_t)C::TDI, 1u};
      if (make_tma_bf16(&tmX, x, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return SSD3D_ERR_TMA;
      tmX2 = tmX;
    }
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
  SSD3D_LAUNCH_PDL((block_dwpw_kernel<S, CIN, COUT>), dim3(grid), dim3(THREADS), C::SMEM, st, tmX, tmX2, tmW2, p);
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
