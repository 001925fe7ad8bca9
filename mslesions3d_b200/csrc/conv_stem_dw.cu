// Stem (dense 3x3x3 Cin -> 32, stride (sd,2,2), BN, ReLU; mobilenet.py:26-31) FUSED with the depthwise conv of the
// first Block (3x3x3, 32 channels, stride 2, BN, ReLU; mobilenet.py:38,44): the stem's activation -- 134 MB
// written and 134 MB read back per batch of eight 2ch 128^3 volumes, half of the whole step's HBM traffic, and
// HBM writes top out at 3.9 TB/s on this part (scripts/probe_hbm.py) -- never leaves the SM.
//
// Stem part = the banded-B tcgen05 GEMM of conv_stem_tz.cu (A operand = raw TMA'd input rows, B = Toeplitz
// weights, M = 16 row slots of 32 stem voxels, N = 4 voxels x 32 channels).  Geometry for stem rows of 64 voxels
// (W = 128):
//   * a PLANE JOB is one stem plane P restricted to 16 stem rows (15 used: 2*7+1) = two M-tiles (left / right
//     half of the row) = 36 UMMAs into two of the four 128-column TMEM accumulators;
//   * the epilogue warps apply BN+ReLU and write the plane to ONE shared-memory buffer, de-interleaved along W
//     (odd columns O[0..32] incl. the zero padding column -1, even columns E[0..31]) and chunk-swizzled
//     (16-byte unit ^= line & 7), so that both the epilogue's 16-byte stores (lanes = consecutive 4-voxel
//     groups) and the depthwise loads (lanes = 8 channel quads x 4 consecutive outputs) are conflict free;
//     rows / planes outside the stem map are written as zeros (the depthwise conv's padding);
//   * depthwise part: thread = (4 channels, output column w), owns the 7 output rows of the tile for TWO output
//     planes in registers (accA = plane d, accB = plane d+1): stem plane 2d adds its kd=1 taps to accA, plane
//     2d+1 its kd=2 taps to accA and its kd=0 taps to accB, then plane d is finished (BN+ReLU, 8-byte stores,
//     a warp writes 256 contiguous bytes) and accB becomes accA.  Taps are applied in (kd, kh, kw) order with
//     fma.rn.f32x2: bit-identical to the stand-alone depthwise kernels on the same stem values.
//   * a CTA marches along D through a contiguous run of (image, 7-row tile, plane d) steps (8.6 per SM for the
//     benchmark); a run that starts at d > 0 first computes stem plane 2d-1 (its kd=0 share).
//   * warp roles: warp 0 = TMA producer (ring of nine (half, kd) units = 1.5 plane jobs; a unit holds the 17 even and
//     16 odd input rows of the tile once: tap kh = 2 reads the even rows one slot further), warp 1 = UMMA issuer,
//     warps 2..9 = epilogue + depthwise.  The UMMAs of job j+1 run while the CUDA cores finish job j.
// Requirements: bf16 volumes, Cin <= 2, W = 128 (stem rows of 64 voxels), even stem H and D.
#include "common.cuh"
#include "tma_host.h"

namespace ssd3d {

struct StemDwParams {
  int N, D, H, W, sd;
  int Ds, Hs;                  // stem map (Ws = 64)
  int Dd, Hd;                  // depthwise output map (Wd = 32)
  int HT;                      // 7-row tiles along Hd
  int steps;                   // N * HT * Dd
  int kpad;                    // row pitch of wt
  const __nv_bfloat16* wt;     // stem (32, kpad) bf16, k = ((ci*3+kd)*3+kh)*3+kw
  const float* scale0;         // stem BN
  const float* shift0;
  const __nv_bfloat16* wd;     // depthwise (27, 32) bf16
  const float* scale1;         // depthwise BN
  const float* shift1;
  __nv_bfloat16* y;            // (N, Dd, Hd, 32, 32)
  float one;                   // 1.0f (a run-time value on purpose: see fadd2)
};

namespace sdw {

constexpr int SLOT_BYTES = 144;
// One (half c, kd) unit of the A ring holds, per input channel, the 33 input rows the 16 stem rows of the tile read:
// the 17 even ones (taps kh = 0 and, one slot further, kh = 2) and the 16 odd ones (kh = 1).
constexpr int EVEN_BYTES = 2560;              // 17 x 144 = 2448, padded to a multiple of 128 (TMA destination)
constexpr int ODD_BYTES = 16 * SLOT_BYTES;    // 2304
constexpr int CI_BYTES = EVEN_BYTES + ODD_BYTES;   // 4864
constexpr int CI_TX = 17 * SLOT_BYTES + ODD_BYTES; // bytes TMA really writes per input channel
constexpr int B_BYTES = 4096;
constexpr int NU = 3;                         // ring of half-plane units (3 kd each): 1.5 plane jobs
constexpr int CWARPS = 16;                     // epilogue + depthwise warps
constexpr int THREADS = 64 + CWARPS * 32;     // + TMA producer warp + UMMA issuer warp
constexpr int TH = 7;                         // depthwise rows per tile
constexpr int O_BYTES = 17 * 128;             // 33 odd-column entries of 64 B
constexpr int ROW_BYTES = 33 * 128;           // O then E (32 entries)
constexpr int PLANE_BYTES = 15 * ROW_BYTES;   // 63360: the 15 stem rows a 7-row tile reads

__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ uint64_t desc_k_nosw(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ float relu_nan1(float v) {   // max.NaN: torch.relu semantics
  float r;
  asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(v));
  return r;
}

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack_f32x2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi)));
  return r;
}
__device__ __forceinline__ f32x2 bf16x2_to_f32x2(uint32_t u) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(u << 16), "r"(u & 0xffff0000u));
  return r;
}
__device__ __forceinline__ void ffma2(f32x2& acc, f32x2 a, f32x2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ void unpack_f32x2(f32x2 v, float& lo, float& hi) {
  uint32_t a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v));
  lo = __uint_as_float(a);
  hi = __uint_as_float(b);
}

// Separately rounded packed multiply and add.  ptxas turns mul.rn.f32x2 + add.rn.f32x2 (and fma(a,b,-0) followed
// by fma(t,1,c), with or without -fmad=false) into ONE FFMA2 -- a contraction the .rn modifiers should forbid, seen
// in the SASS and as 1-ulp flips against the stand-alone kernels.  The add is therefore an fma with a multiplier
// of 1.0 that ptxas cannot see through (read from shared memory at run time): t * 1 + b rounds exactly like add.rn.
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b, f32x2 one) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(one), "l"(b));
  return r;
}
// (separately rounded) a * s + b on two channels, rounded to bf16, then ReLU on the packed pair: rounding is
// monotonic and keeps the sign, so relu(round(v)) == round(relu(v)); max.NaN propagates NaN like torch.relu
__device__ __forceinline__ uint32_t bn_relu_pack(uint32_t a_lo, uint32_t a_hi, f32x2 s, f32x2 b, f32x2 one) {
  f32x2 a;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(a_lo), "r"(a_hi));
  const f32x2 v = fadd2(fmul2(a, s), b, one);
  uint32_t lo, hi, r;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  asm("max.NaN.bf16x2 %0, %0, %1;" : "+r"(r) : "r"(0u));
  return r;
}

// byte offset of 16-byte unit c4 (0..3) of entry e inside a de-interleaved column array (O or E) of the plane
__device__ __forceinline__ uint32_t entry_off(int e, int c4) {
  const int line = e >> 1;
  return (uint32_t)(line * 128 + (((((e & 1) << 2) | c4) ^ (line & 7)) << 4));
}

// The plane jobs of a CTA's run of steps [L0, L1), identical in every warp role.
struct Jobs {
  int L, L0, L1, Dd, stage;
  int col, d, P, role;      // role 0: plane 2d-1 (kd=0 share only), 1: plane 2d (kd=1), 2: plane 2d+1 (kd=2, then finish)
  bool cont;                // role 2: the run goes on with plane d+1 of the same column (kd=0 share into accB)
  __device__ Jobs(int l0, int l1, int dd) : L(l0), L0(l0), L1(l1), Dd(dd), stage(0), col(0), d(0), P(0), role(0), cont(false) {}
  __device__ bool next() {
    if (L >= L1) return false;
    col = L / Dd;
    d = L - col * Dd;
    if (stage == 0) {
      stage = 1;
      if (L == L0 && d > 0) { P = 2 * d - 1; role = 0; return true; }
    }
    if (stage == 1) { stage = 2; P = 2 * d; role = 1; return true; }
    P = 2 * d + 1; role = 2;
    cont = (L + 1 < L1) && (d + 1 < Dd);
    stage = 0;
    ++L;
    return true;
  }
};

// The taps of one stem plane (15 rows in shared memory) for this thread's 2 channels and output column: kd share
// `wa` into accA and, when TOB, the kd = 0 share `wb` into accB.  Stem row r is tap kh = r - 2i of output row i:
// even r -> (i = r/2, kh 0) and (i = r/2 - 1, kh 2); odd r -> kh 1.  Output row i sees its kh = 0, 1, 2 rows in
// ascending r: (kd, kh, kw) order as in the stand-alone depthwise kernels.  The three words of row r+1 are loaded
// before the FMAs of row r (two compute warps per scheduler do not hide a shared-memory round trip by themselves).
template <bool TOB>
__device__ __forceinline__ void dw_plane(f32x2 (&accA)[TH], f32x2 (&accB)[TH], const float* wA, const float* wB,
                                         uint32_t rdO0, uint32_t rdE, uint32_t rdO1) {
  f32x2 wa[9], wb[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    wa[t] = *reinterpret_cast<const f32x2*>(wA + t * 32);
    if (TOB) wb[t] = *reinterpret_cast<const f32x2*>(wB + t * 32);
  }
  uint32_t u[3], un[3];
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u[0]) : "r"(rdO0));
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u[1]) : "r"(rdE));
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u[2]) : "r"(rdO1));
#pragma unroll
  for (int r = 0; r < 15; ++r) {
    if (r + 1 < 15) {
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(un[0]) : "r"(rdO0 + (uint32_t)((r + 1) * ROW_BYTES)));
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(un[1]) : "r"(rdE + (uint32_t)((r + 1) * ROW_BYTES)));
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(un[2]) : "r"(rdO1 + (uint32_t)((r + 1) * ROW_BYTES)));
    }
    const f32x2 x[3] = {bf16x2_to_f32x2(u[0]), bf16x2_to_f32x2(u[1]), bf16x2_to_f32x2(u[2])};
#pragma unroll
    for (int kh = 2; kh >= 0; --kh) {
      if (((r - kh) & 1) != 0 || r - kh < 0 || (r - kh) / 2 >= TH) continue;
      const int i = (r - kh) / 2;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        ffma2(accA[i], x[kw], wa[kh * 3 + kw]);
        if (TOB) ffma2(accB[i], x[kw], wb[kh * 3 + kw]);
      }
    }
    u[0] = un[0]; u[1] = un[1]; u[2] = un[2];
  }
}

}  // namespace sdw

template <int CIN>
__global__ void __launch_bounds__(sdw::THREADS, 1) stem_dw_kernel(const __grid_constant__ CUtensorMap tmE,
                                                                 const __grid_constant__ CUtensorMap tmO,
                                                                 const StemDwParams p) {
  using namespace sdw;
  constexpr int NT = 9 * CIN;
  constexpr int KD_BYTES = CIN * CI_BYTES;           // one kd of a unit: [ci][17 even rows | 16 odd rows][72]
  constexpr int UNIT_BYTES = 3 * KD_BYTES;           // one half (c) of a plane job
  constexpr uint32_t TMEM_COLS = 512;                // four 128-column accumulators

  extern __shared__ uint8_t sdw_raw[];
  const uint32_t raw = smem_u32(sdw_raw);
  uint8_t* smem = sdw_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* sP = smem;                                            // stem plane buffer
  uint8_t* sB = sP + PLANE_BYTES;                                // NT x 4096
  uint8_t* sA = sB + NT * B_BYTES;                               // NU x UNIT_BYTES
  float* sWd = reinterpret_cast<float*>(sA + NU * UNIT_BYTES);   // depthwise weights fp32 [27][32]
  float* sSc0 = sWd + 27 * 32;                                   // stem BN scale[32], shift[32]
  float* sSh0 = sSc0 + 32;
  float* sOne = sSh0 + 32;                                       // {1, 1}: see fadd2
  uint64_t* full = reinterpret_cast<uint64_t*>(sOne + 2);
  uint64_t* empty = full + NU;
  uint64_t* acc_full = empty + NU;                               // [4]
  uint64_t* acc_empty = acc_full + 4;                            // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 4);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int L0 = (int)(((long long)blockIdx.x * p.steps) / gridDim.x);
  const int L1 = (int)(((long long)(blockIdx.x + 1) * p.steps) / gridDim.x);

  if (tid == 0) {
    tma_prefetch_desc(&tmE);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < NU; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], CWARPS); }
    fence_barrier_init();
  }
  __syncthreads();              // the barriers exist

  if (warp == 0) {
    // ===================== TMA producer =====================
    // (does not take part in the weight set-up below: the first units are in flight while the others build B)
    pdl_wait();                 // the input belongs to earlier work
    const uint32_t sA_u = smem_u32(sA);
    int u = 0;
    uint32_t ph = 1;                                   // parity to wait for on empty[u]: the first lap passes
    Jobs jobs(L0, L1, p.Dd);
    while (jobs.next()) {
      const int n = jobs.col / p.HT, ht = jobs.col - n * p.HT;
      const int ch = 4 * (ht * TH) - 3;                // input row of stem row 2*h0 - 1, tap kh = 0
      const int cd = p.sd * jobs.P - 1;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {                    // unit = half c of the plane: 3 kd x CIN x (even, odd) boxes
        mbar_wait(&empty[u], ph);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[u], (uint32_t)(3 * CIN * CI_TX));
          const uint32_t dst = sA_u + (uint32_t)(u * UNIT_BYTES);
#pragma unroll
          for (int kd = 0; kd < 3; ++kd)
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
              tma_load_4d(dst + kd * KD_BYTES + ci * CI_BYTES, &tmE, &full[u], 64 * c - 8, ch, cd + kd, n * CIN + ci);
              tma_load_4d(dst + kd * KD_BYTES + ci * CI_BYTES + EVEN_BYTES, &tmO, &full[u], 64 * c - 8, ch + 1, cd + kd,
                          n * CIN + ci);
            }
        }
        __syncwarp();
        if (++u == NU) { u = 0; ph ^= 1u; }
      }
    }
  } else {
    // ---- weight set-up by warps 1..: banded B operand, depthwise weights, stem BN; the plane buffer starts as
    //      zeros (column -1 stays zero).  Weights are constants of the (inference) step: this runs BEFORE the
    //      grid-dependency wait, i.e. under the tail of the previous kernel.
    constexpr int SETUP = THREADS - 32;
    const int st = tid - 32;
    if (warp == 1) {
      __syncwarp();
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
    for (int i = st; i < (PLANE_BYTES + NT * B_BYTES) / 16; i += SETUP) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = st; i < 27 * 32; i += SETUP) sWd[i] = __bfloat162float(p.wd[i]);
    if (st < 32) { sSc0[st] = __ldg(p.scale0 + st); sSh0[st] = __ldg(p.shift0 + st); }
    if (st < 2) sOne[st] = p.one;
    __nv_bfloat16 wv[(NT * 3 * 32 + SETUP - 1) / SETUP];
#pragma unroll
    for (int it = 0; it < (NT * 3 * 32 + SETUP - 1) / SETUP; ++it) {      // all weight loads in flight together
      const int i = st + it * SETUP;
      wv[it] = (i < NT * 3 * 32) ? p.wt[(i & 31) * p.kpad + (i >> 5)] : __float2bfloat16(0.f);
    }
    asm volatile("bar.sync 2, %0;" ::"r"(SETUP) : "memory");               // zero fill done
#pragma unroll
    for (int it = 0; it < (NT * 3 * 32 + SETUP - 1) / SETUP; ++it) {
      const int i = st + it * SETUP;
      if (i < NT * 3 * 32) {
        const int co = i & 31, tk = i >> 5;
        const int t = tk / 3, kw = tk - 3 * t;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = j * 32 + co, k = 7 + 2 * j + kw;
          *reinterpret_cast<__nv_bfloat16*>(sB + t * B_BYTES + (n >> 3) * 256 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) = wv[it];
        }
      }
    }
    fence_proxy_async_smem();   // generic-proxy writes of B -> visible to the UMMA (async proxy) reads
    tc_fence_before();
    asm volatile("bar.sync 2, %0;" ::"r"(SETUP) : "memory");
    tc_fence_after();
    pdl_wait();                 // the output buffer's previous readers
    pdl_launch_dependents();
  }
  const uint32_t tmem_base = *tmem_slot;   // (a plain load: a select on the warp index here makes ptxas treat every
                                            // UMMA operand as non-uniform -> ELECT loops and five R2UR per UMMA)

  if (warp == 1) {
    // ===================== UMMA issuer =====================
    // One lane runs the whole loop, outer loops rolled, the 3 x CIN UMMAs of a unit unrolled with immediate
    // descriptor offsets: ptxas then keeps the descriptors in UNIFORM registers (UIADD3.64 + UTCHMMA, as in
    // gemm_pw.cu).  With the loops unrolled it precomputes all 36 descriptor pairs in vector registers and pays four
    // R2UR per UMMA: ~110 cycles of issue per UMMA (measured: independent of N), twice the tensor-pipe time.
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint32_t sA_u = smem_u32(sA);
      const uint64_t db = desc_k_nosw(smem_u32(sB), 128u, 256u);
      int j = 0, u = 0;
      uint32_t ph = 0;
      Jobs jobs(L0, L1, p.Dd);
      while (jobs.next()) {
        const uint32_t apar = (uint32_t)((j >> 1) & 1);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const int buf = (j & 1) * 2 + c;
          mbar_wait(&acc_empty[buf], apar ^ 1u);
          mbar_wait(&full[u], ph);
          tc_fence_after();
          const uint32_t dcol = tmem_base + (uint32_t)(buf * 128);
          const uint64_t da = desc_k_nosw(sA_u + (uint32_t)(u * UNIT_BYTES), 16u, (uint32_t)SLOT_BYTES);
          // the 18 UMMAs of the half plane back to back: the issue queue is shallow (issue time == execution time in
          // scripts/ubench/umma_rate.cu), every cycle this thread spends elsewhere is a cycle the tensor pipe idles
#pragma unroll
          for (int kd = 0; kd < 3; ++kd)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              // kh = 0: even rows from slot 0, kh = 1: odd rows, kh = 2: even rows from slot 1
              const int off = kd * KD_BYTES + (kh == 0 ? 0 : (kh == 1 ? EVEN_BYTES : SLOT_BYTES));
#pragma unroll
              for (int ci = 0; ci < CIN; ++ci)
                umma_bf16_ss(dcol, da + (uint64_t)((off + ci * CI_BYTES) >> 4),
                             db + (uint64_t)(((ci * 9 + kd * 3 + kh) * B_BYTES) >> 4), idesc, (kd | kh | ci) != 0 ? 1u : 0u);
            }
          umma_commit(&empty[u]);
          umma_commit(&acc_full[buf]);
          if (++u == NU) { u = 0; ph ^= 1u; }
        }
        ++j;
      }
    }
    __syncwarp();
  } else if (warp >= 2) {
    // ===================== 16 warps: epilogue (TMEM -> BN/ReLU -> plane buffer) + depthwise =====================
    const int ct = tid - 64;                     // 0..511
    // epilogue role: lane quarter q, voxel jv of the 4-voxel group (32 accumulator columns)
    const int q = warp & 3;
    const int jv = (warp - 2) >> 2;
    const int row = (q * 32 + lane) >> 3;        // stem row slot 0..15
    const int g = lane & 7;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(jv * 32);
    const uint32_t sP_u = smem_u32(sP);
    // voxel 32c + 4g + jv: even column -> E[16c + 2g + jv/2], odd column -> O[16c + 2g + (jv+1)/2]
    const uint32_t wr_base = sP_u + (uint32_t)(row * ROW_BYTES) + ((jv & 1) ? 0u : (uint32_t)O_BYTES);
    const int wr_e = 2 * g + ((jv + 1) >> 1);
    // depthwise role: 2 channels x output column w
    const int cp = ct & 15;                      // channels 2*cp, 2*cp + 1
    const int w = ct >> 4;                       // output column 0..31
    const float2 sc1 = __ldg(reinterpret_cast<const float2*>(p.scale1 + 2 * cp));
    const float2 sh1 = __ldg(reinterpret_cast<const float2*>(p.shift1 + 2 * cp));
    const uint32_t sub = (uint32_t)((cp & 3) * 4);
    const uint32_t rdE = sP_u + (uint32_t)O_BYTES + entry_off(w, cp >> 2) + sub;
    const uint32_t rdO0 = sP_u + entry_off(w, cp >> 2) + sub;
    const uint32_t rdO1 = sP_u + entry_off(w + 1, cp >> 2) + sub;
    const f32x2 one2 = *reinterpret_cast<const f32x2*>(sOne);
    f32x2 accA[TH], accB[TH];
#pragma unroll
    for (int i = 0; i < TH; ++i) { accA[i] = 0ull; accB[i] = 0ull; }

    int j = 0;
    Jobs jobs(L0, L1, p.Dd);
    while (jobs.next()) {
      const int n = jobs.col / p.HT, ht = jobs.col - n * p.HT;
      const int h0 = ht * TH;                    // first depthwise row of the tile
      const uint32_t apar = (uint32_t)((j >> 1) & 1);
      const int hs = 2 * h0 - 1 + row;           // stem row of this thread's slot
      const bool row_ok = (hs >= 0) && (hs < p.Hs);
      // ---------------- phase A: both halves of the plane ----------------
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int buf = (j & 1) * 2 + c;
        mbar_wait(&acc_full[buf], apar);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_x32(lane_addr + (uint32_t)(buf * 128), v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        uint32_t out[16];
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          const ulonglong2 sc = *reinterpret_cast<const ulonglong2*>(sSc0 + k);
          const ulonglong2 sh = *reinterpret_cast<const ulonglong2*>(sSh0 + k);
          out[(k >> 1)] = bn_relu_pack(v[k], v[k + 1], sc.x, sh.x, one2);
          out[(k >> 1) + 1] = bn_relu_pack(v[k + 2], v[k + 3], sc.y, sh.y, one2);
        }
        if (!row_ok) {
#pragma unroll
          for (int k = 0; k < 16; ++k) out[k] = 0u;
        }
        // the depthwise pass of the previous job must be done with the plane buffer
        if (c == 0) asm volatile("bar.sync 1, %0;" ::"n"(CWARPS * 32) : "memory");
        if (row < 15) {
          const int e = 16 * c + wr_e;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(wr_base + entry_off(e, c4)), "r"(out[4 * c4]),
                         "r"(out[4 * c4 + 1]), "r"(out[4 * c4 + 2]), "r"(out[4 * c4 + 3])
                         : "memory");
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(CWARPS * 32) : "memory");   // the plane is complete
      // ---------------- phase B: this plane's taps ----------------
      {
        const int role = jobs.role;
        // kd share of accA: role 0 -> kd 0, role 1 -> kd 1, role 2 -> kd 2; accB (the next output plane) gets kd 0
        const float* wA = sWd + (role * 9) * 32 + 2 * cp;
        const float* wB = sWd + 2 * cp;
        if ((role == 2) && jobs.cont) dw_plane<true>(accA, accB, wA, wB, rdO0, rdE, rdO1);
        else dw_plane<false>(accA, accB, wA, wB, rdO0, rdE, rdO1);
        if (role == 2) {
          // depthwise plane d of this tile is complete
          __nv_bfloat16* o = p.y + ((((long long)n * p.Dd + jobs.d) * p.Hd + h0) * 32 + w) * 32 + 2 * cp;
#pragma unroll
          for (int i = 0; i < TH; ++i) {
            if (h0 + i < p.Hd) {
              float a0, a1;
              unpack_f32x2(accA[i], a0, a1);
              a0 = relu_nan1(__fadd_rn(__fmul_rn(a0, sc1.x), sh1.x));
              a1 = relu_nan1(__fadd_rn(__fmul_rn(a1, sc1.y), sh1.y));
              *reinterpret_cast<uint32_t*>(o + (long long)i * 32 * 32) = pack_bf16x2(a0, a1);
            }
            accA[i] = accB[i];
            accB[i] = 0ull;
          }
        }
      }
      ++j;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int CIN>
static int launch_stem_dw(const void* x, const StemDwParams& p, cudaStream_t st) {
  using namespace sdw;
  CUtensorMap tmE, tmO;
  {
    const uint64_t dims[4] = {(uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.D, (uint64_t)p.N * CIN};
    const uint64_t strides[3] = {(uint64_t)p.W * 2, (uint64_t)p.W * p.H * 2, (uint64_t)p.W * p.H * p.D * 2};
    const uint32_t estr[4] = {1u, 2u, 1u, 1u};
    const uint32_t boxE[4] = {72u, 33u, 1u, 1u};      // 17 even input rows of one channel
    const uint32_t boxO[4] = {72u, 31u, 1u, 1u};      // 16 odd ones
    if (make_tma_bf16(&tmE, x, 4, dims, strides, boxE, CU_TENSOR_MAP_SWIZZLE_NONE, estr)) return SSD3D_ERR_TMA;
    if (make_tma_bf16(&tmO, x, 4, dims, strides, boxO, CU_TENSOR_MAP_SWIZZLE_NONE, estr)) return SSD3D_ERR_TMA;
  }
  const size_t smem = 1024 + (size_t)PLANE_BYTES + (size_t)9 * CIN * B_BYTES + (size_t)NU * 3 * CIN * CI_BYTES +
                      27 * 32 * 4 + 256 + 256;
  if (smem > 232448) return SSD3D_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(stem_dw_kernel<CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  long long grid = persistent_sms();
  if (grid > p.steps) grid = p.steps;
  SSD3D_LAUNCH_PDL((stem_dw_kernel<CIN>), dim3((unsigned)grid), dim3(THREADS), smem, st, tmE, tmO, p);
  return SSD3D_OK;
}

}  // namespace ssd3d

using namespace ssd3d;

// 1 when stem + first depthwise conv can run as one kernel: bf16 volumes, Cin <= 2, rows of 128 voxels, even stem map
extern "C" int ssd3d_stem_dw_fused_supported(int x_is_bf16, int Cin, int D, int H, int W, int stride_d) {
  if (!x_is_bf16 || Cin < 1 || Cin > 2 || W != 128) return 0;
  if (stride_d != 1 && stride_d != 2) return 0;
  if (H < 2 || D < 2 || (H & 1)) return 0;
  const int Ds = (D - 1) / stride_d + 1, Hs = (H - 1) / 2 + 1;
  return ((Ds & 1) == 0 && (Hs & 1) == 0) ? 1 : 0;
}

extern "C" int ssd3d_stem_dw_fused(const void* x, int x_is_bf16, const void* w_stem, const float* scale0,
                                   const float* shift0, const void* w_dw, const float* scale1, const float* shift1,
                                   void* y, int N, int Cin, int D, int H, int W, int stride_d, void* stream) {
  if (!x || !w_stem || !scale0 || !shift0 || !w_dw || !scale1 || !shift1 || !y || N <= 0) return SSD3D_ERR_ARG;
  if (!ssd3d_stem_dw_fused_supported(x_is_bf16, Cin, D, H, W, stride_d)) return SSD3D_ERR_UNSUPPORTED;
  StemDwParams p{};
  p.N = N; p.D = D; p.H = H; p.W = W; p.sd = stride_d;
  p.Ds = (D - 1) / stride_d + 1; p.Hs = (H - 1) / 2 + 1;
  p.Dd = (p.Ds - 1) / 2 + 1; p.Hd = (p.Hs - 1) / 2 + 1;
  p.HT = (p.Hd + sdw::TH - 1) / sdw::TH;
  const long long steps = (long long)N * p.HT * p.Dd;
  if (steps > 0x3fffffffll) return SSD3D_ERR_UNSUPPORTED;
  p.steps = (int)steps;
  p.kpad = (27 * Cin <= 64) ? 64 : 128;
  p.wt = static_cast<const __nv_bfloat16*>(w_stem);
  p.scale0 = scale0; p.shift0 = shift0;
  p.wd = static_cast<const __nv_bfloat16*>(w_dw);
  p.scale1 = scale1; p.shift1 = shift1;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.one = 1.0f;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return Cin == 1 ? launch_stem_dw<1>(x, p, st) : launch_stem_dw<2>(x, p, st);
}
