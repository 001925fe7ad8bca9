// Stem convolution (dense 3x3x3, Cin in {1,2} -> 32, stride (sd,2,2), + BN + ReLU; mobilenet.py:26-31 as
// instantiated at ssd3d.py:60-61) with NO per-thread tap gather: the W taps are folded into a banded ("Toeplitz")
// B operand, so that the A operand of the tcgen05 implicit GEMM is the raw input row exactly as TMA drops it into
// shared memory.
//
//   * A row of the GEMM is a GROUP of 4 consecutive output voxels along W (w = 4g .. 4g+3).  With stride 2 and
//     padding 1 they read input columns 8g-1 .. 8g+7 of one input row (ci, 2h+kh-1, sd*d+kd-1): 9 bf16 values
//     inside the 16 consecutive values x[8g-8 .. 8g+7], i.e. inside the two 16-byte chunks g and g+1 of the row.
//   * Shared memory holds, per (kd, kh), a TMA box of input rows [ci][dl][hl][72 columns] (traversal stride 2 along
//     H, sd along D: exactly the rows the tile's outputs need, dense).  The K-major no-swizzle UMMA descriptor of
//     the A operand uses row pitch 16 B inside a core matrix, LBO = 16 B (the second K chunk of row g IS the first
//     chunk of row g+1: the core matrices overlap) and SBO = 144 B (one 32-voxel row slot = 8 groups + the 9th
//     chunk).  M = 128 rows = 16 row slots of 32 output voxels = 512 voxels per accumulator.
//   * B[(j, co)][k] = w[co][ci][kd][kh][k - 7 - 2j] (zero outside 0..2): N = 4 voxels x 32 channels = 128 columns,
//     K = 16.  One tile = 9*Cin UMMAs (M128 N128 K16) accumulating into 128 TMEM columns.  The banded B costs
//     5.7x the MACs of the dense contraction, still only ~17 us of tensor pipe for the 2ch 128^3 batch-8 workload
//     (HBM floor 31 us), and it removes the 36 LDS + byte permutes + tcgen05.st per voxel that bound the gather
//     kernel (conv_stem_tc.cu) by instruction issue.
//   * A TMEM lane holds 4 voxels x 32 channels = 256 contiguous bytes of the channels-last output.  Writing them
//     from the lane's thread (32-byte vector stores, first version) made every warp store touch 32 different
//     lines: ncu counted 64 LSU wavefronts per store instruction, 58 k per SM and launch = the kernel's limit
//     (45 us).  Now each voxel PAIR (128 B) goes to a 128B-swizzled staging line (conflict free: chunk ^= row & 7)
//     and leaves through a TMA store per half tile; the two voxel pairs of a group are handled by two independent
//     sets of 4 epilogue warps with their own staging buffer, named barrier and tensor map (base shifted by 128 B).
//   * Warp-specialised persistent CTA (one per SM): warp 0 = TMA producer (ring of 6 kd groups of 3 boxes = 2 tiles
//     deep, released group by group), warp 1 = UMMA issuer, warps 2..9 = epilogue (two accumulators, so the epilogue of tile i overlaps the
//     UMMAs of tile i+1).  No CTA-wide barrier in the tile loop.
//
// A zero in B still multiplies its A value, so a NaN/Inf input contaminates the 4 voxels of its group instead
// of only its own receptive field; the network raises on the first NaN either way (ssd3d.py:248-263).
#include "common.cuh"
#include "tma_host.h"

namespace ssd3d {

struct StemTzParams {
  int N, D, H, W, Do, Ho, Wo, sd;
  int TH, TD;                  // row slots of one tile: TH * TD = 16
  int cols, tiles_h, tiles_d;  // cols = ceil(Wo / 32)
  int kpad;                    // row pitch of wt
  const __nv_bfloat16* wt;     // (32, kpad) bf16, k = ((ci*3+kd)*3+kh)*3+kw
  const float* scale;
  const float* shift;
  __nv_bfloat16* y;            // (N, Do, Ho, Wo, 32)
  float floor;                 // 0 = ReLU, -inf = identity
  int zero;                    // 0, as a run-time value (keeps the B descriptors out of loop-invariant hoisting)
};

namespace tz {

constexpr int SLOT_BYTES = 144;              // 72 input columns: 8 groups x 16 B + the 9th chunk
constexpr int ROWS = 16;                     // row slots per tile
constexpr int CI_BYTES = ROWS * SLOT_BYTES;  // 2304 (a multiple of 128: every TMA box lands 128-byte aligned)
constexpr int B_BYTES = 4096;                // one (ci,kd,kh) slice of B: 128 rows x 16 bf16
constexpr int NGROUP = 6;                    // ring of kd groups (3 (kd,kh) boxes each): 2 tiles
constexpr int OUT_HALF_BYTES = 16384;        // staging: 128 lines x 128 B (one voxel pair per GEMM row)
constexpr int THREADS = 64 + 2 * 256;         // TMA producer, UMMA issuer, two groups of 8 epilogue warps

__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// K-major operand without swizzle (cute::UMMA canonical layout ((8,m),(T,2)):((1T,SBO),(1,LBO))): 8 rows of a
// core matrix 16 B apart, the two K chunks LBO apart, 8-row groups SBO apart.
__device__ __forceinline__ uint64_t desc_k_nosw(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell); layout type 0 = no swizzle
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

__device__ __forceinline__ float relu_nan1(float v, float floor) {   // max.NaN: torch.relu semantics
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(v), "f"(floor));
  return r;
}

}  // namespace tz

// One tile's share of the producer / issuer loops, with the ring position a compile-time constant (TP = tile
// parity): every shared-memory address, descriptor and barrier is base + immediate.  The issuing warp is a single
// instruction stream that has to feed one UMMA per 64 tensor-pipe cycles: with run-time ring indices the ~45
// dependent integer / R2UR instructions per (kd,kh) box cost 360 cycles per box and bound the whole kernel.
template <int CIN, int TP>
__device__ __forceinline__ void tz_produce_tile(const CUtensorMap* tmX, uint32_t sA_u, uint64_t* full, uint64_t* empty,
                                                uint32_t par, int cw, int ch, int cd, int cn) {
  using namespace tz;
  constexpr int BOX_BYTES = CIN * CI_BYTES;
#pragma unroll
  for (int kd = 0; kd < 3; ++kd) {
    const int g = TP * 3 + kd;
    mbar_wait(&empty[g], par ^ 1u);
    if (elect_one()) {
      mbar_arrive_expect_tx(&full[g], (uint32_t)(3 * BOX_BYTES));
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
        tma_load_4d(sA_u + (uint32_t)((g * 3 + kh) * BOX_BYTES), tmX, &full[g], cw, ch + kh, cd + kd, cn);
    }
    __syncwarp();
  }
}

// Called by ONE lane.  The UMMA issue queue is shallow (scripts/ubench/umma_rate.cu: issue time == execution time, 64
// cycles per M128 N128 K16), so every cycle the issuer spends between two UMMAs on barrier waits, fences and commits
// is a cycle the tensor pipe idles: the wait for the NEXT group's boxes sits between the UMMAs of this group, behind
// UMMAs that are still executing.
template <int CIN>
__device__ __forceinline__ void tz_mma_tile(uint64_t da, uint64_t db0, uint64_t* full, uint64_t* empty,
                                            uint64_t* acc_full, uint32_t par, uint32_t dcol, uint32_t idesc) {
  // da / full / empty already point at this tile's half of the ring (a run-time offset on purpose: with the ring
  // position a template constant ptxas precomputes all 36 descriptor pairs in vector registers and pays four R2UR per
  // UMMA; relative to a per-tile uniform base they are UIADD3.64 with immediates)
  using namespace tz;
  constexpr int BOX_BYTES = CIN * CI_BYTES;
  mbar_wait(&full[0], par);
  tc_fence_after();
#pragma unroll
  for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci)
        umma_bf16_ss(dcol, da + (uint64_t)(((kd * 3 + kh) * BOX_BYTES + ci * CI_BYTES) >> 4),
                     db0 + (uint64_t)(((ci * 9 + kd * 3 + kh) * B_BYTES) >> 4), idesc, (kd | kh | ci) != 0 ? 1u : 0u);
      if (kh == 0 && kd < 2) {
        mbar_wait(&full[kd + 1], par);
        tc_fence_after();
      }
    }
    umma_commit(&empty[kd]);                     // the group's boxes are free once these UMMAs have read them
  }
  umma_commit(acc_full);                         // the accumulator is complete
}

template <int CIN>
__global__ void __launch_bounds__(tz::THREADS, 1) stem_tz_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                const __grid_constant__ CUtensorMap tmY0,
                                                                const __grid_constant__ CUtensorMap tmY1,
                                                                const StemTzParams p) {
  using namespace tz;
  constexpr int NT = 9 * CIN;                       // UMMAs per tile
  constexpr int BOX_BYTES = CIN * CI_BYTES;         // one (kd,kh) box
  constexpr uint32_t TMEM_COLS = 512;               // four 128-column accumulators

  extern __shared__ uint8_t tz_raw[];
  const uint32_t raw = smem_u32(tz_raw);
  uint8_t* smem = tz_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* sOut = smem;                                         // 2 sets x 2 halves x 16 KB, 1024-aligned (128B swizzle)
  uint8_t* sB = sOut + 4 * OUT_HALF_BYTES;                      // NT x 4096
  uint8_t* sA = sB + NT * B_BYTES;                              // NGROUP x 3 x BOX_BYTES
  uint64_t* full = reinterpret_cast<uint64_t*>(sA + NGROUP * 3 * BOX_BYTES);
  uint64_t* empty = full + NGROUP;
  uint64_t* acc_full = empty + NGROUP;                          // [4]
  uint64_t* acc_empty = acc_full + 4;                           // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 4);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int total_tiles = p.cols * p.tiles_h * p.tiles_d * p.N;

  if (tid == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY0);
    tma_prefetch_desc(&tmY1);
    for (int s = 0; s < NGROUP; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
    fence_barrier_init();
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  pdl_wait();                 // weights (training repacks them), the input and the output buffer's readers
  pdl_launch_dependents();

  // ---- banded B operand: zero fill, then scatter the 27*CIN*32 weights to their 4 voxel positions ----
  for (int i = tid; i < NT * B_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(sB)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  for (int i = tid; i < NT * 3 * 32; i += THREADS) {
    const int co = i & 31, tk = i >> 5;             // tk = t*3 + kw = the weight's k index
    const int t = tk / 3, kw = tk - 3 * t;
    const __nv_bfloat16 wv = p.wt[co * p.kpad + tk];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = j * 32 + co, k = 7 + 2 * j + kw;
      *reinterpret_cast<__nv_bfloat16*>(sB + t * B_BYTES + (n >> 3) * 256 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) = wv;
    }
  }
  if (tid < 32) {
    float* s_sc = reinterpret_cast<float*>(tmem_slot + 4);
    s_sc[tid] = __ldg(p.scale + tid);
    s_sc[32 + tid] = __ldg(p.shift + tid);
  }
  fence_proxy_async_smem();   // generic-proxy writes of B -> visible to the UMMA (async proxy) reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    const uint32_t sA_u = smem_u32(sA);
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int t = tile;
      const int col = t % p.cols; t /= p.cols;
      const int h0 = (t % p.tiles_h) * p.TH; t /= p.tiles_h;
      const int d0 = (t % p.tiles_d) * p.TD; t /= p.tiles_d;
      const int cw = 64 * col - 8, ch = 2 * h0 - 1, cd = p.sd * d0 - 1, cn = t * CIN;
      const uint32_t par = (uint32_t)((it >> 1) & 1);
      if (it & 1) tz_produce_tile<CIN, 1>(&tmX, sA_u, full, empty, par, cw, ch, cd, cn);
      else tz_produce_tile<CIN, 0>(&tmX, sA_u, full, empty, par, cw, ch, cd, cn);
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer: one lane, descriptors in uniform registers =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint64_t da0 = desc_k_nosw(smem_u32(sA), 16u, (uint32_t)SLOT_BYTES);
      const uint64_t db0 = desc_k_nosw(smem_u32(sB), 128u, 256u);
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t par = (uint32_t)((it >> 1) & 1);
        const int ab = it & 3;
        mbar_wait(&acc_empty[ab], (uint32_t)(((it >> 2) & 1) ^ 1));
        const uint32_t dcol = tmem_base + (uint32_t)(ab * 128);
        const int tp = it & 1;
        tz_mma_tile<CIN>(da0 + (uint64_t)((uint32_t)(tp * 9 * BOX_BYTES) >> 4), db0 + (uint64_t)(uint32_t)(it & p.zero),
                         full + tp * 3, empty + tp * 3,
                         &acc_full[ab], par, dcol, idesc);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue: two groups (even / odd tiles) x two half-sets (voxel pair 0 / 1) of 4 warps =====================
    // A tile's epilogue is a latency chain (accumulator wait, TMEM load, BN, staging, fence, barrier, TMA store:
    // ~2700 cycles against the 2440 the HBM write stream needs per tile), so two tiles are in the epilogue at once.
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int grp = (warp - 2) >> 3;              // tiles it = grp, grp + 2, ..
    const int half = ((warp - 2) >> 2) & 1;       // voxels j = 2*half, 2*half + 1 of the group
    const int r = q * 32 + lane;                  // GEMM row = group = staging line
    const bool leader = (r == 0);
    const CUtensorMap* tmY = half ? &tmY1 : &tmY0;
    float* s_sc = reinterpret_cast<float*>(tmem_slot + 4);   // BN scale[32], shift[32] (filled before the set-up barrier)
    float* s_sh = s_sc + 32;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64);
    const uint32_t out_base = smem_u32(sOut + (grp * 2 + half) * OUT_HALF_BYTES);   // this group's staging half
    const uint32_t line = out_base + (uint32_t)(r * 128);
    const uint32_t sw = (uint32_t)(r & 7);
    const int bar_id = 1 + grp * 2 + half;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != grp) continue;
      const int ab = it & 3;
      mbar_wait(&acc_full[ab], (uint32_t)((it >> 2) & 1));
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld_x32(lane_addr + (uint32_t)(ab * 128), v0);
      tmem_ld_x32(lane_addr + (uint32_t)(ab * 128 + 32), v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);   // the accumulator can be overwritten
      // the group's previous TMA store (two tiles ago) must have read the staging half before it is overwritten
      if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
      for (int e = 0; e < 2; ++e) {
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 sc0 = *reinterpret_cast<const float4*>(s_sc + 8 * c4), sc1 = *reinterpret_cast<const float4*>(s_sc + 8 * c4 + 4);
          const float4 sh0 = *reinterpret_cast<const float4*>(s_sh + 8 * c4), sh1 = *reinterpret_cast<const float4*>(s_sh + 8 * c4 + 4);
          const uint32_t* v = e ? v1 : v0;
          const int k = 8 * c4;
          const uint32_t o0 = pack_bf16x2(relu_nan1(__fadd_rn(__fmul_rn(__uint_as_float(v[k]), sc0.x), sh0.x), p.floor),
                                          relu_nan1(__fadd_rn(__fmul_rn(__uint_as_float(v[k + 1]), sc0.y), sh0.y), p.floor));
          const uint32_t o1 = pack_bf16x2(relu_nan1(__fadd_rn(__fmul_rn(__uint_as_float(v[k + 2]), sc0.z), sh0.z), p.floor),
                                          relu_nan1(__fadd_rn(__fmul_rn(__uint_as_float(v[k + 3]), sc0.w), sh0.w), p.floor));
          const uint32_t o2 = pack_bf16x2(relu_nan1(__fadd_rn(__fmul_rn(__uint_as_float(v[k + 4]), sc1.x), sh1.x), p.floor),
                                          relu_nan1(__fadd_rn(__fmul_rn(__uint_as_float(v[k + 5]), sc1.y), sh1.y), p.floor));
          const uint32_t o3 = pack_bf16x2(relu_nan1(__fadd_rn(__fmul_rn(__uint_as_float(v[k + 6]), sc1.z), sh1.z), p.floor),
                                          relu_nan1(__fadd_rn(__fmul_rn(__uint_as_float(v[k + 7]), sc1.w), sh1.w), p.floor));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(line + (((uint32_t)(4 * e + c4) ^ sw) << 4)), "r"(o0),
                       "r"(o1), "r"(o2), "r"(o3)
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (leader) {
        int t = tile;
        const int col = t % p.cols; t /= p.cols;
        const int h0 = (t % p.tiles_h) * p.TH; t /= p.tiles_h;
        const int d0 = (t % p.tiles_d) * p.TD; t /= p.tiles_d;
        asm volatile(
            "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                reinterpret_cast<uint64_t>(tmY)),
            "r"(out_base), "r"(0), "r"(8 * col), "r"(h0), "r"(d0), "r"(t)
            : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every output tile has landed
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
static int launch_stem_tz(const void* x, const StemTzParams& p, cudaStream_t st) {
  using namespace tz;
  CUtensorMap tm;
  {
    const uint64_t dims[4] = {(uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.D, (uint64_t)p.N * CIN};
    const uint64_t strides[3] = {(uint64_t)p.W * 2, (uint64_t)p.W * p.H * 2, (uint64_t)p.W * p.H * p.D * 2};
    // rows 2*h0+kh-1, +2, ..: TH of them; planes sd*d0+kd-1, +sd, ..: TD of them
    const uint32_t box[4] = {72u, (uint32_t)(2 * p.TH - 1), (uint32_t)(p.sd * (p.TD - 1) + 1), (uint32_t)CIN};
    const uint32_t estr[4] = {1u, box[1] > 1 ? 2u : 1u, box[2] > 1 ? (uint32_t)p.sd : 1u, 1u};
    if (make_tma_bf16(&tm, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, estr)) return SSD3D_ERR_TMA;
  }
  // output: one voxel pair (2 x 32 channels = 128 B) per element row; pair `half` of every group of 4 voxels through
  // its own map (base shifted by 128 B): dims (64, Wo/4 groups, Ho, Do, N)
  CUtensorMap ty[2];
  for (int half = 0; half < 2; ++half) {
    const uint64_t dims[5] = {64ull, (uint64_t)(p.Wo / 4), (uint64_t)p.Ho, (uint64_t)p.Do, (uint64_t)p.N};
    const uint64_t strides[4] = {256ull, (uint64_t)p.Wo * 64, (uint64_t)p.Ho * p.Wo * 64, (uint64_t)p.Do * p.Ho * p.Wo * 64};
    const uint32_t box[5] = {64u, 8u, (uint32_t)p.TH, (uint32_t)p.TD, 1u};
    if (make_tma_bf16(&ty[half], p.y + half * 64, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return SSD3D_ERR_TMA;
  }
  const size_t smem = 1024 + 4 * (size_t)OUT_HALF_BYTES + (size_t)9 * CIN * B_BYTES + (size_t)NGROUP * 3 * CIN * CI_BYTES + 512;
  cudaError_t e = cudaFuncSetAttribute(stem_tz_kernel<CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const long long tiles = (long long)p.cols * p.tiles_h * p.tiles_d * p.N;
  if (tiles > 0x3fffffffll) return SSD3D_ERR_UNSUPPORTED;
  long long grid = persistent_sms();
  if (grid > tiles) grid = tiles;
  SSD3D_LAUNCH_PDL((stem_tz_kernel<CIN>), dim3((unsigned)grid), dim3(THREADS), smem, st, tm, ty[0], ty[1], p);
  return SSD3D_OK;
}

}  // namespace ssd3d

using namespace ssd3d;

// 1 when the banded-B stem kernel can take this input: bf16 volumes, Cin <= 2, rows of a multiple of 8 voxels
// (TMA row pitch and the 64*col - 8 box start stay 16-byte aligned)
extern "C" int ssd3d_stem_tz_supported(int x_is_bf16, int Cin, int W) {
  return (x_is_bf16 && Cin >= 1 && Cin <= 2 && W >= 8 && W % 8 == 0) ? 1 : 0;
}

extern "C" int ssd3d_stem_conv_affine_tz(const void* x, int x_is_bf16, const void* w, const float* scale,
                                         const float* shift, void* y, int N, int Cin, int D, int H, int W, int stride_d,
                                         int relu, void* stream) {
  if (!x || !w || !scale || !shift || !y || N <= 0 || D <= 0 || H <= 0 || W <= 0) return SSD3D_ERR_ARG;
  if (stride_d != 1 && stride_d != 2) return SSD3D_ERR_ARG;
  if (!ssd3d_stem_tz_supported(x_is_bf16, Cin, W)) return SSD3D_ERR_UNSUPPORTED;
  StemTzParams p{};
  p.N = N; p.D = D; p.H = H; p.W = W; p.sd = stride_d;
  p.Do = (D - 1) / stride_d + 1; p.Ho = (H - 1) / 2 + 1; p.Wo = (W - 1) / 2 + 1;
  int th = 16;
  while (th > 1 && th / 2 >= p.Ho) th >>= 1;     // smallest power of two >= Ho, at most 16
  p.TH = th;
  p.TD = 16 / th;
  p.cols = (p.Wo + 31) / 32;
  p.tiles_h = (p.Ho + p.TH - 1) / p.TH;
  p.tiles_d = (p.Do + p.TD - 1) / p.TD;
  p.kpad = (27 * Cin <= 64) ? 64 : 128;
  p.wt = static_cast<const __nv_bfloat16*>(w);
  p.scale = scale;
  p.shift = shift;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.floor = SSD3D_FLOOR(relu);
  p.zero = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return Cin == 1 ? launch_stem_tz<1>(x, p, st) : launch_stem_tz<2>(x, p, st);
}
