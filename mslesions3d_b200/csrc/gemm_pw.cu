// Persistent pointwise GEMM (mobilenet.py:40,45):  y[M, Cout] = act((x[M, Cin] . w[Cout, Cin]^T) * scale + shift)
// for the large maps, where the one-tile-per-CTA kernel of gemm_tc.cu spends most of its time in per-CTA latency
// (barrier init, TMEM allocation, one TMA round trip, one commit, the epilogue -- all in sequence: 28 % of the HBM
// peak on the 262 144 x 32 -> 64 layer).  Here a CTA walks tiles t = blockIdx.x, + gridDim.x, ...:
//   warp 8   TMA producer: keeps a ring of {A,B} stages full ACROSS tile boundaries
//   warp 9   UMMA issuer: accumulates tile j into TMEM buffer j & 1 (2 x BN fp32 columns)
//   warps 0-7 epilogue, TWO warps per TMEM lane quarter (columns 0-31 / 32-63 of the row: ncu showed the four-warp
//            epilogue, not the loads or the MMAs, setting the tile rate -- barrier and long-scoreboard stalls):
//            drain buffer j & 1 (tcgen05.ld, BN scale/shift from smem, activation floor, bf16 pack)
//            into a 128B-swizzled staging tile and hand it to ONE TMA store (cp.async.bulk.tensor ... global.shared)
//            while the issuer already works on tile j + 1.  (A thread owns a row, so direct stores would make every
//            warp-wide 16-byte store touch 32 different 128-byte lines: half-used sectors, 8 instructions per row.)
// so loads, MMAs and stores of consecutive tiles overlap and the kernel streams at the memory system's pace.
// The tile is 128 rows x 64 output channels (one swizzle atom per row).
#include "common.cuh"
#include "tma_host.h"

namespace ssd3d {

struct PwParams {
  int num_kb, BN, stages, tmem_cols;
  int m_tiles, n_tiles;
  long long M;
  int Cout;
  __nv_bfloat16* y;
  const float* scale;
  const float* shift;
  float floor;
  int* nan_flag;
};

template <int BK>
__device__ __forceinline__ uint64_t pw_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * BK * 2) >> 4) << 32;            // SBO: bytes between 8-row groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)((BK == 64) ? 2u : 4u) << 61;         // SWIZZLE_128B : SWIZZLE_64B
  return d;
}

template <int BK>
__global__ void __launch_bounds__(320) gemm_pw_persistent_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmB,
                                                                    const __grid_constant__ CUtensorMap tmY,
                                                                    const PwParams p) {
  extern __shared__ uint8_t pw_raw[];
  const uint32_t raw = smem_u32(pw_raw);
  uint8_t* smem = pw_raw + ((1024u - (raw & 1023u)) & 1023u);
  constexpr int A_BYTES = 128 * BK * 2;
  const int B_BYTES = p.BN * BK * 2;
  const int STAGE_BYTES = A_BYTES + B_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * STAGE_BYTES);
  uint64_t* empty = full + p.stages;
  uint64_t* tmem_full = empty + p.stages;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  // after the barriers: BN scale / shift of all output channels, then two 16 KB staging tiles (1024-aligned)
  float* s_scale = reinterpret_cast<float*>(smem + (size_t)p.stages * STAGE_BYTES + 256);
  float* s_shift = s_scale + p.Cout;
  uint8_t* stage_out = smem + (((size_t)p.stages * STAGE_BYTES + 256 + (size_t)p.Cout * 8 + 1023) & ~(size_t)1023);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmY);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 256); }
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) { s_scale[i] = p.scale[i]; s_shift[i] = p.shift[i]; }
  __syncthreads();

  const int total_tiles = p.m_tiles * p.n_tiles;
  if (warp == 8) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int m0 = (t / p.n_tiles) * 128, n0 = (t % p.n_tiles) * p.BN;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int s = it % p.stages, use = it / p.stages;
          if (use > 0) mbar_wait(&empty[s], (uint32_t)((use - 1) & 1));
          uint8_t* a_dst = smem + (size_t)s * STAGE_BYTES;
          mbar_arrive_expect_tx(&full[s], (uint32_t)STAGE_BYTES);
          tma_load_2d(a_dst, &tmA, &full[s], kb * BK, m0);
          tma_load_2d(a_dst + A_BYTES, &tmB, &full[s], kb * BK, n0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, p.BN);
      int it = 0, j = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
        const int ab = j & 1, u = j >> 1;
        if (u > 0) {
          mbar_wait(&tmem_empty[ab], (uint32_t)((u - 1) & 1));     // the epilogue has drained this buffer
          tc_fence_after();
        }
        const uint32_t acc = tmem_base + (uint32_t)(ab * p.BN);
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int s = it % p.stages;
          mbar_wait(&full[s], (uint32_t)((it / p.stages) & 1));
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)s * STAGE_BYTES);
          const uint64_t da = pw_desc<BK>(a_addr), db = pw_desc<BK>(a_addr + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_ss(acc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        umma_commit(&tmem_full[ab]);
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3, hf = warp >> 2;          // TMEM lane quarter, column half (32 columns) of this warp
    const int row = q * 32 + lane;
    bool bad = false;
    int j = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
      const int ab = j & 1, u = j >> 1;
      const int m0 = (t / p.n_tiles) * 128, n0 = (t % p.n_tiles) * 64;
      mbar_wait(&tmem_full[ab], (uint32_t)(u & 1));
      __syncwarp();
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * 64 + hf * 32);
      uint32_t v[32];
      tmem_ld_32x32b_x16(taddr, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
      tmem_ld_32x32b_x16(taddr + 16u, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&tmem_empty[ab]);        // accumulator is in registers: 256 arrivals release the TMEM buffer
      // staging buffer ab was handed to a TMA store two tiles ago: wait until that store has READ it
      if (threadIdx.x == 0 && u > 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      uint8_t* dst_row = stage_out + (size_t)ab * 16384 + (size_t)row * 128;
#pragma unroll
      for (int c = 0; c < 32; c += 8) {
        const int cc = hf * 32 + c;            // column inside the 64-wide tile
        const float4 sc0 = *reinterpret_cast<const float4*>(s_scale + n0 + cc);
        const float4 sc1 = *reinterpret_cast<const float4*>(s_scale + n0 + cc + 4);
        const float4 sh0 = *reinterpret_cast<const float4*>(s_shift + n0 + cc);
        const float4 sh1 = *reinterpret_cast<const float4*>(s_shift + n0 + cc + 4);
        const float r0 = clamp_floor(__fadd_rn(__fmul_rn(__uint_as_float(v[c + 0]), sc0.x), sh0.x), p.floor);
        const float r1 = clamp_floor(__fadd_rn(__fmul_rn(__uint_as_float(v[c + 1]), sc0.y), sh0.y), p.floor);
        const float r2 = clamp_floor(__fadd_rn(__fmul_rn(__uint_as_float(v[c + 2]), sc0.z), sh0.z), p.floor);
        const float r3 = clamp_floor(__fadd_rn(__fmul_rn(__uint_as_float(v[c + 3]), sc0.w), sh0.w), p.floor);
        const float r4 = clamp_floor(__fadd_rn(__fmul_rn(__uint_as_float(v[c + 4]), sc1.x), sh1.x), p.floor);
        const float r5 = clamp_floor(__fadd_rn(__fmul_rn(__uint_as_float(v[c + 5]), sc1.y), sh1.y), p.floor);
        const float r6 = clamp_floor(__fadd_rn(__fmul_rn(__uint_as_float(v[c + 6]), sc1.z), sh1.z), p.floor);
        const float r7 = clamp_floor(__fadd_rn(__fmul_rn(__uint_as_float(v[c + 7]), sc1.w), sh1.w), p.floor);
        // NaN check on the sum: NaN iff any term is NaN (or +inf and -inf meet, which the next layer turns into NaN anyway)
        const float chk = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
        bad |= (chk != chk) && ((r0 != r0) | (r1 != r1) | (r2 != r2) | (r3 != r3) | (r4 != r4) | (r5 != r5) | (r6 != r6) | (r7 != r7));
        // 16-byte chunk (cc / 8) of this row, 128-byte swizzle: chunk ^ (row & 7)
        *reinterpret_cast<uint4*>(dst_row + ((((cc >> 3) ^ (row & 7))) << 4)) =
            make_uint4(pack_bf16x2(r0, r1), pack_bf16x2(r2, r3), pack_bf16x2(r4, r5), pack_bf16x2(r6, r7));
      }
      fence_proxy_async_smem();            // generic-proxy writes -> visible to the TMA (async proxy)
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                         reinterpret_cast<uint64_t>(&tmY)),
                     "r"(smem_u32(stage_out + (size_t)ab * 16384)), "r"(n0), "r"(m0)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete
    if (bad && p.nan_flag) atomicOr(p.nan_flag, SSD3D_NAN_BACKBONE);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

static inline int pw_pow2_ceil(int v) {
  int r = 32;
  while (r < v) r <<= 1;
  return r;
}

}  // namespace ssd3d

using namespace ssd3d;

// Returns SSD3D_ERR_UNSUPPORTED when the problem has too few tiles to be worth a persistent grid.
int ssd3d_pwconv_persistent(const void* x, const void* w, const float* scale, const float* shift, void* y, int64_t M,
                            int Cin, int Cout, float floor, int* nan_flag, cudaStream_t st) {
  const int n_sm = persistent_sms();
  const int BK = (Cin % 64 == 0) ? 64 : 32;
  if (Cout % 64 || Cout > 1024) return SSD3D_ERR_UNSUPPORTED;
  const int BN = 64;                                  // one 128-byte swizzle atom per output row
  const long long m_tiles = (M + 127) / 128;
  const long long tiles = m_tiles * (Cout / BN);
  if (tiles < 2ll * n_sm || m_tiles > 0x7fffffff / (Cout / BN)) return SSD3D_ERR_UNSUPPORTED;
  PwParams p{};
  p.num_kb = Cin / BK;
  p.BN = BN;
  p.m_tiles = (int)m_tiles;
  p.n_tiles = Cout / BN;
  p.M = M; p.Cout = Cout;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.scale = scale; p.shift = shift; p.floor = floor; p.nan_flag = nan_flag;
  p.tmem_cols = pw_pow2_ceil(2 * BN);
  // CTAs per SM: the epilogue (4 warps per CTA) is the slow stage, so several CTAs share an SM; bounded by the
  // 512 TMEM columns and by ~200 KB of shared memory for the operand rings
  int per_sm = 512 / p.tmem_cols;
  if (per_sm > 4) per_sm = 4;
  const int stage_bytes = 128 * BK * 2 + BN * BK * 2;
  const int fixed = 256 + Cout * 8 + 1024 + 2 * 16384 + 1024;     // barriers, scale/shift, 2 staging tiles, alignment
  if (per_sm > 3) per_sm = 3;
  int stages = 0;
  for (; per_sm >= 1; --per_sm) {
    stages = ((220 * 1024) / per_sm - fixed) / stage_bytes;
    if (stages >= 3 || per_sm == 1) break;
  }
  if (stages > 8) stages = 8;
  if (stages < 2) return SSD3D_ERR_UNSUPPORTED;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + fixed;
  CUtensorMap tmA, tmB;
  const CUtensorMapSwizzle sw = (BK == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  {
    const uint64_t dims[2] = {(uint64_t)Cin, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)Cin * 2};
    const uint32_t box[2] = {(uint32_t)BK, 128u};
    if (make_tma_bf16(&tmA, x, 2, dims, strides, box, sw)) return SSD3D_ERR_TMA;
  }
  {
    const uint64_t dims[2] = {(uint64_t)Cin, (uint64_t)Cout};
    const uint64_t strides[1] = {(uint64_t)Cin * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)BN};
    if (make_tma_bf16(&tmB, w, 2, dims, strides, box, sw)) return SSD3D_ERR_TMA;
  }
  CUtensorMap tmY;
  {
    const uint64_t dims[2] = {(uint64_t)Cout, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)Cout * 2};
    const uint32_t box[2] = {64u, 128u};
    if (make_tma_bf16(&tmY, y, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return SSD3D_ERR_TMA;
  }
  const long long want = (long long)per_sm * n_sm;
  const unsigned grid = (unsigned)(tiles < want ? tiles : want);
  cudaError_t e;
  if (BK == 64) {
    e = cudaFuncSetAttribute(gemm_pw_persistent_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    SSD3D_LAUNCH_PDL(gemm_pw_persistent_kernel<64>, dim3(grid), dim3(320), smem, st, tmA, tmB, tmY, p);
  } else {
    e = cudaFuncSetAttribute(gemm_pw_persistent_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    SSD3D_LAUNCH_PDL(gemm_pw_persistent_kernel<32>, dim3(grid), dim3(320), smem, st, tmA, tmB, tmY, p);
  }
  return SSD3D_OK;
}
