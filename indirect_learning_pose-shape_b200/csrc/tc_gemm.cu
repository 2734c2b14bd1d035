// K1 on the 5th-generation tensor cores: error-compensated split-operand GEMMs,  D[M][N] = A[M][K] * B[N][K]^T (+ bias[N]).
//
// Both blend products of the SMPL layer are this GEMM (batch_smpl.py:106-108,126-128 forward, their TF autodiff
// backward):
//   forward   v_posed[N][LD]  = X[N][224]     * BT[LD][224]^T + v_template      (BT = [shapedirs; posedirs]^T, K-major)
//   backward  g_X[N][224]     = g_vp[N][Kp]   * Bs[224][Kp]^T
// fp32 parity (1e-5 on vertices, ~3 ulp on projections) rules out a plain TF32 or fp16 product, so every operand is
// supplied as an exact split x = hi + lo and the product is accumulated as lo*hi + hi*lo + hi*hi (the dropped lo*lo
// term is 2^-22 relative):
//   backward  3xTF32 (split3_gemm_kernel<.., F16 = false>): hi = x with its 13 low mantissa bits cleared, lo = x - hi;
//             hi is exactly representable in TF32, so the result does not depend on how the tensor core rounds inputs.
//             The A operand is the caller's gradient: no fixed scale would keep an fp16 split in range.
//   forward   fp16 split (blend_f16_panel_kernel; kind::f16 runs at twice the TF32 rate on half the operand bytes): the
//             blend coefficients are scaled by 2^6 and the blend matrix by the power of two that puts max |B| in
//             [2^13, 2^14) before hi = fp16(x), lo = fp16(x - hi): 11 + 11 bits, fp16 subnormals bound lo's own error at
//             1e-9 absolute, fp16 x fp16 products are exact in the fp32 accumulator, and the scales come off exactly in
//             the epilogue.  A degenerate model (no finite non-zero blend shape) keeps the forward on 3xTF32 (api.cu).
// The tensor core's own fp32 accumulation is not round-to-nearest, and its error grows with the length of the
// accumulation chain (measured here: 1.1e-6 on vertices at K=224, 1e-4 relative on gradients at K=4160 when the whole
// K loop accumulates in TMEM).  So TMEM only ever accumulates ONE K block (128 bytes of K: 12 MMAs); the epilogue warps
// add each block's partial tile into fp32 REGISTER accumulators with round-to-nearest (Ootomo & Yokota's scheme), while
// the tensor core works on the next blocks in the other TMEM buffers.
//
// Structure of split3_gemm_kernel (one CTA per SM, persistent over output tiles of 128 x BN):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d of the four operand boxes (A_hi, A_lo, B_hi, B_lo; 128 B of K
//               wide, SWIZZLE_128B) into a 3-stage ring, completion on a transaction mbarrier
//   warp 1      MMA issuer: one elected thread issues 12 tcgen05.mma.cta_group::1 (M=128, N=BN, 32 B of K each) per K
//               block into one of four TMEM buffers; tcgen05.commit releases the smem stage and publishes the buffer.
//               Owns the TMEM allocation.
//   warps 2-9   accumulate + epilogue, two warps per TMEM lane quadrant (one per half of the columns): all tcgen05.ld
//               32x32b.x16 of the block's partial half tile in flight, ONE wait, FADD into BN/2 registers per thread (half
//               an output row each); after the last block add bias, st.global.v8.f32 (full 32-byte sectors)
// blend_f16_panel_kernel keeps the same roles and adds a resident B panel (see its own comment).
// Descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <mutex>
#include "common.cuh"

namespace smplb200 {

namespace {

constexpr int kBM = 128;         // rows of an output tile (UMMA M, cta_group::1)
constexpr int kBK = 32;          // fp32 per K block = one 128-byte swizzle row
constexpr int kStages = 3;
constexpr int kThreads = 320;    // 10 warps: TMA, MMA, 2 x 4 accumulate/epilogue (each group owns half of the columns)
constexpr int kTmemBufs = 4;     // partial-tile buffers in TMEM: the tensor core may run this many K blocks ahead of the adds
constexpr int kTmemCols = 512;   // kTmemBufs buffers of up to 128 columns = all of TMEM (one CTA per SM)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {   // arrives on `bar` when all prior MMAs have completed
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <bool F16>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  if (F16) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// K-major, SWIZZLE_128B operand tile: rows 128 B apart, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);     // [0,14)  start address >> 4
  d |= (uint64_t)1 << 16;                       // [16,30) leading byte offset >> 4 (ignored for swizzled K-major; 1)
  d |= (uint64_t)(1024 >> 4) << 32;             // [32,46) stride byte offset >> 4: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                       // [46,48) descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                       // [61,64) layout type: SWIZZLE_128B
  return d;
}
// cute::UMMA::InstrDescriptor for kind::tf32 (operand format 2 = TF32) or kind::f16 (format 0 = F16), fp32 accumulate,
// both operands K-major.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool f16) {
  return (1u << 4)                 // [4,6)   c_format = F32
         | ((f16 ? 0u : 2u) << 7)  // [7,10)  a_format
         | ((f16 ? 0u : 2u) << 10) // [10,13) b_format
         | ((uint32_t)(N >> 3) << 17)   // [17,23) n_dim
         | ((uint32_t)(M >> 4) << 24);  // [24,29) m_dim
}

// F16: operands are fp16 hi/lo pairs (64 per 128-byte K block, UMMA K = 16); the accumulated sum is multiplied by
// out_scale (the inverse of the operands' power-of-two scales) before the bias.
template <int BN, bool BIAS, bool F16, bool DENSE>   // DENSE: the general epilogue of the regressor's layers (ncols / epi)
__global__ void __launch_bounds__(kThreads, 1)
split3_gemm_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                   float* __restrict__ D, int ldd, const float* __restrict__ bias, int M, int kblocks, int tiles_m,
                   int tiles_n, float out_scale, int ksplit, size_t plane_stride, int ncols, int epi) {
  // ncols: valid output columns (the last column tile may hang over it); epi: bit 0 = ReLU, bit 1 = accumulate into D,
  // bit 2 = D rows are 32-byte aligned with whole 8-column groups (vector stores)
  // Work unit = (output tile, K slice s of ksplit): slice s covers K blocks [s kblocks / ksplit, (s+1) kblocks / ksplit)
  // and writes its partial tile to plane s of D (D + s * plane_stride); the consumer adds the planes in fixed order
  // (deterministic, unlike atomics).  ksplit = 1: the plain GEMM.  The backward blend product has only 2 N / 128 output
  // tiles of K = 4160: without the split a 2048-sample shard kept 32 of 148 SMs busy.
  constexpr int kBKe = F16 ? 2 * kBK : kBK;            // operand elements per K block (one 128-byte swizzle row)
  constexpr int kABytes = kBM * kBK * 4, kBBytes = BN * kBK * 4;
  constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  constexpr int kAccStride = 128;                      // TMEM columns between partial-tile buffers
  static_assert(BN % 16 == 0 && BN <= 128, "BN / 2 registers per epilogue thread");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;                   // partial tile of one K block ready in TMEM buffer b
  uint64_t* tempty = tfull + kTmemBufs;                // TMEM buffer b drained into registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kTmemBufs);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = tiles_m * tiles_n * ksplit;       // work units

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < kTmemBufs; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {                                     // TMEM allocation is warp-collective; this warp also frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < ntiles; unit += gridDim.x) {
        const int tile = unit / ksplit, ks = unit - tile * ksplit;
        const int m0 = (tile / tiles_n) * kBM, n0 = (tile % tiles_n) * BN;
        const int kb0 = (int)(((long long)kblocks * ks) / ksplit), kb1 = (int)(((long long)kblocks * (ks + 1)) / ksplit);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * kStageBytes;
          mbar_expect_tx(&full[stage], kStageBytes);
          tma_load_2d(st, &tmAh, &full[stage], kb * kBKe, m0);
          tma_load_2d(st + kABytes, &tmAl, &full[stage], kb * kBKe, m0);
          tma_load_2d(st + 2 * kABytes, &tmBh, &full[stage], kb * kBKe, n0);
          tma_load_2d(st + 2 * kABytes + kBBytes, &tmBl, &full[stage], kb * kBKe, n0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kBM, BN, F16);
      int stage = 0, buf = 0;
      uint32_t phase = 0, buf_phase = 0;
      for (int unit = blockIdx.x; unit < ntiles; unit += gridDim.x) {
        const int ks = unit % ksplit;
        const int kb0 = (int)(((long long)kblocks * ks) / ksplit), kb1 = (int)(((long long)kblocks * (ks + 1)) / ksplit);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&tempty[buf], buf_phase ^ 1);      // the accumulate warps have drained this TMEM buffer
          mbar_wait(&full[stage], phase);              // TMA has landed this stage
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tmem_d = tmem_base + (uint32_t)(buf * kAccStride);
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t dAh = make_smem_desc(sa), dAl = make_smem_desc(sa + kABytes);
          const uint64_t dBh = make_smem_desc(sa + 2 * kABytes), dBl = make_smem_desc(sa + 2 * kABytes + kBBytes);
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {          // UMMA K = 8 tf32 / 16 f16 = 32 bytes: advance the start address by 2 (x16 B)
            const uint64_t o = (uint64_t)(k * 2);
            umma<F16>(tmem_d, dAl + o, dBh + o, idesc, k ? 1u : 0u);           // small terms first; fresh per K block
            umma<F16>(tmem_d, dAh + o, dBl + o, idesc, 1u);
          }
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {
            const uint64_t o = (uint64_t)(k * 2);
            umma<F16>(tmem_d, dAh + o, dBh + o, idesc, 1u);
          }
          umma_commit(&empty[stage]);                  // frees this smem stage once the MMAs above have read it
          umma_commit(&tfull[buf]);                    // this K block's partial tile is complete
          if (++stage == kStages) { stage = 0; phase ^= 1; }
          if (++buf == kTmemBufs) { buf = 0; buf_phase ^= 1; }
        }
      }
    }
  } else {
    // ===== accumulate (fp32 registers, round-to-nearest) + epilogue =====
    // Eight warps: warp w may only touch TMEM lanes [32*(w%4), +32), so warps 2-5 take columns [0, BN/2) and warps 6-9
    // columns [BN/2, BN) of the same lanes.  Every thread issues ALL its tcgen05.ld of a K block before the single wait:
    // the drain of a partial tile was four serialised load->wait round trips per warp and bound the kernel.
    constexpr int HN = BN / 2;
    static_assert(HN % 8 == 0, "half rows leave as 32-byte stores");
    const int q = warp & 3;
    const int hcol = (warp - 2) >= 4 ? HN : 0;
    int buf = 0;
    uint32_t buf_phase = 0;
    for (int unit = blockIdx.x; unit < ntiles; unit += gridDim.x) {
      const int tile = unit / ksplit, ks = unit - tile * ksplit;
      const int m0 = (tile / tiles_n) * kBM, n0 = (tile % tiles_n) * BN;
      const int kb0 = (int)(((long long)kblocks * ks) / ksplit), kb1 = (int)(((long long)kblocks * (ks + 1)) / ksplit);
      float acc[HN];
#pragma unroll
      for (int i = 0; i < HN; ++i) acc[i] = 0.f;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&tfull[buf], buf_phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kAccStride) + (uint32_t)hcol;
        uint32_t r[HN / 16 + 1][16];
#pragma unroll
        for (int c = 0; c < HN; c += 16) {
          if (c + 16 <= HN) tmem_ld_32x32b_x16(taddr + c, r[c / 16]);
        }
        if (HN % 16) tmem_ld_32x32b_x8(taddr + (HN / 16) * 16, r[HN / 16]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < HN; ++c) acc[c] += __uint_as_float(r[c / 16][c % 16]);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[buf]);
        if (++buf == kTmemBufs) { buf = 0; buf_phase ^= 1; }
      }
      const int row = m0 + q * 32 + lane;
      if (row < M) {
        float* drow = D + (size_t)ks * plane_stride + (size_t)row * ldd + n0 + hcol;
#pragma unroll
        for (int c = 0; c < HN; c += 8) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = F16 ? acc[c + e] * out_scale : acc[c + e];
          if (BIAS && ks == 0) {
            const float4 b0 = *reinterpret_cast<const float4*>(bias + n0 + hcol + c);
            const float4 b1 = *reinterpret_cast<const float4*>(bias + n0 + hcol + c + 4);
            o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w;
            o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
          }
          const int col0 = n0 + hcol + c;
          if (DENSE && (epi & 2)) {
#pragma unroll
            for (int e = 0; e < 8; ++e) if (col0 + e < ncols) o[e] += drow[c + e];
          }
          if (DENSE && (epi & 1)) {
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = fmaxf(o[e], 0.f);
          }
          if (!DENSE || ((epi & 4) && col0 + 8 <= ncols)) {
            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(drow + c), "f"(o[0]), "f"(o[1]), "f"(o[2]),
                         "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7])
                         : "memory");
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) if (col0 + e < ncols) drow[c + e] = o[e];
          }
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

// ---- forward blend, fp16 split, B panel resident --------------------------------------------------------------------
// With fp16 operands the tile-per-CTA kernel above is bound by L2 -> shared traffic: every 128 x 128 output tile pulls
// 128 KB of A and 128 KB of B through TMA (8.4 TB/s at 0.63 ms).  Here a CTA keeps one B panel (128 output columns, the
// whole K = 4 blocks of 64, hi and lo: 128 KB) in shared memory and streams kPanelRows consecutive A tiles past it, so a
// tile costs 128 KB + 128 KB / kPanelRows.  Work unit = (column panel, chunk of row tiles); units go to the CTAs
// round-robin.  Pipeline roles and the per-K-block register accumulation are those of the kernel above.
constexpr int kPanelKB = 4;          // K blocks of the forward (K = 217 <= 256 halfs)
template <int BN, int kPanelStages>   // output columns per panel, A stages (hi + lo = 32 KB each)
__global__ void __launch_bounds__(kThreads, 1)
blend_f16_panel_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                       const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                       float* __restrict__ D, int ldd, const float* __restrict__ bias, int M, int tiles_m, int tiles_n,
                       float out_scale, int kPanelRows) {   // kPanelRows: row tiles per unit, chosen by the launcher
  constexpr int kBKe = 2 * kBK;
  constexpr int kTile = kBM * kBK * 4;                 // one 128-row x 128-byte operand box: 16 KB
  constexpr int kBTile = BN * kBK * 4;                 // one BN-row box of B
  constexpr int kPanelBytes = kPanelKB * 2 * kBTile;   // B panel: [kb][hi, lo]
  constexpr int kStageBytes = 2 * kTile;               // A stage: hi, lo
  constexpr int kAccStride = 128;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stages = smem + kPanelBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(stages + kPanelStages * kStageBytes);
  uint64_t* empty = full + kPanelStages;
  uint64_t* tfull = empty + kPanelStages;
  uint64_t* tempty = tfull + kTmemBufs;
  uint64_t* bfull = tempty + kTmemBufs;                // B panel landed
  uint64_t* bempty = bfull + 1;                        // every MMA of the unit has read the panel
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bempty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks_m = (tiles_m + kPanelRows - 1) / kPanelRows;
  const int nunits = tiles_n * chunks_m;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kPanelStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < kTmemBufs; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
    mbar_init(bfull, 1); mbar_init(bempty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, bphase = 0;
      for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
        const int np = unit / chunks_m, mc = unit - np * chunks_m;
        const int n0 = np * BN;
        const int mt0 = mc * kPanelRows, mt1 = min(mt0 + kPanelRows, tiles_m);
        mbar_wait(bempty, bphase ^ 1);                 // the previous unit's MMAs are done with the panel
        mbar_expect_tx(bfull, kPanelBytes);
        for (int kb = 0; kb < kPanelKB; ++kb) {
          tma_load_2d(smem + (kb * 2) * kBTile, &tmBh, bfull, kb * kBKe, n0);
          tma_load_2d(smem + (kb * 2 + 1) * kBTile, &tmBl, bfull, kb * kBKe, n0);
        }
        bphase ^= 1;
        for (int mt = mt0; mt < mt1; ++mt) {
          for (int kb = 0; kb < kPanelKB; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* st = stages + stage * kStageBytes;
            mbar_expect_tx(&full[stage], kStageBytes);
            tma_load_2d(st, &tmAh, &full[stage], kb * kBKe, mt * kBM);
            tma_load_2d(st + kTile, &tmAl, &full[stage], kb * kBKe, mt * kBM);
            if (++stage == kPanelStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kBM, BN, true);
      int stage = 0, buf = 0;
      uint32_t phase = 0, buf_phase = 0, bphase = 0;
      for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
        const int mc = unit % chunks_m;
        const int mt0 = mc * kPanelRows, mt1 = min(mt0 + kPanelRows, tiles_m);
        mbar_wait(bfull, bphase);
        bphase ^= 1;
        for (int mt = mt0; mt < mt1; ++mt) {
          for (int kb = 0; kb < kPanelKB; ++kb) {
            mbar_wait(&tempty[buf], buf_phase ^ 1);
            mbar_wait(&full[stage], phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + (uint32_t)(buf * kAccStride);
            const uint32_t sa = smem_u32(stages + stage * kStageBytes);
            const uint32_t sb = smem_u32(smem + (kb * 2) * kBTile);
            const uint64_t dAh = make_smem_desc(sa), dAl = make_smem_desc(sa + kTile);
            const uint64_t dBh = make_smem_desc(sb), dBl = make_smem_desc(sb + kBTile);
#pragma unroll
            for (int k = 0; k < kBK / 8; ++k) {
              const uint64_t o = (uint64_t)(k * 2);
              umma<true>(tmem_d, dAl + o, dBh + o, idesc, k ? 1u : 0u);
              umma<true>(tmem_d, dAh + o, dBl + o, idesc, 1u);
            }
#pragma unroll
            for (int k = 0; k < kBK / 8; ++k) {
              const uint64_t o = (uint64_t)(k * 2);
              umma<true>(tmem_d, dAh + o, dBh + o, idesc, 1u);
            }
            umma_commit(&empty[stage]);
            umma_commit(&tfull[buf]);
            if (++stage == kPanelStages) { stage = 0; phase ^= 1; }
            if (++buf == kTmemBufs) { buf = 0; buf_phase ^= 1; }
          }
        }
        umma_commit(bempty);                           // arrives once every MMA above has completed
      }
    }
  } else {
    // ===== accumulate (fp32 registers, round-to-nearest) + epilogue: as in split3_gemm_kernel =====
    constexpr int HN = BN / 2;
    static_assert(HN % 16 == 0, "whole x16 TMEM loads");
    const int q = warp & 3;
    const int hcol = (warp - 2) >= 4 ? HN : 0;
    int buf = 0;
    uint32_t buf_phase = 0;
    for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
      const int np = unit / chunks_m, mc = unit - np * chunks_m;
      const int n0 = np * BN;
      const int mt0 = mc * kPanelRows, mt1 = min(mt0 + kPanelRows, tiles_m);
      for (int mt = mt0; mt < mt1; ++mt) {
        float acc[HN];
#pragma unroll
        for (int i = 0; i < HN; ++i) acc[i] = 0.f;
        for (int kb = 0; kb < kPanelKB; ++kb) {
          mbar_wait(&tfull[buf], buf_phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kAccStride) + (uint32_t)hcol;
          uint32_t r[HN / 16][16];
#pragma unroll
          for (int c = 0; c < HN; c += 16) tmem_ld_32x32b_x16(taddr + c, r[c / 16]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int c = 0; c < HN; ++c) acc[c] += __uint_as_float(r[c / 16][c % 16]);
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[buf]);
          if (++buf == kTmemBufs) { buf = 0; buf_phase ^= 1; }
        }
        const int row = mt * kBM + q * 32 + lane;
        if (row < M) {
          float* drow = D + (size_t)row * ldd + n0 + hcol;
#pragma unroll
          for (int c = 0; c < HN; c += 8) {
            const float4 b0 = *reinterpret_cast<const float4*>(bias + n0 + hcol + c);
            const float4 b1 = *reinterpret_cast<const float4*>(bias + n0 + hcol + c + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __fadd_rn(__fmul_rn(acc[c + e], out_scale), bb[e]);
            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(drow + c), "f"(o[0]), "f"(o[1]), "f"(o[2]),
                         "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7])
                         : "memory");
          }
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

cudaError_t load_encode() {
  if (g_encode) return cudaSuccess;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess) return e;
  if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return cudaSuccess;
}

// 2-D tensor [rows][cols] of fp32 (or fp16) with row pitch ld (elements); box = 128 bytes of columns x box_rows,
// SWIZZLE_128B, zero OOB fill (a K that is not a multiple of the box reads zeros past its end).
// The encoding is a pure function of its arguments, and a training loop presents the same few (pointer, shape) tuples
// step after step (torch's caching allocator hands the same blocks back): a small ring of recent encodings replaces
// the eight driver calls per step with eight 40-byte compares.
struct MapKey {
  const void* base; int rows, cols, ld, box_rows, f16;
  bool operator==(const MapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && f16 == o.f16;
  }
};
constexpr int kMapCache = 32;
std::mutex g_map_mutex;
MapKey g_map_keys[kMapCache];
CUtensorMap g_map_vals[kMapCache];
int g_map_used = 0, g_map_next = 0;

cudaError_t make_map(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_rows, bool f16) {
  const MapKey key{base, rows, cols, ld, box_rows, f16 ? 1 : 0};
  {
    std::lock_guard<std::mutex> lk(g_map_mutex);
    for (int i = 0; i < g_map_used; ++i)
      if (g_map_keys[i] == key) { *map = g_map_vals[i]; return cudaSuccess; }
  }
  const size_t es = f16 ? 2 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * es};
  cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
  std::lock_guard<std::mutex> lk(g_map_mutex);
  g_map_keys[g_map_next] = key;
  g_map_vals[g_map_next] = *map;
  g_map_next = (g_map_next + 1) % kMapCache;
  g_map_used = g_map_used < kMapCache ? g_map_used + 1 : kMapCache;
  return cudaSuccess;
}

template <int BN, bool BIAS, bool F16, bool DENSE = false>
cudaError_t launch_gemm(const void* Ah, const void* Al, int lda, const void* Bh, const void* Bl, int ldb, float* D,
                        int ldd, const float* bias, int M, int Ntot, int K, float out_scale, int num_sms, cudaStream_t st,
                        int ksplit = 1, size_t plane_stride = 0, int ncols = -1, int epi = 4) {
  cudaError_t e = load_encode();
  if (e != cudaSuccess) return e;
  CUtensorMap mAh, mAl, mBh, mBl;
  if ((e = make_map(&mAh, Ah, M, K, lda, kBM, F16)) != cudaSuccess) return e;
  if ((e = make_map(&mAl, Al, M, K, lda, kBM, F16)) != cudaSuccess) return e;
  if ((e = make_map(&mBh, Bh, Ntot, K, ldb, BN, F16)) != cudaSuccess) return e;
  if ((e = make_map(&mBl, Bl, Ntot, K, ldb, BN, F16)) != cudaSuccess) return e;
  constexpr int kStageBytes = 2 * kBM * kBK * 4 + 2 * BN * kBK * 4;
  const size_t smem = (size_t)kStages * kStageBytes + 256 + 1024;
  e = cudaFuncSetAttribute(split3_gemm_kernel<BN, BIAS, F16, DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int tiles_m = (M + kBM - 1) / kBM, tiles_n = Ntot / BN;
  const int grid = min(tiles_m * tiles_n * ksplit, num_sms);
  constexpr int kBKe = F16 ? 2 * kBK : kBK;
  split3_gemm_kernel<BN, BIAS, F16, DENSE><<<grid, kThreads, smem, st>>>(mAh, mAl, mBh, mBl, D, ldd, bias, M, (K + kBKe - 1) / kBKe,
                                                                   tiles_m, tiles_n, out_scale, ksplit, plane_stride,
                                                                   ncols < 0 ? Ntot : ncols, epi);
  return cudaGetLastError();
}

}  // namespace

// v_posed[N][LD] = X[N][224] * BT[LD][224]^T + v_template   (LD is a multiple of 768, hence of BN = 128)
cudaError_t launch_blend_fwd_tc(const SmplB200Model* m, const float* Xh, const float* Xl, int N, float* v_posed,
                                cudaStream_t st) {
  LaunchScope scope(KID_BLEND_FWD, st);
  if (m->BT16_hi) {   // fp16 split: Xh / Xl hold [N][kKPad] halfs (launch_pose_fwd with the fp16 tables)
    cudaError_t e = load_encode();
    if (e != cudaSuccess) return e;
    CUtensorMap mAh, mAl, mBh, mBl;
    if ((e = make_map(&mAh, Xh, N, kK, kKPad, kBM, true)) != cudaSuccess) return e;
    if ((e = make_map(&mAl, Xl, N, kK, kKPad, kBM, true)) != cudaSuccess) return e;
    // measured (N = 16384): <128, 3> with 16 rows per unit 0.55 ms; 2 stages 0.61; 8 rows 0.57; 32 rows 0.60; <96, 4> 0.58;
    // <64, 5> 0.66
    constexpr int kPanelBN = 128, kPanelStages = 3;
    if ((e = make_map(&mBh, m->BT16_hi, m->LD, kK, kKPad, kPanelBN, true)) != cudaSuccess) return e;
    if ((e = make_map(&mBl, m->BT16_lo, m->LD, kK, kKPad, kPanelBN, true)) != cudaSuccess) return e;
    const size_t smem = (size_t)kPanelKB * 2 * kPanelBN * 128 + (size_t)kPanelStages * 2 * kBM * 128 + 256 + 1024;
    e = cudaFuncSetAttribute(blend_f16_panel_kernel<kPanelBN, kPanelStages>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int tiles_m = (N + kBM - 1) / kBM, tiles_n = m->LD / kPanelBN;
    // Row tiles per unit: units go to the CTAs round-robin, so the launch takes ceil(units / SMs) rounds of (rows + the
    // unit's exposed panel load, ~0.85 of a row tile: 128 KB against 32 KB + the MMAs).  16 rows minimise that at 16384
    // samples (9 rounds of 16); a 2048-sample shard (16 row tiles) would run 2 rounds of 16 with 14 CTAs in the second.
    int rows = 16;
    {
      double best = 1e30;
      for (int r = 1; r <= 32; ++r) {
        const int chunks = (tiles_m + r - 1) / r;
        const int rr = (tiles_m + chunks - 1) / chunks;            // the even split with that many chunks
        const long long units = (long long)tiles_n * chunks;
        const double cost = (double)((units + m->num_sms - 1) / m->num_sms) * (rr + 0.85);
        if (cost < best - 1e-9) { best = cost; rows = rr; }
      }
    }
    const int nunits = tiles_n * ((tiles_m + rows - 1) / rows);
    blend_f16_panel_kernel<kPanelBN, kPanelStages><<<min(nunits, m->num_sms), kThreads, smem, st>>>(
        mAh, mAl, mBh, mBl, v_posed, m->LD, m->vt_pad, N, tiles_m, tiles_n, 1.0f / (kXScale16 * m->bt16_scale), rows);
    return cudaGetLastError();
  }
  return launch_gemm<128, true, false>(Xh, Xl, kKPad, m->BT_hi, m->BT_lo, kKPad, v_posed, m->LD, m->vt_pad, N, m->LD, kKPad,
                                       1.0f, m->num_sms, st);
}

// K slices of the backward blend product for batch N and depth Kp: 2 ceil(N / 128) output tiles of Kp / 32 K blocks each
// go to the CTAs round-robin; a slice costs its K blocks plus ~8 K-block times of pipeline fill, epilogue and plane sum
// (measured: 16384 samples 0.168 ms unsplit / 0.176 in 4 slices; 2048 samples 0.080 unsplit / 0.033 in 4 slices).
int blend_bwd_ksplit(const SmplB200Model* m, int N, int Kp) {
  const int tiles = 2 * ((N + kBM - 1) / kBM), kblocks = (Kp + kBK - 1) / kBK;
  int best_s = 1;
  double best = 1e30;
  for (int s = 1; s <= kMaxBlendBwdSplit; ++s) {
    const long long units = (long long)tiles * s;
    const double cost = (double)((units + m->num_sms - 1) / m->num_sms) * ((kblocks + s - 1) / s + 8.0);
    if (cost < best - 1e-9) { best = cost; best_s = s; }
  }
  return best_s;
}

// g_X[s][N][224] (s < ksplit planes, summed by pose_bwd) = g_vp[N][Kp] * Bs[224][Kp]^T   (Kp a multiple of 32)
cudaError_t launch_blend_bwd_tc(const SmplB200Model* m, const VsTables* t, const float* gvp_hi, const float* gvp_lo,
                                size_t gvp_ld, int N, float* g_X, int* planes, cudaStream_t st) {
  LaunchScope scope(KID_BLEND_BWD, st);
  const int ks = blend_bwd_ksplit(m, N, t->Kp);
  *planes = ks;
  return launch_gemm<kKPad / 2, false, false>(gvp_hi, gvp_lo, (int)gvp_ld, t->Bs_hi, t->Bs_lo, t->Kp, g_X, kKPad, nullptr,
                                              N, kKPad, t->Kp, 1.0f, m->num_sms, st, ks, (size_t)N * kKPad);
}

// D[M][ncols] (+)= A[M][K] * B[Np][K]^T (+ bias) (ReLU): the regressor's Dense products (model.py:63-105) as 3xTF32
// tcgen05 GEMMs.  A / B arrive already split (hi + lo, rows of ldk floats, K padded with zeros to a multiple of 4; B padded
// with zero rows to Np = a multiple of the column tile).  bias: Np floats or null.
cudaError_t launch_dense_gemm(const float* Ah, const float* Al, const float* Bh, const float* Bl, int ldk, float* D, int ldd,
                              const float* bias, int M, int Np, int ncols, int K, bool relu, bool accumulate, int num_sms,
                              cudaStream_t st) {
  const bool vec = (ldd % 8 == 0) && (reinterpret_cast<uintptr_t>(D) % 32 == 0);
  const int epi = (relu ? 1 : 0) | (accumulate ? 2 : 0) | (vec ? 4 : 0);
  if (Np % 128 == 0) {
    if (bias) return launch_gemm<128, true, false, true>(Ah, Al, ldk, Bh, Bl, ldk, D, ldd, bias, M, Np, K, 1.0f, num_sms, st, 1, 0, ncols, epi);
    return launch_gemm<128, false, false, true>(Ah, Al, ldk, Bh, Bl, ldk, D, ldd, nullptr, M, Np, K, 1.0f, num_sms, st, 1, 0, ncols, epi);
  }
  if (Np % 96 == 0) {
    if (bias) return launch_gemm<96, true, false, true>(Ah, Al, ldk, Bh, Bl, ldk, D, ldd, bias, M, Np, K, 1.0f, num_sms, st, 1, 0, ncols, epi);
    return launch_gemm<96, false, false, true>(Ah, Al, ldk, Bh, Bl, ldk, D, ldd, nullptr, M, Np, K, 1.0f, num_sms, st, 1, 0, ncols, epi);
  }
  return cudaErrorInvalidValue;
}

}  // namespace smplb200
