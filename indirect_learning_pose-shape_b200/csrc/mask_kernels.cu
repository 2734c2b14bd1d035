// K4: per-sample point z-buffer on the reference's hard-coded 64x64 grid -> per-vertex weight 1 (front-most at
// its pixel) or 500.  No gradient (both map_fn calls have back_prop=False).
//
// Reference semantics (keras_smpl/compute_mask.py:12-108):
//   :22      pixel = tf.round(u, v)                 -> round half to even (rintf)
//   :44      img_wh = 64, independent of the output resolution
//   :90-92   a vertex belongs to pixel (c, r) iff its rounded u == c and rounded v == r (c, r in 0..63)
//   :100-102 winner = tf.argmax(depth)              -> LARGEST z, lowest vertex index on ties
//   :98-99   an empty pixel yields tf.ones([1,1,4]) -> "winner" index 1
//   :66-70   mask = 500 everywhere, 1 at the unique winner indices
// One block per sample; the z-buffer lives in shared memory (two 64x64 word planes).
#include "common.cuh"

namespace smplb200 {

namespace {

constexpr int kCells = kMaskGrid * kMaskGrid;

// monotone map float -> uint (larger float <=> larger uint); -0 and +0 must compare equal, so canonicalise first
__device__ __forceinline__ unsigned int orderable(float z) {
  z += 0.0f;                                   // -0 -> +0
  const unsigned int b = __float_as_uint(z);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ int cell_of(float u, float v) {
  const float pu = rintf(u), pv = rintf(v);    // tf.round: half to even
  if (!(pu >= 0.f && pu <= (float)(kMaskGrid - 1) && pv >= 0.f && pv <= (float)(kMaskGrid - 1))) return -1;
  return (int)pv * kMaskGrid + (int)pu;        // column = u, row = v (meshgrid 'xy', :49-54)
}

// kPer = vertices per thread kept in registers between the passes (cell and depth key are computed once); launches whose
// Vs exceeds kPer * blockDim fall back to re-reading (kPer = 0).
template <int kPer>
__global__ void __launch_bounds__(1024)
mask_kernel(const float* __restrict__ projects, int N, int Vs, float* __restrict__ mask) {
  __shared__ unsigned int zmax[kCells];        // 0 = empty (orderable() never returns 0 for a non-NaN float)
  __shared__ unsigned int win[kCells];
  __shared__ int occupied;
  const int n = blockIdx.x, tid = threadIdx.x;
  const float* p = projects + (size_t)n * Vs * 3;
  for (int i = tid; i < kCells; i += blockDim.x) { zmax[i] = 0u; win[i] = 0xffffffffu; }
  if (tid == 0) occupied = 0;
  int cellr[kPer > 0 ? kPer : 1];
  unsigned keyr[kPer > 0 ? kPer : 1];
  if (kPer > 0) {
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int i = tid + k * blockDim.x;
      cellr[k] = -1; keyr[k] = 0u;
      if (i < Vs) {
        const float z = p[i * 3 + 2];
        const int c = cell_of(p[i * 3], p[i * 3 + 1]);
        if (c >= 0 && z == z) { cellr[k] = c; keyr[k] = orderable(z); }
      }
    }
  }
  __syncthreads();
  if (kPer > 0) {
#pragma unroll
    for (int k = 0; k < kPer; ++k)
      if (cellr[k] >= 0) atomicMax(&zmax[cellr[k]], keyr[k]);
  } else {
    for (int i = tid; i < Vs; i += blockDim.x) {
      const float z = p[i * 3 + 2];
      const int c = cell_of(p[i * 3], p[i * 3 + 1]);
      if (c >= 0 && z == z) atomicMax(&zmax[c], orderable(z));
    }
  }
  __syncthreads();
  if (kPer > 0) {
#pragma unroll
    for (int k = 0; k < kPer; ++k)
      if (cellr[k] >= 0 && keyr[k] == zmax[cellr[k]]) atomicMin(&win[cellr[k]], (unsigned int)(tid + k * blockDim.x));
  } else {
    for (int i = tid; i < Vs; i += blockDim.x) {
      const float z = p[i * 3 + 2];
      const int c = cell_of(p[i * 3], p[i * 3 + 1]);
      if (c >= 0 && z == z && orderable(z) == zmax[c]) atomicMin(&win[c], (unsigned int)i);
    }
  }
  int occ = 0;
  for (int i = tid; i < kCells; i += blockDim.x) occ += zmax[i] != 0u;
  occ = __reduce_add_sync(0xffffffffu, occ);
  if ((tid & 31) == 0 && occ) atomicAdd(&occupied, occ);
  __syncthreads();
  const bool any_empty = occupied < kCells;
  if (kPer > 0) {
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int i = tid + k * blockDim.x;
      if (i < Vs) {
        bool vis = (cellr[k] >= 0) && (win[cellr[k]] == (unsigned int)i);
        if (i == 1 && any_empty) vis = true;     // every empty pixel votes for index 1 (:98-99)
        mask[(size_t)n * Vs + i] = vis ? 1.0f : kMaskInvisible;
      }
    }
  } else {
    for (int i = tid; i < Vs; i += blockDim.x) {
      const int c = cell_of(p[i * 3], p[i * 3 + 1]);
      const float z = p[i * 3 + 2];
      bool vis = (c >= 0) && (z == z) && (win[c] == (unsigned int)i);
      if (i == 1 && any_empty) vis = true;       // every empty pixel votes for index 1 (:98-99)
      mask[(size_t)n * Vs + i] = vis ? 1.0f : kMaskInvisible;
    }
  }
}

}  // namespace

cudaError_t launch_mask_fwd(const float* projects, int N, int Vs, float* mask, cudaStream_t st) {
  LaunchScope scope(KID_MASK, st);
  const int threads = N < 64 ? 1024 : 256;    // few samples: spend threads on each
  if (Vs <= 7 * threads) mask_kernel<7><<<N, threads, 0, st>>>(projects, N, Vs, mask);
  else mask_kernel<0><<<N, threads, 0, st>>>(projects, N, Vs, mask);
  return cudaGetLastError();
}

}  // namespace smplb200
