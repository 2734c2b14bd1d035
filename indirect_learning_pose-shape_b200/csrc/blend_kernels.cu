// K1: shape + pose blend products.
//   forward  (batch_smpl.py:106-108,126-128):  v_posed[N][LD] = X[N][224] * Bm[224][LD] + v_template
//            with X = [beta(10) | pose_feature(207) | 0(7)], Bm = [shapedirs; posedirs; 0] (one GEMM for both blends)
//   backward (TF autodiff of the same lines):  g_X[N][224]   = g_vp[N][Kp] * BmT[Kp][224]
//
// This file holds the skinny (small-batch, bandwidth-bound) fp32 kernels used when the batch does not make the
// product dense (N < kDenseBatch); from kDenseBatch on both products run on the tensor cores (tc_gemm.cu).
#include "common.cuh"

namespace smplb200 {

namespace {

// ---- skinny forward: up to kRows samples per block, one column quad per thread; streams Bm once per block ---
// Block = 32 column quads (128 columns) x 8 k-groups; group g owns blend rows g, g+8, ...; the 8 partial sums meet in
// shared memory.  162 blocks x 256 threads keep ~2.6 MB of Bm loads in flight, which is what a batch-1 call needs.
constexpr int kSkinnyRows = 8;
constexpr int kSkinnyGroups = 8;
__global__ void __launch_bounds__(256)
blend_fwd_skinny_kernel(const float* __restrict__ X, const float* __restrict__ Bm, const float* __restrict__ vt,
                        float* __restrict__ VP, int N, int LD) {
  __shared__ float xs[kSkinnyRows][kKPad];
  __shared__ __align__(16) float red[kSkinnyGroups][kSkinnyRows][128];
  const int n0 = blockIdx.y * kSkinnyRows;
  const int rows = min(kSkinnyRows, N - n0);
  for (int i = threadIdx.x; i < kSkinnyRows * kKPad; i += blockDim.x) {
    const int r = i / kKPad, k = i % kKPad;
    xs[r][k] = (r < rows) ? X[(size_t)(n0 + r) * kKPad + k] : 0.f;
  }
  __syncthreads();
  const int q = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + q * 4;        // LD is a multiple of 128
  float4 acc[kSkinnyRows];
#pragma unroll
  for (int r = 0; r < kSkinnyRows; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int k = g; k < kK; k += kSkinnyGroups) {
    const float4 b = *reinterpret_cast<const float4*>(Bm + (size_t)k * LD + c);
#pragma unroll
    for (int r = 0; r < kSkinnyRows; ++r) {
      const float x = xs[r][k];
      acc[r].x = fmaf(x, b.x, acc[r].x); acc[r].y = fmaf(x, b.y, acc[r].y);
      acc[r].z = fmaf(x, b.z, acc[r].z); acc[r].w = fmaf(x, b.w, acc[r].w);
    }
  }
#pragma unroll
  for (int r = 0; r < kSkinnyRows; ++r) *reinterpret_cast<float4*>(&red[g][r][q * 4]) = acc[r];
  __syncthreads();
  // thread (g, q) finishes row g: sum the 8 groups in fixed order, add the template, store
  if (g < rows) {
    float4 s = *reinterpret_cast<const float4*>(vt + c);
#pragma unroll
    for (int h = 0; h < kSkinnyGroups; ++h) {
      const float4 p = *reinterpret_cast<const float4*>(&red[h][g][q * 4]);
      s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
    }
    *reinterpret_cast<float4*>(VP + (size_t)(n0 + g) * LD + c) = s;
  }
}

// ---- skinny backward: split-K over column chunks, partial sums combined with atomicAdd into a zeroed g_X ---------
constexpr int kBwdChunk = 128;   // columns (K of the backward product) per block
__global__ void __launch_bounds__(kKPad)
blend_bwd_skinny_kernel(const float* __restrict__ gvp, int gvp_ld, const float* __restrict__ BmT, int Kp,
                        float* __restrict__ gX, int N) {
  __shared__ float gs[kSkinnyRows][kBwdChunk];
  const int n0 = blockIdx.y * kSkinnyRows;
  const int rows = min(kSkinnyRows, N - n0);
  const int c0 = blockIdx.x * kBwdChunk;
  const int cn = min(kBwdChunk, Kp - c0);
  for (int i = threadIdx.x; i < kSkinnyRows * kBwdChunk; i += blockDim.x) {
    const int r = i / kBwdChunk, c = i % kBwdChunk;
    gs[r][c] = (r < rows && c < cn) ? gvp[(size_t)(n0 + r) * gvp_ld + c0 + c] : 0.f;
  }
  __syncthreads();
  const int k = threadIdx.x;      // 224 threads, one blend row each
  float acc[kSkinnyRows];
#pragma unroll
  for (int r = 0; r < kSkinnyRows; ++r) acc[r] = 0.f;
  for (int c = 0; c < cn; ++c) {
    const float b = BmT[(size_t)(c0 + c) * kKPad + k];
#pragma unroll
    for (int r = 0; r < kSkinnyRows; ++r) acc[r] = fmaf(gs[r][c], b, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < kSkinnyRows; ++r)
    if (r < rows) atomicAdd(gX + (size_t)(n0 + r) * kKPad + k, acc[r]);
}

}  // namespace


cudaError_t launch_blend_fwd(const SmplB200Model* m, const float* X, int N, float* v_posed, cudaStream_t st) {
  LaunchScope scope(KID_BLEND_FWD, st);
  dim3 grid(m->LD / 128, (N + kSkinnyRows - 1) / kSkinnyRows);
  blend_fwd_skinny_kernel<<<grid, 256, 0, st>>>(X, m->Bm, m->vt_pad, v_posed, N, m->LD);
  return cudaGetLastError();
}

cudaError_t launch_blend_bwd(const SmplB200Model* m, const VsTables* t, const float* g_vp, size_t gvp_ld, int N,
                             float* g_X, cudaStream_t st) {
  (void)m;
  LaunchScope scope(KID_BLEND_BWD, st);
  cudaError_t e = cudaMemsetAsync(g_X, 0, sizeof(float) * (size_t)N * kKPad, st);
  if (e != cudaSuccess) return e;
  dim3 grid((t->Kp + kBwdChunk - 1) / kBwdChunk, (N + kSkinnyRows - 1) / kSkinnyRows);
  blend_bwd_skinny_kernel<<<grid, kKPad, 0, st>>>(g_vp, (int)gvp_ld, t->BmT, t->Kp, g_X, N);
  return cudaGetLastError();
}

}  // namespace smplb200
