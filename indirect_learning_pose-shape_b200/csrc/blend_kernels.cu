// K1: shape + pose blend products.
//   forward  (batch_smpl.py:106-108,126-128):  v_posed[N][LD] = X[N][224] * Bm[224][LD] + v_template
//            with X = [beta(10) | pose_feature(207) | 0(7)], Bm = [shapedirs; posedirs; 0] (one GEMM for both blends)
//   backward (TF autodiff of the same lines):  g_X[N][224]   = g_vp[N][Kp] * BmT[Kp][224]
//
// This file holds the fp32 CUDA-core tiled GEMM (128x128x16 tiles, 8x8 register micro-tiles, double-buffered
// shared memory) and the skinny (small-batch, bandwidth-bound) kernels used when the batch does not make the
// product dense.
#include "common.cuh"

namespace smplb200 {

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int kThreads = 256;

// C[M][ldc] = A[M][lda] * B[K][ldb] (+ bias[col]) ; K % 16 == 0, Ncols % 4 == 0, all row strides % 4 == 0.
template <bool kBias>
__global__ void __launch_bounds__(kThreads)
sgemm_128x128_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                     float* __restrict__ C, int ldc, const float* __restrict__ bias, int M, int Ncols, int K) {
  __shared__ __align__(16) float As[2][BK][BM + 4];   // transposed: As[k][m]
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = tid & 15, ty = tid >> 4;              // 16 x 16 threads, each 8 (m) x 8 (n)

  // global->smem mapping. A tile: 128 rows x 16 k = 512 float4 -> 2 per thread. B tile: 16 k x 128 n = 512 float4.
  const int a_row = tid >> 2, a_k4 = (tid & 3) * 4;    // rows a_row and a_row + 64
  const int b_k = tid >> 5, b_n4 = (tid & 31) * 4;     // k rows b_k and b_k + 8
  float4 ra[2], rb[2];

  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = m0 + a_row + i * 64;
      ra[i] = (r < M) ? *reinterpret_cast<const float4*>(A + (size_t)r * lda + k0 + a_k4) : make_float4(0.f, 0.f, 0.f, 0.f);
      const int c = n0 + b_n4;
      rb[i] = (c < Ncols) ? *reinterpret_cast<const float4*>(B + (size_t)(k0 + b_k + i * 8) * ldb + c)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = a_row + i * 64;
      As[buf][a_k4 + 0][r] = ra[i].x; As[buf][a_k4 + 1][r] = ra[i].y;
      As[buf][a_k4 + 2][r] = ra[i].z; As[buf][a_k4 + 3][r] = ra[i].w;
      *reinterpret_cast<float4*>(&Bs[buf][b_k + i * 8][b_n4]) = rb[i];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  gload(0);
  sstore(0);
  __syncthreads();
  const int nk = K / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      // rows ty*4..+3 and 64+ty*4..+3 ; cols tx*4..+3 and 64+tx*4..+3 (conflict-free float4 reads)
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = n0 + h * 64 + tx * 4;
      if (c >= Ncols) continue;
      float4 v = make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
      if (kBias) {
        const float4 bb = *reinterpret_cast<const float4*>(bias + c);
        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
      }
      *reinterpret_cast<float4*>(C + (size_t)r * ldc + c) = v;
    }
  }
}

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

constexpr int kDenseBatch = 64;   // below this the blend is a bandwidth-bound skinny product

cudaError_t launch_blend_fwd(const SmplB200Model* m, const float* X, int N, float* v_posed, cudaStream_t st) {
  LaunchScope scope(KID_BLEND_FWD, st);
  if (N < kDenseBatch) {
    dim3 grid(m->LD / 128, (N + kSkinnyRows - 1) / kSkinnyRows);
    blend_fwd_skinny_kernel<<<grid, 256, 0, st>>>(X, m->Bm, m->vt_pad, v_posed, N, m->LD);
  } else {
    dim3 grid(m->LD / BN, (N + BM - 1) / BM);
    sgemm_128x128_kernel<true><<<grid, kThreads, 0, st>>>(X, kKPad, m->Bm, m->LD, v_posed, m->LD, m->vt_pad, N, m->LD,
                                                         kKPad);
  }
  return cudaGetLastError();
}

cudaError_t launch_blend_bwd(const SmplB200Model* m, const VsTables* t, const float* g_vp, size_t gvp_ld, int N,
                             float* g_X, cudaStream_t st) {
  (void)m;
  LaunchScope scope(KID_BLEND_BWD, st);
  if (N < kDenseBatch) {
    cudaError_t e = cudaMemsetAsync(g_X, 0, sizeof(float) * (size_t)N * kKPad, st);
    if (e != cudaSuccess) return e;
    dim3 grid((t->Kp + kBwdChunk - 1) / kBwdChunk, (N + kSkinnyRows - 1) / kSkinnyRows);
    blend_bwd_skinny_kernel<<<grid, kKPad, 0, st>>>(g_vp, (int)gvp_ld, t->BmT, t->Kp, g_X, N);
  } else {
    dim3 grid((kKPad + BN - 1) / BN, (N + BM - 1) / BM);
    sgemm_128x128_kernel<false><<<grid, kThreads, 0, st>>>(g_vp, (int)gvp_ld, t->BmT, kKPad, g_X, kKPad, nullptr, N, kKPad,
                                                          t->Kp);
  }
  return cudaGetLastError();
}

}  // namespace smplb200
