// The regression module ahead of the decoder: Dense layers of model.py:63-105 -- the IEF loop
// (three passes through shared Dense 1024 / 1024 / 86, :63-97) and the plain regressor (Dense 2048 / 1024 / 86, :99-105).
//
// Reference arithmetic per layer (Keras Dense, fp32):  Y = act(X W + b), W = kernel (in, out), act = relu / linear.
// TF autodiff:  gZ = gY * [Y > 0] (relu) ;  gX = gZ W^T ;  gW = X^T gZ ;  gb = sum_rows gZ.
//
// All three products run on the tensor cores as 3xTF32 split GEMMs (csrc/tc_gemm.cu: hi = x with 13 low mantissa bits
// cleared, lo = x - hi, lo*hi + hi*lo + hi*hi with per-K-block fp32 register accumulation), i.e. at fp32 accuracy like
// the reference's sgemm.  The kernels here prepare the K-major split operands (with the transposes the products need)
// and do the small elementwise work around them.
#include <algorithm>
#include "common.cuh"

namespace smplb200 {

namespace {

// src [rows][cols] (row stride ld) -> hi/lo [rows_p][ldk]: exact TF32 split, zero padding beyond (rows, cols)
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ src, int ld, int rows, int cols, float* __restrict__ hi, float* __restrict__ lo,
             int ldk, int rows_p, const float* __restrict__ gate, int ldg) {
  const long long total = (long long)rows_p * ldk;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ldk), c = (int)(i - (long long)r * ldk);
    float x = (r < rows && c < cols) ? src[(size_t)r * ld + c] : 0.f;
    if (gate && r < rows && c < cols && !(gate[(size_t)r * ldg + c] > 0.f)) x = 0.f;      // relu backward: gZ = gY [Y > 0]
    const float h = tf32_hi(x);
    hi[i] = h; lo[i] = x - h;
  }
}

// src [rows][cols] -> hi/lo [cols_p][ldk] = split(src^T), ldk >= rows; 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256)
transpose_split_kernel(const float* __restrict__ src, int ld, int rows, int cols, float* __restrict__ hi,
                       float* __restrict__ lo, int ldk, int cols_p, const float* __restrict__ gate, int ldg) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    float x = (r < rows && c < cols) ? src[(size_t)r * ld + c] : 0.f;
    if (gate && r < rows && c < cols && !(gate[(size_t)r * ldg + c] > 0.f)) x = 0.f;
    tile[j][tx] = x;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;                            // output row = source column
    if (c < cols_p && r < ldk) {
      const float x = tile[tx][j];
      const float h = tf32_hi(x);
      hi[(size_t)c * ldk + r] = h; lo[(size_t)c * ldk + r] = x - h;
    }
  }
}

// gb[c] (+)= sum_r gY[r][c] [Y[r][c] > 0]   (one block per 32 columns, fixed-order tree: deterministic)
__global__ void __launch_bounds__(256)
bias_grad_kernel(const float* __restrict__ gY, int ldg, const float* __restrict__ Y, int ldy, int rows, int cols,
                 float* __restrict__ gb, int accumulate) {
  __shared__ float part[8][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
  float s = 0.f;
  if (c < cols)
    for (int r = ty; r < rows; r += 8) {
      const float g = gY[(size_t)r * ldg + c];
      s += (!Y || Y[(size_t)r * ldy + c] > 0.f) ? g : 0.f;
    }
  part[ty][threadIdx.x & 31] = s;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x & 31];
    gb[c] = accumulate ? gb[c] + t : t;
  }
}

// out[r][c] = a[r][c] + scale * d[r][c]   (param_{k+1} = param_k + scaledown * delta_k, model.py:80-82; also the IEF state
// assembly: columns copied into the [features | params] state row)
__global__ void __launch_bounds__(256)
axpy_cols_kernel(const float* __restrict__ a, int lda, const float* __restrict__ d, int ldd, float scale, int rows, int cols,
                 float* __restrict__ out, int ldo) {
  const long long total = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    const float av = a ? a[(size_t)r * lda + c] : 0.f;
    out[(size_t)r * ldo + c] = d ? fmaf(scale, d[(size_t)r * ldd + c], av) : av;
  }
}

inline int ru_i(int x, int a) { return (x + a - 1) / a * a; }
inline size_t ru_z(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline unsigned grid_for(long long total) { return (unsigned)std::min<long long>((total + 255) / 256, 148 * 16); }

}  // namespace

int dense_np(int n) { return n >= 128 ? ru_i(n, 128) : ru_i(n, 96); }      // column tile: 128, or one 96-wide tile for narrow outputs

// workspace of one GEMM with contraction depth K, M rows and N columns: the four split operands (+ padded bias)
size_t dense_gemm_ws(int M, int N, int K) {
  const int ldk = ru_i(K, 4), Np = dense_np(N);
  return 2 * ru_z((size_t)ru_i(M, 128) * ldk * 4, 256) + 2 * ru_z((size_t)Np * ldk * 4, 256) + ru_z((size_t)Np * 4, 256);
}

// Y[M][out] = act(X[M][in] W[in][out] + b)
cudaError_t launch_dense_fwd(const float* X, int ldx, const float* W, const float* b, int M, int in, int out, bool relu,
                             float* Y, int ldy, void* ws, int num_sms, cudaStream_t st) {
  const int ldk = ru_i(in, 4), Np = dense_np(out), Mp = ru_i(M, 128);
  char* p = (char*)ws;
  float* Ah = (float*)p; p += ru_z((size_t)Mp * ldk * 4, 256);
  float* Al = (float*)p; p += ru_z((size_t)Mp * ldk * 4, 256);
  float* Bh = (float*)p; p += ru_z((size_t)Np * ldk * 4, 256);
  float* Bl = (float*)p; p += ru_z((size_t)Np * ldk * 4, 256);
  float* bp = (float*)p;
  LaunchScope scope(KID_DENSE, st);
  split_kernel<<<grid_for((long long)Mp * ldk), 256, 0, st>>>(X, ldx, M, in, Ah, Al, ldk, Mp, nullptr, 0);
  transpose_split_kernel<<<dim3((Np + 31) / 32, (ldk + 31) / 32), 256, 0, st>>>(W, out, in, out, Bh, Bl, ldk, Np, nullptr, 0);
  if (b) axpy_cols_kernel<<<1, 256, 0, st>>>(b, 0, nullptr, 0, 0.f, 1, out, bp, 0);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (b && Np > out) {
    e = cudaMemsetAsync(bp + out, 0, (size_t)(Np - out) * 4, st);
    if (e != cudaSuccess) return e;
  }
  return launch_dense_gemm(Ah, Al, Bh, Bl, ldk, Y, ldy, b ? bp : nullptr, M, Np, out, ldk, relu, false, num_sms, st);
}

// gX[M][in] = gZ W^T (nullable), gW[in][out] (+)= X^T gZ, gb[out] (+)= colsum gZ, with gZ = gY [Y > 0] when relu
cudaError_t launch_dense_bwd(const float* X, int ldx, const float* W, const float* Y, int ldy, const float* gY, int ldg, int M,
                             int in, int out, bool relu, float* gX, int ldgx, float* gW, float* gb, bool accumulate, void* ws,
                             int num_sms, cudaStream_t st) {
  const float* gate = relu ? Y : nullptr;
  LaunchScope scope(KID_DENSE, st);
  cudaError_t e;
  if (gb) bias_grad_kernel<<<(out + 31) / 32, 256, 0, st>>>(gY, ldg, gate, ldy, M, out, gb, accumulate ? 1 : 0);
  if (gX) {       // A = gZ [M][out] (depth out), B = W [in][out] (rows = in: already K-major)
    const int ldk = ru_i(out, 4), Np = dense_np(in), Mp = ru_i(M, 128);
    char* p = (char*)ws;
    float* Ah = (float*)p; p += ru_z((size_t)Mp * ldk * 4, 256);
    float* Al = (float*)p; p += ru_z((size_t)Mp * ldk * 4, 256);
    float* Bh = (float*)p; p += ru_z((size_t)Np * ldk * 4, 256);
    float* Bl = (float*)p;
    split_kernel<<<grid_for((long long)Mp * ldk), 256, 0, st>>>(gY, ldg, M, out, Ah, Al, ldk, Mp, gate, ldy);
    split_kernel<<<grid_for((long long)Np * ldk), 256, 0, st>>>(W, out, in, out, Bh, Bl, ldk, Np, nullptr, 0);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    e = launch_dense_gemm(Ah, Al, Bh, Bl, ldk, gX, ldgx, nullptr, M, Np, in, ldk, false, false, num_sms, st);
    if (e != cudaSuccess) return e;
  }
  if (gW) {       // A = X^T [in][M] (depth M), B = gZ^T [out][M]
    const int ldk = ru_i(M, 4), Np = dense_np(out), Mp = ru_i(in, 128);
    char* p = (char*)ws;
    float* Ah = (float*)p; p += ru_z((size_t)Mp * ldk * 4, 256);
    float* Al = (float*)p; p += ru_z((size_t)Mp * ldk * 4, 256);
    float* Bh = (float*)p; p += ru_z((size_t)Np * ldk * 4, 256);
    float* Bl = (float*)p;
    transpose_split_kernel<<<dim3((Mp + 31) / 32, (ldk + 31) / 32), 256, 0, st>>>(X, ldx, M, in, Ah, Al, ldk, Mp, nullptr, 0);
    transpose_split_kernel<<<dim3((Np + 31) / 32, (ldk + 31) / 32), 256, 0, st>>>(gY, ldg, M, out, Bh, Bl, ldk, Np, gate, ldy);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    e = launch_dense_gemm(Ah, Al, Bh, Bl, ldk, gW, out, nullptr, in, Np, out, ldk, false, accumulate, num_sms, st);
    if (e != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

cudaError_t launch_axpy_cols(const float* a, int lda, const float* d, int ldd, float scale, int rows, int cols, float* out,
                             int ldo, cudaStream_t st) {
  LaunchScope scope(KID_DENSE, st);
  axpy_cols_kernel<<<grid_for((long long)rows * cols), 256, 0, st>>>(a, lda, d, ldd, scale, rows, cols, out, ldo);
  return cudaGetLastError();
}

}  // namespace smplb200
