// K6: soft silhouette from projected vertices, forward and backward.
//
// Reference arithmetic (keras_smpl/projects_to_silhouette.py:14-44), per pixel g = (column c, row r):
//   s[g] = max_i exp(-||p_i - g||_2 / 1.2)     :35-38   (all vertices, no visibility weights; true division by 1.2f)
//   out[n, wh-1-r, c, :] = [1 - s, s]           :40-42   (rows flipped)
// max_i exp(-d_i/1.2) = exp(-(min_i d_i)/1.2): an exact nearest-vertex query.  The reference evaluates all wh^2 x V
// pairs; here every sample is counting-sorted into <= 32x32 grid cells in shared memory and each warp resolves an
// 8x4 pixel tile by visiting cell rings of growing Chebyshev radius until the ring's lower bound (rho-1)*B exceeds the
// worst current distance in the tile.  Lower bounds stay valid for vertices clamped into border cells, so the result
// is the exact minimum, bit-identical to the brute-force value.
// Backward (TF autodiff): d s/d p_i* = -(1/1.2) s (p_i* - g)/d at the first arg-min (ties: measure zero), 0 when d == 0
// (TF: NaN); upstream is g[...,1] - g[...,0].  Lanes that share an arg-min are merged with __match_any_sync before
// the shared-memory atomicAdd.
#include <math_constants.h>
#include "common.cuh"

namespace smplb200 {

namespace {

constexpr int kTileW = 8, kTileH = 4;
constexpr int kMaxGrid = 32;

struct SilSmem {
  float2* suv;    // [Vs] vertices sorted by cell
  int* svid;      // [Vs] their vertex ids
  int* start;     // [cells + 1]
  int* cursor;    // [cells]
  float* gacc;    // [Vs][2] (backward only)
};

__device__ __forceinline__ int cell_coord(float x, float invB, int G) {
  const float f = floorf(x * invB);
  return (int)fminf(fmaxf(f, 0.f), (float)(G - 1));     // NaN -> 0
}

__device__ void bin_vertices(const SilSmem& sm, const float* __restrict__ proj, int Vs, int B, int G) {
  const int cells = G * G;
  const float invB = 1.0f / (float)B;                    // B is a power of two or small integer; used for binning only
  for (int i = threadIdx.x; i <= cells; i += blockDim.x) sm.start[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < Vs; i += blockDim.x) {
    const int c = cell_coord(proj[i * 3 + 1], invB, G) * G + cell_coord(proj[i * 3], invB, G);
    atomicAdd(&sm.start[c + 1], 1);
  }
  __syncthreads();
  if (threadIdx.x < 32) {                                // warp scan over <= 1024 counts
    int carry = 0;
    for (int base = 0; base < cells; base += 32) {
      const int i = base + threadIdx.x;
      int v = (i < cells) ? sm.start[i + 1] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if ((int)threadIdx.x >= o) v += t;
      }
      if (i < cells) sm.start[i + 1] = v + carry;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cells; i += blockDim.x) sm.cursor[i] = sm.start[i];
  __syncthreads();
  for (int i = threadIdx.x; i < Vs; i += blockDim.x) {
    const float u = proj[i * 3], v = proj[i * 3 + 1];
    const int c = cell_coord(v, invB, G) * G + cell_coord(u, invB, G);
    const int pos = atomicAdd(&sm.cursor[c], 1);
    sm.suv[pos] = make_float2(u, v);
    sm.svid[pos] = i;
  }
  __syncthreads();
}

// Nearest vertex (squared distance, sorted position) for this lane's pixel; the whole warp searches together.
__device__ __forceinline__ void tile_search(const SilSmem& sm, int B, int G, int cx0, int cx1, int cy0, int cy1,
                                            bool active, float gx, float gy, float& best, int& barg) {
  const int lane = threadIdx.x & 31;
  best = CUDART_INF_F;
  barg = -1;
  for (int rho = 0; rho <= G; ++rho) {
    const int x0 = cx0 - rho, x1 = cx1 + rho, y0 = cy0 - rho, y1 = cy1 + rho;
    const int W = x1 - x0 + 1, H = y1 - y0 + 1;
    const int count = (rho == 0) ? W * H : 2 * W + 2 * (H - 2);
    for (int base = 0; base < count; base += 32) {
      const int t = base + lane;
      int x = -1, y = -1;
      if (t < count) {
        if (rho == 0) { y = y0 + t / W; x = x0 + t % W; }
        else if (t < W) { y = y0; x = x0 + t; }
        else if (t < 2 * W) { y = y1; x = x0 + (t - W); }
        else { const int q = t - 2 * W; y = y0 + 1 + (q >> 1); x = (q & 1) ? x1 : x0; }
      }
      const bool inr = x >= 0 && x < G && y >= 0 && y < G;
      int s0 = 0, s1 = 0;
      if (inr) { s0 = sm.start[y * G + x]; s1 = sm.start[y * G + x + 1]; }
      unsigned todo = __ballot_sync(0xffffffffu, s1 > s0);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int a = __shfl_sync(0xffffffffu, s0, src), b = __shfl_sync(0xffffffffu, s1, src);
        for (int i = a; i < b; ++i) {
          const float2 p = sm.suv[i];                    // broadcast
          const float du = __fsub_rn(p.x, gx), dv = __fsub_rn(p.y, gy);
          const float d2 = __fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv));
          if (d2 < best) { best = d2; barg = i; }
        }
      }
    }
    // every vertex in ring rho+1 or beyond is at least rho*B away from every pixel of the tile
    float worst = active ? best : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    const float lb = (float)(rho * B);
    if (lb * lb > worst) break;
    if (x0 <= 0 && y0 <= 0 && x1 >= G - 1 && y1 >= G - 1) break;   // the whole grid has been visited
  }
}

__device__ __forceinline__ SilSmem carve_sil(unsigned char* raw, int Vs, int cells, bool bwd) {
  SilSmem sm;
  size_t off = 0;
  sm.suv = reinterpret_cast<float2*>(raw + off); off += (size_t)Vs * 8;
  sm.svid = reinterpret_cast<int*>(raw + off); off += (size_t)Vs * 4;
  sm.start = reinterpret_cast<int*>(raw + off); off += (size_t)(cells + 1) * 4;
  sm.cursor = reinterpret_cast<int*>(raw + off); off += (size_t)cells * 4;
  sm.gacc = bwd ? reinterpret_cast<float*>(raw + off) : nullptr;
  return sm;
}

template <bool BWD>
__global__ void __launch_bounds__(256)
sil_kernel(const float* __restrict__ projects, const float* __restrict__ g_sil, int N, int Vs, int wh, int B, int G,
           float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char raw[];
  const SilSmem sm = carve_sil(raw, Vs, G * G, BWD);
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (BWD)
    for (int i = threadIdx.x; i < Vs * 2; i += blockDim.x) sm.gacc[i] = 0.f;
  bin_vertices(sm, projects + (size_t)n * Vs * 3, Vs, B, G);
  const int tx = (wh + kTileW - 1) / kTileW, ty = (wh + kTileH - 1) / kTileH;
  const int ntiles = tx * ty;
  const int t0 = (int)(((long long)ntiles * blockIdx.y) / gridDim.y);
  const int t1 = (int)(((long long)ntiles * (blockIdx.y + 1)) / gridDim.y);
  for (int t = t0 + warp; t < t1; t += nwarps) {
    const int c0 = (t % tx) * kTileW, r0 = (t / tx) * kTileH;
    const int c = c0 + (lane & (kTileW - 1)), r = r0 + (lane >> 3);
    const bool active = c < wh && r < wh;
    const int c1 = min(c0 + kTileW, wh) - 1, r1 = min(r0 + kTileH, wh) - 1;
    float best;
    int barg;
    tile_search(sm, B, G, min(c0 / B, G - 1), min(c1 / B, G - 1), min(r0 / B, G - 1), min(r1 / B, G - 1), active,
                (float)c, (float)r, best, barg);
    const float d = sqrtf(best);
    const float s = expf(__fdiv_rn(-d, 1.2f));           // tf.exp(tf.negative(norm) / 1.2) (:37)
    const size_t o = (((size_t)n * wh + (wh - 1 - r)) * wh + c) * 2;
    if (!BWD) {
      if (active) *reinterpret_cast<float2*>(out + o) = make_float2(1.0f - s, s);
    } else {
      int vid = -1;
      float cu = 0.f, cv = 0.f;
      if (active && barg >= 0) {
        const float2 g = *reinterpret_cast<const float2*>(g_sil + o);
        const float2 p = sm.suv[barg];
        const float du = __fsub_rn(p.x, (float)c), dv = __fsub_rn(p.y, (float)r);
        const float coef = (d > 0.f) ? (-(g.y - g.x) * s / 1.2f) / d : 0.f;
        cu = coef * du; cv = coef * dv;
        vid = sm.svid[barg];
      }
      const unsigned grp = __match_any_sync(0xffffffffu, vid);
      float su = 0.f, sv = 0.f;
#pragma unroll
      for (int src = 0; src < 32; ++src) {
        const float a = __shfl_sync(0xffffffffu, cu, src), b = __shfl_sync(0xffffffffu, cv, src);
        if ((grp >> src) & 1u) { su += a; sv += b; }
      }
      if (vid >= 0 && lane == __ffs(grp) - 1) {
        atomicAdd(&sm.gacc[vid * 2], su);
        atomicAdd(&sm.gacc[vid * 2 + 1], sv);
      }
    }
  }
  if (BWD) {
    __syncthreads();
    // gridDim.y blocks share one g_projects row: block 0 stores, the others add (row zeroed by the launcher)
    float* gp = out + (size_t)n * Vs * 3;
    for (int i = threadIdx.x; i < Vs * 2; i += blockDim.x) {
      const int v = i >> 1, k = i & 1;
      if (gridDim.y == 1) gp[v * 3 + k] = sm.gacc[i];
      else if (sm.gacc[i] != 0.f) atomicAdd(&gp[v * 3 + k], sm.gacc[i]);
    }
    if (gridDim.y == 1)
      for (int i = threadIdx.x; i < Vs; i += blockDim.x) gp[i * 3 + 2] = 0.f;
  }
}

size_t sil_smem_bytes(int Vs, int G, bool bwd) {
  return (size_t)Vs * 12 + (size_t)(2 * G * G + 1) * 4 + (bwd ? (size_t)Vs * 8 : 0) + 16;
}

void sil_grid(int wh, int& B, int& G) {
  B = 4;
  while ((wh + B - 1) / B > kMaxGrid) B *= 2;
  G = (wh + B - 1) / B;
}

template <bool BWD>
cudaError_t launch_sil(const float* projects, const float* g_sil, int N, int Vs, int wh, float* out, cudaStream_t st) {
  int B, G;
  sil_grid(wh, B, G);
  const size_t smem = sil_smem_bytes(Vs, G, BWD);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(sil_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int ntiles = ((wh + kTileW - 1) / kTileW) * ((wh + kTileH - 1) / kTileH);
  int split = 1;
  if (N < 2 * 148) split = max(1, min(ntiles / 8, (2 * 148 + N - 1) / N));
  if (BWD && split > 1) {
    e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N * Vs * 3, st);
    if (e != cudaSuccess) return e;
  }
  dim3 grid(N, split);
  LaunchScope scope(BWD ? KID_SIL_BWD : KID_SIL_FWD, st);
  sil_kernel<BWD><<<grid, 256, smem, st>>>(projects, g_sil, N, Vs, wh, B, G, out);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_sil_fwd(const float* projects, int N, int Vs, int wh, float* sil, cudaStream_t st) {
  return launch_sil<false>(projects, nullptr, N, Vs, wh, sil, st);
}

cudaError_t launch_sil_bwd(const float* projects, const float* g_sil, int N, int Vs, int wh, float* g_projects,
                           cudaStream_t st) {
  return launch_sil<true>(projects, g_sil, N, Vs, wh, g_projects, st);
}

}  // namespace smplb200
