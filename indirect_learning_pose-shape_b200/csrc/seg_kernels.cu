// K5: 31-part soft segmentation from projected vertices and the visibility weights, forward and backward.
//
// Reference arithmetic (keras_smpl/projects_to_seg.py:9-69), per part k and pixel g = (column c, row r):
//   s_k[g] = max_i exp(-(||p_i - g||_2 * w_i))        :52-56   (tf.norm = sqrt(du*du + dv*dv), no FMA)
//   bg[g]  = 1 - clip(sum_k s_k[g], 0, 1)             :61-64
//   out[n, wh-1-r, c, :] = [bg, s_0 .. s_30]          :66-68   (rows flipped)
//
// exp, sqrt and the multiply by w are monotone, so  max_i exp(-(d_i w_i)) = exp(-min_i (d_i w_i))  bit for bit: both
// directions are an exact weighted-nearest-vertex query.  Vertices are split by weight once per sample:
//   light    w == 1        one per occupied z-buffer cell after compute_mask; min over SQUARED distances in the hot
//                           loop, a single sqrt at the end (sqrt is monotone and correctly rounded)
//   heavy    w >= 256      d*w > 128 unless d < 0.5, and exp(-128) is exactly 0 in fp32, so a heavy vertex can only
//                           reach the one pixel it rounds to: chained per pixel, visited by that pixel alone
//   generic  anything else evaluated against every pixel (never produced by compute_mask; kept for drop-in inputs)
// The backward keeps (argmin slot, distance*weight) per (pixel, part) in a shared tile, then transposes the work:
// lane k owns part k and walks the 32 pixels of the group, merging runs of equal arg-min vertices in registers, so
// a run costs one shared-memory atomicAdd pair instead of one per pixel (run-length warp aggregation).
// Gradient conventions (TF autodiff): first arg-min takes the whole gradient on exact ties (TF splits
// evenly; ties have measure zero); d == 0 yields 0 where TF yields NaN; the clip gate is inclusive (0 <= sum <= 1).
#include <math_constants.h>
#include "common.cuh"

namespace smplb200 {

namespace {

constexpr float kHeavyMin = 256.0f;
constexpr float kDropX = 110.0f;      // exp(-x) == 0 in fp32 (denormals included) for x > 103.98
constexpr int kNone16 = 0xffff;

struct SegSmem {
  float4* ent;     // [E]  light: {u, v, 1, vid}   heavy/generic: {u, v, w, entry | next << 16}
  int* head;       // [wh*wh] first heavy slot of each pixel, -1 none
  int* lcount;     // [32] light entries per part (packed at the front of the part's CSR segment)
  int* ghead;      // [1]  chain of generic slots, -1 none
  float* tiles;    // per-warp staging
};

__device__ __forceinline__ float dist2(float u, float v, float gx, float gy) {
  const float du = __fsub_rn(u, gx), dv = __fsub_rn(v, gy);
  return __fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv));
}

// Split the sample's part vertices into weight classes (one warp per part, ballot compaction).
__device__ void classify(const SegSmem& sm, const float* __restrict__ proj, const float* __restrict__ mask,
                         const int* __restrict__ ptr, const int* __restrict__ idx, int P, int wh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < wh * wh; i += blockDim.x) sm.head[i] = -1;
  if (threadIdx.x == 0) *sm.ghead = -1;
  __syncthreads();
  for (int k = warp; k < P; k += nwarps) {
    const int p0 = ptr[k], p1 = ptr[k + 1];
    int nl = 0, no = 0;
    for (int base = p0; base < p1; base += 32) {
      const int e = base + lane;
      const bool in = e < p1;
      float u = 0.f, v = 0.f, w = 0.f;
      int vid = 0;
      if (in) {
        vid = idx[e];
        u = proj[vid * 3]; v = proj[vid * 3 + 1]; w = mask[vid];
      }
      const bool light = in && (w == 1.0f);
      const bool other = in && !light;
      const unsigned bl = __ballot_sync(0xffffffffu, light), bo = __ballot_sync(0xffffffffu, other);
      const unsigned lt = (1u << lane) - 1u;
      if (light) sm.ent[p0 + nl + __popc(bl & lt)] = make_float4(u, v, 1.0f, __int_as_float(vid));
      if (other) {
        const int slot = p1 - 1 - (no + __popc(bo & lt));
        int next = kNone16;
        bool keep = true;
        if (w >= kHeavyMin) {
          const float pu = rintf(u), pv = rintf(v);
          keep = pu >= 0.f && pu <= (float)(wh - 1) && pv >= 0.f && pv <= (float)(wh - 1);
          if (keep) keep = __fmul_rn(sqrtf(dist2(u, v, pu, pv)), w) <= kDropX;
          if (keep) next = atomicExch(&sm.head[(int)pv * wh + (int)pu], slot) & 0xffff;
        } else {
          next = atomicExch(sm.ghead, slot) & 0xffff;
        }
        // dropped heavy entries keep their slot but are never linked
        sm.ent[slot] = make_float4(u, v, w, __int_as_float((e & 0xffff) | (next << 16)));
        (void)keep;
      }
      nl += __popc(bl);
      no += __popc(bo);
    }
    if (lane == 0) sm.lcount[k] = nl;
  }
  __syncthreads();
}

// x = min_i d_i*w_i over the chained (heavy or generic) entries of part k; returns the best slot through `arg`.
__device__ __forceinline__ void walk_chain(const SegSmem& sm, int first, int k, const int* __restrict__ ptr, float gx,
                                           float gy, float& x, int& arg) {
  int slot = first;
  while (slot >= 0) {
    const float4 e = sm.ent[slot];
    const int meta = __float_as_int(e.w);
    if (slot >= ptr[k] && slot < ptr[k + 1]) {           // slots never leave their part's CSR segment
      const float xe = __fmul_rn(sqrtf(dist2(e.x, e.y, gx, gy)), e.z);
      if (xe < x) { x = xe; arg = slot; }
    }
    const int nx = (meta >> 16) & 0xffff;
    slot = (nx == kNone16) ? -1 : nx;
  }
}

// Weighted nearest vertex of part k for this lane's pixel.  TRACK also returns the winning slot.
template <bool TRACK>
__device__ __forceinline__ float part_query(const SegSmem& sm, const int* __restrict__ ptr, int k, float gx, float gy,
                                            int myhead, int ghead, int& arg) {
  const int p0 = ptr[k], nl = sm.lcount[k];
  float best = CUDART_INF_F;
  int barg = -1;
#pragma unroll 4
  for (int i = 0; i < nl; ++i) {
    const float4 e = sm.ent[p0 + i];                      // same address on every lane: broadcast
    const float d2 = dist2(e.x, e.y, gx, gy);
    if (TRACK) {
      if (d2 < best) { best = d2; barg = p0 + i; }
    } else {
      best = fminf(best, d2);
    }
  }
  float x = sqrtf(best);                                  // w == 1: d*w == d
  if (myhead >= 0) walk_chain(sm, myhead, k, ptr, gx, gy, x, barg);
  if (ghead >= 0) walk_chain(sm, ghead, k, ptr, gx, gy, x, barg);
  arg = barg;
  return x;
}

__device__ __forceinline__ SegSmem carve(unsigned char* raw, int E, int wh, int Vs_acc) {
  SegSmem sm;
  size_t off = 0;
  sm.ent = reinterpret_cast<float4*>(raw + off); off += (size_t)((E + 1) & ~1) * 16;
  sm.head = reinterpret_cast<int*>(raw + off); off += (size_t)wh * wh * 4;
  sm.lcount = reinterpret_cast<int*>(raw + off); off += 32 * 4;
  sm.ghead = reinterpret_cast<int*>(raw + off); off += 16;
  off += (size_t)Vs_acc * 8;                              // backward accumulator sits here (see seg_bwd_kernel)
  sm.tiles = reinterpret_cast<float*>(raw + off);
  return sm;
}

__global__ void __launch_bounds__(256)
seg_fwd_kernel(const float* __restrict__ projects, const float* __restrict__ mask, int N, int Vs,
               const int* __restrict__ ptr, const int* __restrict__ idx, int P, int E, int wh,
               float* __restrict__ seg) {
  extern __shared__ __align__(16) unsigned char raw[];
  const SegSmem sm = carve(raw, E, wh, 0);
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  classify(sm, projects + (size_t)n * Vs * 3, mask + (size_t)n * Vs, ptr, idx, P, wh);
  const int ghead = *sm.ghead;
  const int C = P + 1;
  float* tile = sm.tiles + warp * (32 * 33);
  const int npix = wh * wh, ngroups = (npix + 31) / 32;
  // this block's share of the pixel groups (gridDim.y splits a sample when the batch alone cannot fill the GPU)
  const int g0 = (int)(((long long)ngroups * blockIdx.y) / gridDim.y);
  const int g1 = (int)(((long long)ngroups * (blockIdx.y + 1)) / gridDim.y);
  for (int grp = g0 + warp; grp < g1; grp += nwarps) {
    const int pix = grp * 32 + lane;
    const bool active = pix < npix;
    const int r = pix / wh, c = pix - r * wh;
    const float gx = (float)c, gy = (float)r;             // grid = (column, row) (:26-31)
    const int myhead = active ? sm.head[pix] : -1;
    float S = 0.f;
    for (int k = 0; k < P; ++k) {
      int arg;
      const float x = part_query<false>(sm, ptr, k, gx, gy, myhead, ghead, arg);
      const float s = expf(-x);
      tile[lane * 33 + 1 + k] = s;
      S += s;
    }
    tile[lane * 33] = 1.0f - fminf(fmaxf(S, 0.f), 1.f);
    __syncwarp();
    const int cnt = min(32, npix - grp * 32);
    for (int j = 0; j < cnt; ++j) {
      const int pj = grp * 32 + j;
      const int rj = pj / wh, cj = pj - rj * wh;
      if (lane < C) seg[(((size_t)n * wh + (wh - 1 - rj)) * wh + cj) * C + lane] = tile[j * 33 + lane];
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256)
seg_bwd_kernel(const float* __restrict__ projects, const float* __restrict__ mask, const float* __restrict__ g_seg,
               int N, int Vs, const int* __restrict__ ptr, const int* __restrict__ idx, int P, int E, int wh,
               float* __restrict__ g_projects) {
  extern __shared__ __align__(16) unsigned char raw[];
  const SegSmem sm = carve(raw, E, wh, Vs);
  float* gacc = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sm.ghead) + 16);   // [Vs][2]
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < Vs * 2; i += blockDim.x) gacc[i] = 0.f;
  classify(sm, projects + (size_t)n * Vs * 3, mask + (size_t)n * Vs, ptr, idx, P, wh);
  const int ghead = *sm.ghead;
  const int C = P + 1;
  float* tileX = sm.tiles + warp * (2 * 32 * 33 + 32);
  int* tileI = reinterpret_cast<int*>(tileX + 32 * 33);
  float* gate = tileX + 2 * 32 * 33;
  const int npix = wh * wh, ngroups = (npix + 31) / 32;
  for (int grp = warp; grp < ngroups; grp += nwarps) {
    // ---- phase 1: lane = pixel; weighted nearest vertex of every part, and the clip gate ----------------------
    {
      const int pix = grp * 32 + lane;
      const bool active = pix < npix;
      const int r = pix / wh, c = pix - r * wh;
      const float gx = (float)c, gy = (float)r;
      const int myhead = active ? sm.head[pix] : -1;
      float S = 0.f;
      for (int k = 0; k < P; ++k) {
        int arg;
        const float x = part_query<true>(sm, ptr, k, gx, gy, myhead, ghead, arg);
        tileX[lane * 33 + k] = x;
        tileI[lane * 33 + k] = arg;
        S += expf(-x);
      }
      gate[lane] = (S >= 0.f && S <= 1.f) ? 1.f : 0.f;    // clip_by_value passes gradient on the closed interval
    }
    __syncwarp();
    // ---- phase 2: lane = part; walk the group's pixels, merge runs of equal arg-min vertices -------------------
    {
      const int cnt = min(32, npix - grp * 32);
      int run_vid = -1;
      float run_u = 0.f, run_v = 0.f;
      for (int j = 0; j < cnt; ++j) {
        const int pj = grp * 32 + j;
        const int rj = pj / wh, cj = pj - rj * wh;
        const float gch = (lane < C) ? g_seg[(((size_t)n * wh + (wh - 1 - rj)) * wh + cj) * C + lane] : 0.f;
        const float g0 = __shfl_sync(0xffffffffu, gch, 0);
        const float gk = __shfl_down_sync(0xffffffffu, gch, 1);     // lane k <- channel k+1
        if (lane < P) {
          const int arg = tileI[j * 33 + lane];
          const float x = tileX[j * 33 + lane];
          int vid = -1;
          float cu = 0.f, cv = 0.f;
          if (arg >= 0) {
            const float4 e = sm.ent[arg];
            const float s = expf(-x);
            const float G = gk - gate[j] * g0;
            const float du = __fsub_rn(e.x, (float)cj), dv = __fsub_rn(e.y, (float)rj);
            const float d = sqrtf(__fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv)));
            const float coef = (d > 0.f) ? (-e.z * s * G) / d : 0.f;  // d(exp(-d w))/dp = -w s (p - g)/d
            cu = coef * du; cv = coef * dv;
            const int meta = __float_as_int(e.w);
            vid = (e.z == 1.0f) ? meta : idx[meta & 0xffff];
          }
          if (vid != run_vid) {
            if (run_vid >= 0) { atomicAdd(&gacc[run_vid * 2], run_u); atomicAdd(&gacc[run_vid * 2 + 1], run_v); }
            run_vid = vid; run_u = cu; run_v = cv;
          } else {
            run_u += cu; run_v += cv;
          }
        }
      }
      if (run_vid >= 0) { atomicAdd(&gacc[run_vid * 2], run_u); atomicAdd(&gacc[run_vid * 2 + 1], run_v); }
    }
    __syncwarp();
  }
  __syncthreads();
  float* out = g_projects + (size_t)n * Vs * 3;
  for (int i = threadIdx.x; i < Vs * 3; i += blockDim.x) {
    const int v = i / 3, c = i - v * 3;
    out[i] = (c < 2) ? gacc[v * 2 + c] : 0.f;             // z receives no gradient from the rasteriser
  }
}

size_t seg_smem_bytes(int E, int wh, int Vs_acc, int warps, bool bwd) {
  size_t b = (size_t)((E + 1) & ~1) * 16 + (size_t)wh * wh * 4 + 32 * 4 + 16 + (size_t)Vs_acc * 8;
  b += (size_t)warps * (bwd ? (2 * 32 * 33 + 32) : (32 * 33)) * 4;
  return b;
}

constexpr size_t kMaxSmem = 227 * 1024;

}  // namespace

cudaError_t launch_seg_fwd(const SmplB200Parts* p, const float* projects, const float* mask, int N, int Vs, int wh,
                           float* seg, cudaStream_t st) {
  int warps = 8;
  while (warps > 1 && seg_smem_bytes(p->E, wh, 0, warps, false) > kMaxSmem) warps >>= 1;
  const size_t smem = seg_smem_bytes(p->E, wh, 0, warps, false);
  if (smem > kMaxSmem) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(seg_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // split a sample's pixel groups over gridDim.y when the batch alone leaves SMs idle
  const int ngroups = (wh * wh + 31) / 32;
  int split = 1;
  if (N < 2 * 148) split = max(1, min(ngroups / warps, (2 * 148 + N - 1) / N));
  dim3 grid(N, split);
  LaunchScope scope(KID_SEG_FWD, st);
  seg_fwd_kernel<<<grid, warps * 32, smem, st>>>(projects, mask, N, Vs, p->ptr, p->idx, p->P, p->E, wh, seg);
  return cudaGetLastError();
}

cudaError_t launch_seg_bwd(const SmplB200Parts* p, const float* projects, const float* mask, const float* g_seg, int N,
                           int Vs, int wh, float* g_projects, cudaStream_t st) {
  int warps = 8;
  while (warps > 1 && seg_smem_bytes(p->E, wh, Vs, warps, true) > kMaxSmem) warps >>= 1;
  const size_t smem = seg_smem_bytes(p->E, wh, Vs, warps, true);
  if (smem > kMaxSmem) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(seg_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  LaunchScope scope(KID_SEG_BWD, st);
  seg_bwd_kernel<<<N, warps * 32, smem, st>>>(projects, mask, g_seg, N, Vs, p->ptr, p->idx, p->P, p->E, wh, g_projects);
  return cudaGetLastError();
}

}  // namespace smplb200
