// K5: 31-part soft segmentation from projected vertices and the visibility weights, forward and backward.
//
// Reference arithmetic (keras_smpl/projects_to_seg.py:9-69), per part k and pixel g = (column c, row r):
//   s_k[g] = max_i exp(-(||p_i - g||_2 * w_i))        :52-56
//   bg[g]  = 1 - clip(sum_k s_k[g], 0, 1)             :61-64
//   out[n, wh-1-r, c, :] = [bg, s_0 .. s_30]          :66-68   (rows flipped)
//
// exp, sqrt and the multiply by w are monotone, so  max_i exp(-(d_i w_i)) = exp(-min_i (d_i w_i)): both directions are
// an exact weighted-nearest-vertex query (arg-min over squared distances computed as fl(fl(du^2)+fl(dv^2)), the same
// roundings as tf.norm).  Vertices are split by weight once per sample:
//   light    w == 1        one per occupied z-buffer cell after compute_mask; min over SQUARED distances in the hot loop
//   heavy    w >= 256      d*w > 128 unless d < 0.5, and exp(-128) is exactly 0 in fp32, so a heavy vertex can only
//                           reach the one pixel it rounds to: chained per pixel, visited by that pixel alone
//   generic  anything else evaluated against every pixel (never produced by compute_mask; kept for drop-in inputs)
//
// Forward: a warp owns a tile of (LX*BW) x (LY*BH) pixels, a lane a BW x BH block of it, so the per-axis squared
// offsets du^2 (BW values) and dv^2 (BH values) are shared by the block's pixels.  Eight channels of each pixel are
// kept in registers and leave as ONE 32-byte store (st.global.v8.f32 = a full DRAM sector; partial-sector stores cost
// a read-fill).  The score epilogue is d = d2*rsqrt(d2), s = ex2(-d*log2 e): <= 3e-7 absolute from exp(-sqrt(d2)),
// an order of magnitude inside the 1e-5 geometry tolerance that bounds the inputs.
// When a backward will follow, the forward also records per pixel the clip gate and the arg-min of every part as one
// byte (`saved`, 32 B per pixel, laid out in the forward's own (tile, block, lane) order), so the backward never
// searches.
//
// Backward: lane = channel.  Records are read back in the forward's order (no index arithmetic beyond shifts), the
// upstream gradient row of the pixel is one coalesced 128-byte load, s is recomputed at the recorded arg-min, and the
// per-vertex sums accumulate in per-warp PRIVATE shared-memory slots indexed by the part's light slot: lane k is the
// only writer of part k's slots, so plain load/add/store replaces atomics (shared fp32 atomicAdd is a CAS loop on
// sm_100).  Runs of equal arg-min vertices along a row are merged in registers first.
// Gradient conventions (TF autodiff): the first arg-min takes the whole gradient on exact ties (TF splits
// evenly; measure zero); d == 0 yields 0 where TF yields NaN; the clip gate is inclusive (0 <= sum <= 1).
#include <math_constants.h>
#include "common.cuh"

namespace smplb200 {

namespace {

constexpr float kHeavyMin = 256.0f;
constexpr float kDropX = 110.0f;      // exp(-x) == 0 in fp32 (denormals included) for x > 103.98
constexpr unsigned kNone16 = 0xffffu;
constexpr int kArgNone = 0xff;        // saved byte: the part has no vertex that reaches this pixel (zero gradient)
constexpr int kArgSlow = 0xfe;        // saved byte: winner is a heavy/generic vertex or index >= 254: re-query
constexpr float kLog2e = 1.4426950408889634f;

struct SegSmem {
  float4* ent;     // [E]  light: {u, v, 1, vid}   heavy/generic: {u, v, w, entry | next << 16}
  int* head;       // [wh*wh] first heavy slot of each pixel, -1 none
  int* lcount;     // [32] light entries per part (packed at the front of the part's CSR segment)
  int* lbase;      // [32] exclusive prefix sum of lcount
  int* pptr;       // [36] the part table's CSR pointers (P+1 used)
  int* ghead;      // [1]  chain of generic slots, -1 none
  int* nheavy;     // [1]  number of chained heavy entries
  unsigned char* rest;
};

__device__ __forceinline__ float dist2(float u, float v, float gx, float gy) {
  const float du = __fsub_rn(u, gx), dv = __fsub_rn(v, gy);
  return __fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv));
}

// exp(-sqrt(d2)) for the light class, fast path (see header).  d2 = +inf (empty part) -> 0.
__device__ __forceinline__ float score_from_d2(float d2) {
  const float d = d2 * rsqrtf(fmaxf(d2, 1e-30f));            // d2 == 0 -> 0 ; inf*0 is NaN, handled below
  return (d2 < CUDART_INF_F) ? exp2f(-d * kLog2e) : 0.f;     // exp2f lowers to ex2.approx with a range fix-up
}

__device__ __forceinline__ SegSmem carve(unsigned char* raw, int E, int wh) {
  SegSmem sm;
  size_t off = 0;
  sm.ent = reinterpret_cast<float4*>(raw + off); off += (size_t)((E + 1) & ~1) * 16;
  sm.head = reinterpret_cast<int*>(raw + off); off += ((size_t)wh * wh * 4 + 15) & ~(size_t)15;
  sm.lcount = reinterpret_cast<int*>(raw + off); off += 32 * 4;
  sm.lbase = reinterpret_cast<int*>(raw + off); off += 32 * 4;
  sm.pptr = reinterpret_cast<int*>(raw + off); off += 36 * 4;
  sm.ghead = reinterpret_cast<int*>(raw + off);
  sm.nheavy = sm.ghead + 1; off += 16;
  sm.rest = raw + off;
  return sm;
}
size_t seg_base_smem(int E, int wh) {
  return (size_t)((E + 1) & ~1) * 16 + (((size_t)wh * wh * 4 + 15) & ~(size_t)15) + 2 * 32 * 4 + 36 * 4 + 16;
}

// Split the sample's part vertices into weight classes (one warp per part, ballot compaction).
__device__ void classify(const SegSmem& sm, const float* __restrict__ proj, const float* __restrict__ mask,
                         const int* __restrict__ ptr, const int* __restrict__ idx, int P, int wh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < wh * wh; i += blockDim.x) sm.head[i] = -1;
  if (threadIdx.x == 0) { *sm.ghead = -1; *sm.nheavy = 0; }
  if (threadIdx.x < 32) sm.lcount[threadIdx.x] = 0;
  if (threadIdx.x < 36) sm.pptr[threadIdx.x] = ptr[min((int)threadIdx.x, P)];
  __syncthreads();
  for (int k = warp; k < P; k += nwarps) {
    const int p0 = ptr[k], p1 = ptr[k + 1];
    int nl = 0, no = 0;
    constexpr int kU = 4;                                // chunks of 32 entries whose gather chains overlap
    for (int base0 = p0; base0 < p1; base0 += 32 * kU) {
      int vids[kU];
      float us[kU], vv[kU], ws[kU];
#pragma unroll
      for (int c = 0; c < kU; ++c) {                     // idx -> (proj, mask) is a dependent chain: issue kU of them
        const int e = base0 + c * 32 + lane;
        vids[c] = (e < p1) ? idx[e] : 0;
      }
#pragma unroll
      for (int c = 0; c < kU; ++c) {
        const bool in = base0 + c * 32 + lane < p1;
        us[c] = in ? proj[vids[c] * 3] : 0.f;
        vv[c] = in ? proj[vids[c] * 3 + 1] : 0.f;
        ws[c] = in ? mask[vids[c]] : 0.f;
      }
#pragma unroll
      for (int c = 0; c < kU; ++c) {
        const int base = base0 + c * 32;
        if (base >= p1) break;
        const int e = base + lane;
        const bool in = e < p1;
        const float u = us[c], v = vv[c], w = ws[c];
        const int vid = vids[c];
        const bool light = in && (w == 1.0f);
        const bool other = in && !light;
        const unsigned bl = __ballot_sync(0xffffffffu, light), bo = __ballot_sync(0xffffffffu, other);
        const unsigned lt = (1u << lane) - 1u;
        if (light) sm.ent[p0 + nl + __popc(bl & lt)] = make_float4(u, v, 1.0f, __int_as_float(vid));
        if (other) {
          const int slot = p1 - 1 - (no + __popc(bo & lt));
          unsigned next = kNone16;
          // heavy entries are linked in a second pass (below), once the part's light list is complete
          if (!(w >= kHeavyMin)) next = (unsigned)atomicExch(sm.ghead, slot) & 0xffffu;
          sm.ent[slot] = make_float4(u, v, w, __uint_as_float(((unsigned)e & 0xffffu) | (next << 16)));
        }
        nl += __popc(bl);
        no += __popc(bo);
      }
    }
    if (lane == 0) sm.lcount[k] = nl;
    __syncwarp();
    // Second pass over this part's heavy entries (they sit at the back of the segment): a heavy vertex reaches only
    // the pixel it rounds to, and matters only if it BEATS the part's light minimum there -- the same comparison
    // (x_heavy < sqrt(min d2_light)) the per-pixel path makes, so dropping the losers is exact.  Winners are chained.
    for (int base = p1 - no; base < p1; base += 32) {
      const int slot = base + lane;
      if (slot < p1) {
        const float4 h = sm.ent[slot];
        if (h.z >= kHeavyMin) {
          const float pu = rintf(h.x), pv = rintf(h.y);
          if (pu >= 0.f && pu <= (float)(wh - 1) && pv >= 0.f && pv <= (float)(wh - 1)) {
            const float xh = __fmul_rn(sqrtf(dist2(h.x, h.y, pu, pv)), h.z);
            if (xh <= kDropX) {
              float best = CUDART_INF_F;
              for (int i = 0; i < nl; ++i) {
                const float4 l = sm.ent[p0 + i];
                best = fminf(best, dist2(l.x, l.y, pu, pv));
              }
              if (xh < sqrtf(best)) {
                const unsigned next = (unsigned)atomicExch(&sm.head[(int)pv * wh + (int)pu], slot) & 0xffffu;
                sm.ent[slot].w = __uint_as_float((__float_as_uint(h.w) & 0xffffu) | (next << 16));
                atomicAdd(sm.nheavy, 1);
              }
            }
          }
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) {                               // exclusive scan of the light counts
    const int c = sm.lcount[threadIdx.x];
    int s = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, o);
      if ((int)threadIdx.x >= o) s += t;
    }
    sm.lbase[threadIdx.x] = s - c;
  }
  __syncthreads();
}

// x = min_i d_i*w_i over the chained (heavy or generic) entries of part [p0,p1); returns the best slot through `arg`.
__device__ __forceinline__ void walk_chain(const SegSmem& sm, int first, int p0, int p1, float gx, float gy, float& x,
                                           int& arg) {
  int slot = first;
  while (slot >= 0) {
    const float4 e = sm.ent[slot];
    if (slot >= p0 && slot < p1) {                       // slots never leave their part's CSR segment
      const float xe = __fmul_rn(sqrtf(dist2(e.x, e.y, gx, gy)), e.z);
      if (xe < x) { x = xe; arg = slot; }
    }
    const unsigned nx = (__float_as_uint(e.w) >> 16) & 0xffffu;
    slot = (nx == kNone16) ? -1 : (int)nx;
  }
}

// Out-of-line slow path of the forward: does a heavy/generic vertex beat the light minimum d2 at this pixel?
// Returns the winning score (exact expf) or a negative value if the light minimum stands.
__device__ __noinline__ float slow_pixel_score(const SegSmem& sm, int p0, int p1, float gx, float gy, int head, int ghead,
                                               float d2_light) {
  float x = sqrtf(d2_light);
  int a = -1;
  if (head >= 0) walk_chain(sm, head, p0, p1, gx, gy, x, a);
  if (ghead >= 0) walk_chain(sm, ghead, p0, p1, gx, gy, x, a);
  return (a >= 0) ? expf(-x) : -1.0f;
}

// Out-of-line slow path of the backward: full weighted nearest-vertex query of part [p0,p1) for one pixel.
__device__ __noinline__ int slow_pixel_query(const SegSmem& sm, int p0, int p1, int nl, float gx, float gy, int head,
                                             int ghead) {
  float best = CUDART_INF_F;
  int barg = -1;
  for (int i = 0; i < nl; ++i) {
    const float4 e = sm.ent[p0 + i];
    const float d2 = dist2(e.x, e.y, gx, gy);
    if (d2 < best) { best = d2; barg = p0 + i; }
  }
  float x = sqrtf(best);
  if (head >= 0) walk_chain(sm, head, p0, p1, gx, gy, x, barg);
  if (ghead >= 0) walk_chain(sm, ghead, p0, p1, gx, gy, x, barg);
  return barg;
}

__device__ __forceinline__ void st_global_v8(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// Tile geometry shared by forward and backward (the `saved` layout depends on it).
struct SegGeom {
  int BW, BH, LX, LY, TW, TH, NB, tiles_x, tiles_y, ntiles;
};
SegGeom seg_geom(int wh) {
  SegGeom g;
  if (wh % 12 == 0) { g.BW = 3; g.BH = 2; g.LX = 4; }         // 12 x 16 pixel tiles (48 = 4 x 3 tiles)
  else { g.BW = 4; g.BH = 2; g.LX = 4; }                      // 16 x 16 pixel tiles
  g.LY = 32 / g.LX; g.TW = g.LX * g.BW; g.TH = g.LY * g.BH; g.NB = g.BW * g.BH;
  g.tiles_x = (wh + g.TW - 1) / g.TW; g.tiles_y = (wh + g.TH - 1) / g.TH; g.ntiles = g.tiles_x * g.tiles_y;
  return g;
}

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
constexpr float kBigD2 = 1e30f;      // "no vertex yet": rsqrt/ex2 map it to a score of exactly 0 without a branch
constexpr int kNoPrune = 3;          // parts with this few visible vertices skip the pruning pass
constexpr float kPruneMargin = 0.01f;

__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// explicit shared-window accesses (a generic pointer into dynamic smem makes the compiler rebuild the window base)
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
  float2 r;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(a));
  return r;
}
__device__ __forceinline__ void sts_f2(uint32_t a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}

// saved layout: 4 planes (channels 8q..8q+7); plane q holds, for record id = (tile*NB + b)*32 + lane, 8 bytes.
// byte of channel 0: bit 0 = clip gate.  byte of channel 1+k: 0 none, 1..254 light index + 1, 255 re-query.
template <int BW, int BH, int LX, bool TRACK>
__global__ void __launch_bounds__(256, 3)
seg_fwd_kernel(const float* __restrict__ projects, const float* __restrict__ mask, int N, int Vs,
               const int* __restrict__ ptr, const int* __restrict__ idx, int P, int E, int wh,
               float* __restrict__ seg, unsigned char* __restrict__ saved) {
  constexpr int LY = 32 / LX, TW = LX * BW, TH = LY * BH, NB = BW * BH;
  extern __shared__ __align__(16) unsigned char raw[];
  const SegSmem sm = carve(raw, E, wh);
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  classify(sm, projects + (size_t)n * Vs * 3, mask + (size_t)n * Vs, ptr, idx, P, wh);
  const int ghead = *sm.ghead;
  const bool any_heavy = *sm.nheavy > 0;
  const int C = P + 1;
  const int tiles_x = (wh + TW - 1) / TW, tiles_y = (wh + TH - 1) / TH, ntiles = tiles_x * tiles_y;
  const int t0 = (int)(((long long)ntiles * blockIdx.y) / gridDim.y);
  const int t1 = (int)(((long long)ntiles * (blockIdx.y + 1)) / gridDim.y);
  const int lx = lane % LX, ly = lane / LX;
  const size_t plane = (size_t)ntiles * NB * 32 * 8;               // bytes of one saved plane of one sample
  unsigned char* sv = TRACK ? saved + (size_t)n * plane * 4 : nullptr;
  float* seg_n = seg ? seg + (size_t)n * wh * wh * C : nullptr;
  // per-lane staging of one chunk: stage[(sub*NB + q)*32 + lane]  (lane-contiguous: conflict-free, no sync needed)
  float* stage = reinterpret_cast<float*>(sm.rest) + (size_t)warp * (8 * NB * 32) + lane;

  for (int t = t0 + warp; t < t1; t += nwarps) {
    const int ty = t / tiles_x, tx = t - ty * tiles_x;
    const int c0 = tx * TW + lx * BW, r0 = ty * TH + ly * BH;      // this lane's block origin (grid = (column,row), :26-31)
    const float cx0 = (float)(tx * TW), cx1 = (float)(tx * TW + TW - 1);   // tile corners (pixel centres) and centre
    const float cy0 = (float)(ty * TH), cy1 = (float)(ty * TH + TH - 1);
    const float tcx = 0.5f * (cx0 + cx1), tcy = 0.5f * (cy0 + cy1);
    float gxs[BW], gys[BH];
#pragma unroll
    for (int i = 0; i < BW; ++i) gxs[i] = (float)(c0 + i);
#pragma unroll
    for (int j = 0; j < BH; ++j) gys[j] = (float)(r0 + j);
    bool blk_slow = ghead >= 0;                                    // any heavy vertex chained to one of my pixels?
    if (any_heavy) {
#pragma unroll
      for (int j = 0; j < BH; ++j)
#pragma unroll
        for (int i = 0; i < BW; ++i)
          if (c0 + i < wh && r0 + j < wh) blk_slow |= sm.head[(r0 + j) * wh + c0 + i] >= 0;
    }
    float S[NB];
#pragma unroll
    for (int q = 0; q < NB; ++q) S[q] = 0.f;

    // Exact pruning, once per tile, lane k working on part k.  i0 = the part's vertex nearest the tile centre.
    // f(g) = d_j^2(g) - d_i0^2(g) is linear in the pixel g, so f >= margin at the tile's four corners implies f >= margin on
    // every pixel of the tile: vertex j can never be the arg-min there and is dropped from the part's survivor mask.
    // The margin (0.01 px^2) dwarfs the fp32 rounding of the squared distances (< 3e-3 at wh <= 128), so the survivors
    // always contain the exact fp32 arg-min.
    unsigned keepmask = 0xffffffffu;
    {
      const int kp = lane < P ? lane : 0;
      const int q0 = sm.pptr[kp], n0 = (lane < P) ? sm.lcount[kp] : 0;
      if (n0 > kNoPrune && n0 <= 32) {
        float bd = CUDART_INF_F;
        int bi = 0;
        for (int v = 0; v < n0; ++v) {
          const float2 e = *reinterpret_cast<const float2*>(&sm.ent[q0 + v]);
          const float d = dist2(e.x, e.y, tcx, tcy);
          if (d < bd) { bd = d; bi = v; }
        }
        const float2 e0 = *reinterpret_cast<const float2*>(&sm.ent[q0 + bi]);
        const float a0 = dist2(e0.x, e0.y, cx0, cy0), a1 = dist2(e0.x, e0.y, cx1, cy0);
        const float a2 = dist2(e0.x, e0.y, cx0, cy1), a3 = dist2(e0.x, e0.y, cx1, cy1);
        keepmask = 0u;
        for (int v = 0; v < n0; ++v) {
          const float2 e = *reinterpret_cast<const float2*>(&sm.ent[q0 + v]);
          const bool dominated = dist2(e.x, e.y, cx0, cy0) - a0 >= kPruneMargin && dist2(e.x, e.y, cx1, cy0) - a1 >= kPruneMargin &&
                                 dist2(e.x, e.y, cx0, cy1) - a2 >= kPruneMargin && dist2(e.x, e.y, cx1, cy1) - a3 >= kPruneMargin;
          keepmask |= dominated ? 0u : (1u << v);
        }
      }
    }

    // channel chunks 1, 2, 3, then chunk 0 last: its channel 0 (background) needs the sum over all parts
    for (int cc = 1; cc <= 4; ++cc) {
      const int chunk = cc & 3;
      if (chunk * 8 >= C) continue;
      unsigned clo[NB], chi[NB];                                   // packed saved bytes, channels 0-3 / 4-7 of the chunk
#pragma unroll
      for (int q = 0; q < NB; ++q) { clo[q] = 0u; chi[q] = 0u; }
#pragma unroll 1
      for (int sub = 0; sub < 8; ++sub) {
        const int ch = chunk * 8 + sub;
        float sc[NB];
        int code[NB];
        if (ch == 0 || ch >= C) {                                  // background is filled in after the loop
#pragma unroll
          for (int q = 0; q < NB; ++q) stage[(sub * NB + q) * 32] = 0.f;
          continue;
        }
        const int p0 = sm.pptr[ch - 1], p1 = sm.pptr[ch], nl = sm.lcount[ch - 1];
        float best[NB];
        int barg[NB];
#pragma unroll
        for (int q = 0; q < NB; ++q) { best[q] = kBigD2; barg[q] = -1; }
        const unsigned pm = __shfl_sync(0xffffffffu, keepmask, ch - 1);   // survivors of this part (lane ch-1 pruned it)
        for (int cb = 0; cb < nl; cb += 32) {
          const int cnt = min(32, nl - cb);
          unsigned m = (cnt == 32) ? 0xffffffffu : ((1u << cnt) - 1u);
          if (nl <= 32) m &= pm;                                     // parts with more than 32 visible vertices are not pruned
          while (m) {
            const int v = cb + __ffs(m) - 1;
            m &= m - 1;
            const float2 e = *reinterpret_cast<const float2*>(&sm.ent[p0 + v]);   // same address on every lane: broadcast
            float du2[BW], dv2[BH], d2[NB];
#pragma unroll
            for (int i = 0; i < BW; ++i) { const float d = __fsub_rn(e.x, gxs[i]); du2[i] = __fmul_rn(d, d); }
#pragma unroll
            for (int j = 0; j < BH; ++j) { const float d = __fsub_rn(e.y, gys[j]); dv2[j] = __fmul_rn(d, d); }
#pragma unroll
            for (int j = 0; j < BH; ++j)
#pragma unroll
              for (int i = 0; i < BW; ++i) d2[j * BW + i] = __fadd_rn(du2[i], dv2[j]);
#pragma unroll
            for (int q = 0; q < NB; ++q) {
              if (TRACK) barg[q] = (d2[q] < best[q]) ? v : barg[q];
              best[q] = fminf(best[q], d2[q]);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < NB; ++q) {
          const float rs = rsqrt_approx(fmaxf(best[q], 1e-30f));
          sc[q] = ex2_approx((best[q] * rs) * (-kLog2e));          // exp(-sqrt(d2)); kBigD2 -> 0, d2 == 0 -> 1
          code[q] = min(barg[q] + 1, 255);                         // 0 none, 1..254 index+1, 255 re-query
        }
        if (blk_slow) {                                            // rare: heavy / generic vertices
#pragma unroll
          for (int q = 0; q < NB; ++q) {
            const int i = q % BW, j = q / BW;
            if (c0 + i < wh && r0 + j < wh) {
              const int hd = any_heavy ? sm.head[(r0 + j) * wh + c0 + i] : -1;
              if (hd >= 0 || ghead >= 0) {
                const float ss = slow_pixel_score(sm, p0, p1, gxs[i], gys[j], hd, ghead, best[q]);
                if (ss >= 0.f) { sc[q] = ss; code[q] = 255; }
              }
            }
          }
        }
        const unsigned sh = 8u * (sub & 3);
#pragma unroll
        for (int q = 0; q < NB; ++q) {
          stage[(sub * NB + q) * 32] = sc[q];
          S[q] += sc[q];
          if (TRACK) {
            if (sub < 4) clo[q] |= (unsigned)code[q] << sh;
            else chi[q] |= (unsigned)code[q] << sh;
          }
        }
      }
      if (chunk == 0) {
#pragma unroll
        for (int q = 0; q < NB; ++q) {
          stage[q * 32] = 1.0f - fminf(fmaxf(S[q], 0.f), 1.f);     // :61-64
          if (TRACK) clo[q] |= (S[q] >= 0.f && S[q] <= 1.f) ? 1u : 0u;   // clip gate (inclusive), for the backward
        }
      }
      // one 32-byte sector per pixel; rows flipped (:68)
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        const int r = r0 + q / BW, c = c0 + q % BW;
        if (seg_n && r < wh && c < wh) {
          float* o = seg_n + ((wh - 1 - r) * wh + c) * C + chunk * 8;
          float v8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v8[e] = stage[(e * NB + q) * 32];
          if (C == 32) {
            st_global_v8(o, v8);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (chunk * 8 + e < C) o[e] = v8[e];
          }
        }
        if (TRACK) {
          uint2* so = reinterpret_cast<uint2*>(sv + (size_t)chunk * plane) + ((t * NB + q) * 32 + lane);
          *so = make_uint2(clo[q], chi[q]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBatch = 8;          // records whose loads are in flight together (consecutive lanes of one block row)
constexpr int kAccSlots = 512;     // private accumulator slots per warp (light entries beyond this use atomics)

// rare: winner is a heavy/generic vertex (or a light index that did not fit a byte): exact re-query, atomics
__device__ __noinline__ void slow_pixel_grad(const SegSmem& sm, const int* __restrict__ idx, float* gacc, int p0, int p1,
                                             int nl, float gx, float gy, int head, int ghead, float G) {
  const int slot = slow_pixel_query(sm, p0, p1, nl, gx, gy, head, ghead);
  if (slot < 0) return;
  const float4 e = sm.ent[slot];
  const float du = __fsub_rn(e.x, gx), dv = __fsub_rn(e.y, gy);
  const float d = sqrtf(__fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv)));
  const float s = expf(-__fmul_rn(d, e.z));
  const float coef = (d > 0.f) ? (-e.z * s * G) / d : 0.f;         // d(exp(-d w))/dp = -w s (p - g)/d
  const int vid = (e.z == 1.0f) ? __float_as_int(e.w) : idx[__float_as_uint(e.w) & 0xffffu];
  atomicAdd(&gacc[vid * 2], coef * du);
  atomicAdd(&gacc[vid * 2 + 1], coef * dv);
}

template <int BW, int BH, int LX, bool C32, bool ALLIN>
__global__ void __launch_bounds__(256, 3)
seg_bwd_kernel(const float* __restrict__ projects, const float* __restrict__ mask, const float* __restrict__ g_seg,
               const unsigned char* __restrict__ saved, int N, int Vs, const int* __restrict__ ptr,
               const int* __restrict__ idx, int P, int E, int wh, int tiles_x, int ntiles,
               float* __restrict__ g_projects) {
  constexpr int LY = 32 / LX, TW = LX * BW, TH = LY * BH, NB = BW * BH;
  static_assert(LX == 4 && kBatch == 8, "a batch of 8 records = two lane rows of 4 blocks");
  extern __shared__ __align__(16) unsigned char raw[];
  const SegSmem sm = carve(raw, E, wh);
  float* gacc = reinterpret_cast<float*>(sm.rest);                  // [Vs][2]
  float2* wacc_all = reinterpret_cast<float2*>(gacc + (size_t)((Vs * 2 + 3) & ~3));   // [nwarps][kAccSlots]
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < Vs * 2; i += blockDim.x) gacc[i] = 0.f;
  for (int i = threadIdx.x; i < nwarps * kAccSlots; i += blockDim.x) wacc_all[i] = make_float2(0.f, 0.f);
  classify(sm, projects + (size_t)n * Vs * 3, mask + (size_t)n * Vs, ptr, idx, P, wh);
  const int ghead = *sm.ghead;
  const int C = C32 ? 32 : P + 1;
  const bool live = lane >= 1 && lane < C;                          // lane = channel; channel 0 carries the gate
  const int k = live ? lane - 1 : 0;
  const int p0 = sm.pptr[k], p1 = sm.pptr[k + 1], nl = sm.lcount[k];
  const float4* ent_k = sm.ent + p0;
  float2* wacc = wacc_all + (size_t)warp * kAccSlots + sm.lbase[k]; // this lane's part, private to (warp, lane)
  const uint32_t ent_sa = (uint32_t)__cvta_generic_to_shared(ent_k);   // shared-window addresses of the two hot arrays
  const uint32_t wacc_sa = (uint32_t)__cvta_generic_to_shared(wacc);
  const int acc_cap = kAccSlots - sm.lbase[k];                      // light indices >= acc_cap overflow to atomics
  const size_t plane = (size_t)ntiles * NB * 32 * 8;
  const unsigned char* sv = saved + (size_t)n * plane * 4 + (size_t)(lane >> 3) * plane + (lane & 7);
  const float* g_n = g_seg + (size_t)n * wh * wh * C + (lane < C ? lane : 0);
  const int nrec = ntiles * NB * 32;

  // Runs of equal arg-min vertices are merged in registers; a run ends with one load/add/store on the lane's private
  // slot.  Light indices beyond the private capacity (only possible when the sample has more than kAccSlots visible
  // part vertices) fall back to atomics behind a warp-uniform flag, so the common path carries no such branch.
  const bool overflow = (sm.lbase[31] + sm.lcount[31]) > kAccSlots;
  int run_li = -1;                                                  // current run: light index within this lane's part
  float run_u = 0.f, run_v = 0.f;
  auto flush_slow = [&]() {
    if (run_li < acc_cap) {
      float2 a = wacc[run_li];
      a.x += run_u; a.y += run_v;
      wacc[run_li] = a;
    } else {
      const int vid = __float_as_int(ent_k[run_li].w);
      atomicAdd(&gacc[vid * 2], run_u); atomicAdd(&gacc[vid * 2 + 1], run_v);
    }
  };
  // Records follow the forward's order, id = (tile*NB + b)*32 + l32, l32 = ly*4 + lx.  Each warp owns a contiguous
  // range of whole 32-record groups; a batch is 8 records = two lane rows visited in serpentine order (even rows left
  // to right, odd rows right to left), so consecutive records are neighbouring pixel blocks and runs stay long.
  struct Batch { int code[kBatch]; float g[kBatch]; int r0, c0; };
  auto load_batch = [&](int base, Batch& bt) {
    const int l0 = base & 31, tb = base >> 5;
    const int b = tb % NB, t = tb / NB;                             // compile-time divisors
    const int ty = (tiles_x == 1) ? t : t / tiles_x, tx = t - ty * tiles_x;
    bt.r0 = ty * TH + (l0 >> 2) * BH + b / BW;                      // first lane row of the batch; the second is +BH
    bt.c0 = tx * TW + b % BW;                                       // lx = 0 ; +BW per lane column
    const unsigned char* srow = sv + (size_t)base * 8;
    const float* grow = g_n + ((wh - 1 - bt.r0) * wh + bt.c0) * C;  // rows flipped (:68): next lane row is -BH*wh*C
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      const int ly = j >> 2, lx = (j & 4) ? 3 - (j & 3) : (j & 3);  // serpentine
      const bool in = ALLIN || (bt.r0 + ly * BH < wh && bt.c0 + lx * BW < wh);
      bt.code[j] = in ? (int)srow[(ly * 4 + lx) * 8] : 0;
      bt.g[j] = (in && (C32 || lane < C)) ? grow[(lx * BW - ly * BH * wh) * C] : 0.f;
    }
  };
  const int ngroups = nrec >> 5;
  const int grp0 = (int)(((long long)ngroups * warp) / nwarps), grp1 = (int)(((long long)ngroups * (warp + 1)) / nwarps);
  const int rec_end = grp1 * 32;
  Batch cur, nxt;
  int base = grp0 * 32;
  if (base < rec_end) load_batch(base, cur);
  for (; base < rec_end; base += kBatch) {
    if (base + kBatch < rec_end) load_batch(base + kBatch, nxt);    // next batch's loads fly while this one is consumed
    const float gxb = (float)cur.c0, gyb = (float)cur.r0;
    unsigned slowmask = 0u;
#pragma unroll
    for (int h = 0; h < kBatch; h += 4) {
      // (a) four records at a time, branch-free and mutually independent: the compiler interleaves the four chains
      int li4[4];
      float cu4[4], cv4[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = h + jj;
        const int ly = j >> 2, lx = (j & 4) ? 3 - (j & 3) : (j & 3);
        const int gate = __shfl_sync(0xffffffffu, cur.code[j], 0);
        const float g0 = __shfl_sync(0xffffffffu, cur.g[j], 0);
        const float G = cur.g[j] - ((gate & 1) ? g0 : 0.f);         // d bg / d s_k = -gate
        const float gx = gxb + (float)(lx * BW), gy = gyb + (float)(ly * BH);   // small integers: exact in fp32
        int li = live ? cur.code[j] - 1 : -1;                       // 0 -> -1 none
        if (li == 254) { slowmask |= 1u << j; li = -1; }            // code 255: rare exact re-query, deferred
        // light entry (w == 1); lanes without a vertex read slot 0 and contribute zero
        const float2 e = lds_f2(ent_sa + (uint32_t)(li < 0 ? 0 : li) * 16u);
        const float du = __fsub_rn(e.x, gx), dv = __fsub_rn(e.y, gy);
        const float d2 = __fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv));
        const float rs = rsqrt_approx(fmaxf(d2, 1e-30f));
        const float s = ex2_approx((d2 * rs) * (-kLog2e));
        const float coef = (li < 0) ? 0.f : -(s * G) * rs;          // -s (p - g)/d ; d == 0 -> du = dv = 0 -> 0
        li4[jj] = li; cu4[jj] = coef * du; cv4[jj] = coef * dv;
      }
      // (b) sequential run merge over the four results
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const bool brk = li4[jj] != run_li;
        if (brk && run_li >= 0) {
          if (!overflow) {                                          // private slot: this lane is its only writer
            const uint32_t a_sa = wacc_sa + (uint32_t)run_li * 8u;
            float2 a = lds_f2(a_sa);
            a.x += run_u; a.y += run_v;
            sts_f2(a_sa, a);
          } else {
            flush_slow();
          }
        }
        run_u = brk ? cu4[jj] : run_u + cu4[jj];
        run_v = brk ? cv4[jj] : run_v + cv4[jj];
        run_li = li4[jj];
      }
    }
    if (__any_sync(0xffffffffu, slowmask != 0u)) {                  // rare: heavy / generic winners, exact, atomics
#pragma unroll 1
      for (int j = 0; j < kBatch; ++j) {                            // warp-uniform loop: every lane takes the shuffles
        const int ly = j >> 2, lx = (j & 4) ? 3 - (j & 3) : (j & 3);
        const int gate = __shfl_sync(0xffffffffu, cur.code[j], 0);
        const float g0 = __shfl_sync(0xffffffffu, cur.g[j], 0);
        if ((slowmask >> j) & 1u) {
          const float G = cur.g[j] - ((gate & 1) ? g0 : 0.f);
          slow_pixel_grad(sm, idx, gacc, p0, p1, nl, gxb + (float)(lx * BW), gyb + (float)(ly * BH),
                          sm.head[(cur.r0 + ly * BH) * wh + cur.c0 + lx * BW], ghead, G);
        }
      }
    }
    cur = nxt;
  }
  if (run_li >= 0) flush_slow();
  __syncthreads();
  // fold the warps' private slots into the per-vertex sums (a vertex may sit in more than one part)
  const int nlight = min(sm.lbase[31] + sm.lcount[31], kAccSlots);
  for (int li = threadIdx.x; li < nlight; li += blockDim.x) {
    float su = 0.f, sv2 = 0.f;
    for (int w = 0; w < nwarps; ++w) { const float2 a = wacc_all[(size_t)w * kAccSlots + li]; su += a.x; sv2 += a.y; }
    int kk = 0;                                                     // part kk with lbase[kk] <= li < lbase[kk] + lcount[kk]
    while (kk < 31 && li >= sm.lbase[kk + 1]) ++kk;
    const int vid = __float_as_int(sm.ent[sm.pptr[kk] + (li - sm.lbase[kk])].w);
    atomicAdd(&gacc[vid * 2], su); atomicAdd(&gacc[vid * 2 + 1], sv2);
  }
  __syncthreads();
  float* out = g_projects + (size_t)n * Vs * 3;
  for (int i = threadIdx.x; i < Vs * 3; i += blockDim.x) {
    const int v = i / 3, c = i - v * 3;
    out[i] = (c < 2) ? gacc[v * 2 + c] : 0.f;                       // z receives no gradient from the rasteriser
  }
}

constexpr size_t kMaxSmem = 227 * 1024;

template <int BW, int BH, int LX>
cudaError_t launch_fwd_cfg(const SmplB200Parts* p, const float* projects, const float* mask, int N, int Vs, int wh,
                           float* seg, unsigned char* saved, cudaStream_t st) {
  const SegGeom g = seg_geom(wh);
  int warps = 1;
  for (int w = 8; w >= 1; --w)                 // the largest warp count <= 8 that divides the tile count evenly
    if (g.ntiles % w == 0) { warps = w; break; }
  // split a sample's tiles over gridDim.y when the batch alone leaves SMs idle
  int split = 1;
  if (N < 2 * 148) {
    split = max(1, min(g.ntiles, (2 * 148 + N - 1) / N));
    warps = 8;     // every block re-classifies the sample's vertices: keep 8 warps for that even if it owns few tiles
  }
  const size_t smem = seg_base_smem(p->E, wh) + (size_t)warps * 8 * g.NB * 32 * 4;
  if (smem > kMaxSmem) return cudaErrorInvalidConfiguration;
  dim3 grid(N, split);
  LaunchScope scope(KID_SEG_FWD, st);
  cudaError_t e;
  if (saved) {
    e = cudaFuncSetAttribute(seg_fwd_kernel<BW, BH, LX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    seg_fwd_kernel<BW, BH, LX, true><<<grid, warps * 32, smem, st>>>(projects, mask, N, Vs, p->ptr, p->idx, p->P, p->E, wh,
                                                                    seg, saved);
  } else {
    e = cudaFuncSetAttribute(seg_fwd_kernel<BW, BH, LX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    seg_fwd_kernel<BW, BH, LX, false><<<grid, warps * 32, smem, st>>>(projects, mask, N, Vs, p->ptr, p->idx, p->P, p->E, wh,
                                                                     seg, nullptr);
  }
  return cudaGetLastError();
}

}  // namespace

size_t seg_saved_bytes(int N, int wh) {
  const SegGeom g = seg_geom(wh);
  return (size_t)N * g.ntiles * g.NB * 32 * 8 * 4;
}

cudaError_t launch_seg_fwd(const SmplB200Parts* p, const float* projects, const float* mask, int N, int Vs, int wh,
                           float* seg, unsigned char* saved, cudaStream_t st) {
  if (wh % 12 == 0) return launch_fwd_cfg<3, 2, 4>(p, projects, mask, N, Vs, wh, seg, saved, st);
  return launch_fwd_cfg<4, 2, 4>(p, projects, mask, N, Vs, wh, seg, saved, st);
}

cudaError_t launch_seg_bwd(const SmplB200Parts* p, const float* projects, const float* mask, const float* g_seg,
                           const unsigned char* saved, int N, int Vs, int wh, float* g_projects, cudaStream_t st) {
  const int warps = 8;
  const size_t smem = seg_base_smem(p->E, wh) + (size_t)((Vs * 2 + 3) & ~3) * 4 + (size_t)warps * kAccSlots * 8;
  if (smem > kMaxSmem) return cudaErrorInvalidConfiguration;
  const SegGeom g = seg_geom(wh);
  LaunchScope scope(KID_SEG_BWD, st);
#define SMPL_SEG_BWD(BW, BH, LX, C32, ALLIN)                                                                           \
  do {                                                                                                                 \
    cudaError_t e = cudaFuncSetAttribute(seg_bwd_kernel<BW, BH, LX, C32, ALLIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                                    \
    seg_bwd_kernel<BW, BH, LX, C32, ALLIN><<<N, warps * 32, smem, st>>>(projects, mask, g_seg, saved, N, Vs, p->ptr,    \
                                                                        p->idx, p->P, p->E, wh, g.tiles_x, g.ntiles,    \
                                                                        g_projects);                                   \
  } while (0)
  const bool c32 = p->P == 31;
  const bool allin = (wh % g.TW == 0) && (wh % g.TH == 0);          // every record is a pixel of the image
  if (wh % 12 == 0) {
    if (c32 && allin) SMPL_SEG_BWD(3, 2, 4, true, true);
    else if (c32) SMPL_SEG_BWD(3, 2, 4, true, false);
    else SMPL_SEG_BWD(3, 2, 4, false, false);
  } else {
    if (c32 && allin) SMPL_SEG_BWD(4, 2, 4, true, true);
    else if (c32) SMPL_SEG_BWD(4, 2, 4, true, false);
    else SMPL_SEG_BWD(4, 2, 4, false, false);
  }
#undef SMPL_SEG_BWD
  return cudaGetLastError();
}

}  // namespace smplb200
