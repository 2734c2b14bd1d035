// K6: soft silhouette from projected vertices, forward and backward.
//
// Reference arithmetic (keras_smpl/projects_to_silhouette.py:14-44), per pixel g = (column c, row r):
//   s[g] = max_i exp(-||p_i - g||_2 / 1.2)     :35-38   (all vertices, no visibility weights; true division by 1.2f)
//   out[n, wh-1-r, c, :] = [1 - s, s]           :40-42   (rows flipped)
// max_i exp(-d_i/1.2) = exp(-(min_i d_i)/1.2): an exact nearest-vertex query over squared distances computed as
// fl(fl(du^2)+fl(dv^2)).  The reference evaluates all wh^2 x V pairs.  Here every sample is counting-sorted into a grid
// of at most 64x64 cells in shared memory (cells row-major, so 8 consecutive cells of a row -- a "strip" -- are one
// contiguous range of the sorted array) and the image is resolved hierarchically, one warp per top tile (8x8 cells):
//   1. an upper bound U on the distance from the tile centre c to its nearest vertex (nearest non-empty strip, scanned);
//   2. every vertex that is nearest to SOME pixel of the tile lies within U + 2R of c (R = half diagonal), so only strips
//      whose box comes that close are visited: first to find i0 = the vertex nearest c, then to prune EXACTLY --
//      f(g) = d_j^2(g) - d_i0^2(g) is linear in g, so f(c) - 2(|du_j0| hw + |dv_j0| hh) >= margin means j is never the
//      arg-min inside the tile -- and the survivors go to a per-warp list;
//   3. the list is pruned again for each 16x8 sub-tile and, where it is still long, for each 8x4 leaf; the pixels of a
//      sub-tile / leaf are then compared against its survivors only (a handful for background tiles, a few dozen inside
//      the body at 256x256).
// Where the vertices are much denser than the pixels (low resolutions) lists would not shorten anything: those launches,
// and any tile whose survivor list overflows, use the plain search instead -- 8x4 pixel tiles visiting cell rings of
// growing Chebyshev radius until the ring's lower bound (rho-1)*B exceeds the worst current distance in the tile.  All
// lower bounds stay valid for vertices clamped into border cells, so every path returns the exact minimum,
// bit-identical to the brute-force value.
// Backward (TF autodiff): d s/d p_i* = -(1/1.2) s (p_i* - g)/d at the first arg-min (ties: measure zero), 0 when d == 0
// (TF: NaN); upstream is g[...,1] - g[...,0].  When the forward was given a `saved` buffer it records the arg-min vertex
// of every pixel (16 bits, in the output's pixel order) and the backward is a streaming kernel (sil_bwd_saved_kernel:
// no search; runs of pixels that share their vertex -- the whole background -- are summed across the warp before one
// shared-memory add).  Without it the search is repeated.
#include <math_constants.h>
#include <algorithm>
#include "common.cuh"

namespace smplb200 {

namespace {

constexpr int kMaxGrid = 64;        // fine grid: at most kMaxGrid x kMaxGrid cells
constexpr int kStrip = 8;           // cells per strip; a top tile is kStrip x kStrip cells
constexpr int kCap1 = 1536, kCap2 = 512, kCap3 = 256;   // survivor lists per warp (sorted positions, 16 bits each)
#ifndef SIL_LEAFMAX
#define SIL_LEAFMAX 64
#endif
constexpr int kLeafMax = SIL_LEAFMAX;   // a 16x8 sub-tile with at most this many survivors is evaluated directly, a 2 x 2 block per
                                        // lane (measured, ms per 2048 samples at 256x256: 8 -> 6.19, 24 -> 6.04, 64 -> 5.98, 160 -> 6.00)
constexpr float kMargin = 0.01f, kRel = 4e-6f;          // pruning margin: 0.01 px^2 + 4e-6 d^2 (>> fp32 rounding)
constexpr int kSilWarps = 32;   // measured at 256x256, 6890 vertices, ms per 2048 samples (lists kCap1/kCap2/kCap3): 16 warps 2048/512/256
                                // 9.51; 32 warps 1024/512/256 8.26, 1536/512/256 7.51, 1280/384/256 7.52, 1728/384/192 7.90,
                                // 1920/256/128 9.50 (a tile whose list overflows falls back to the ring search: a third of
                                // the kernel's samples at 1024); a 128 x 128 cell grid with 16-pixel top tiles 11.9

struct SilSmem {
  float2* pts;             // [Vs] vertices sorted by cell
  unsigned short* vid;     // [Vs] their vertex ids
  unsigned short* cstart;  // [cells + 1]
  int* scratch;            // [cells + 1] counting / cursors while binning; afterwards the survivor lists live here
  float* gacc;             // [Vs][2] (backward only)
  unsigned short* saved;   // global: [N][wh][wh] arg-min vertex ids in output pixel order (forward only; may be null)
};

struct Grid {
  int B, G, S, wh;         // pixels per cell, cells per side, strips per row
};

__device__ __forceinline__ int cell_coord(float x, float invB, int G) {
  const float f = floorf(x * invB);
  return (int)fminf(fmaxf(f, 0.f), (float)(G - 1));     // NaN -> 0
}

__device__ void bin_vertices(const SilSmem& sm, const float* __restrict__ proj, int Vs, const Grid& gr) {
  const int G = gr.G, cells = G * G;
  const float invB = 1.0f / (float)gr.B;                 // B is a power of two; used for binning only
  for (int i = threadIdx.x; i <= cells; i += blockDim.x) sm.scratch[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < Vs; i += blockDim.x) {
    const int c = cell_coord(proj[i * 3 + 1], invB, G) * G + cell_coord(proj[i * 3], invB, G);
    atomicAdd(&sm.scratch[c + 1], 1);
  }
  __syncthreads();
  if (threadIdx.x < 32) {                                // warp scan over <= 4096 counts
    int carry = 0;
    for (int base = 0; base < cells; base += 32) {
      const int i = base + threadIdx.x;
      int v = (i < cells) ? sm.scratch[i + 1] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if ((int)threadIdx.x >= o) v += t;
      }
      if (i < cells) sm.scratch[i + 1] = v + carry;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= cells; i += blockDim.x) sm.cstart[i] = (unsigned short)sm.scratch[i];
  __syncthreads();
  for (int i = threadIdx.x; i < Vs; i += blockDim.x) {
    const float u = proj[i * 3], v = proj[i * 3 + 1];
    const int c = cell_coord(v, invB, G) * G + cell_coord(u, invB, G);
    const int pos = atomicAdd(&sm.scratch[c], 1);        // scratch[c] starts at the cell's first slot
    sm.pts[pos] = make_float2(u, v);
    sm.vid[pos] = (unsigned short)i;
  }
  __syncthreads();
}

__device__ __forceinline__ float sq_dist(float u, float v, float gx, float gy) {
  const float du = __fsub_rn(u, gx), dv = __fsub_rn(v, gy);
  return __fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv));
}

// squared distance from (cx, cy) to the pixel box of strip (row, sx); a lower bound for every vertex binned there
// (vertices outside the image are clamped INTO border cells, which only moves their cell closer)
__device__ __forceinline__ float strip_lb2(const Grid& gr, int row, int sx, float cx, float cy) {
  const float x0 = (float)(sx * kStrip * gr.B), x1 = (float)(min((sx + 1) * kStrip, gr.G) * gr.B);
  const float y0 = (float)(row * gr.B), y1 = (float)((row + 1) * gr.B);
  const float dx = fmaxf(fmaxf(x0 - cx, cx - x1), 0.f), dy = fmaxf(fmaxf(y0 - cy, cy - y1), 0.f);
  return dx * dx + dy * dy;
}
__device__ __forceinline__ void strip_range(const SilSmem& sm, const Grid& gr, int row, int sx, int& a, int& b) {
  a = sm.cstart[row * gr.G + sx * kStrip];
  b = sm.cstart[row * gr.G + min((sx + 1) * kStrip, gr.G)];
}

// ---------------------------------------------------------------------------------------------------------------
// plain search: nearest vertex for this lane's pixel of an 8x4 tile; the whole warp walks cell rings together
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ring_search(const SilSmem& sm, const Grid& gr, int cx0, int cx1, int cy0, int cy1,
                                            bool active, float gx, float gy, float& best, int& barg) {
  const int lane = threadIdx.x & 31, G = gr.G, B = gr.B;
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
      if (inr) { s0 = sm.cstart[y * G + x]; s1 = sm.cstart[y * G + x + 1]; }
      unsigned todo = __ballot_sync(0xffffffffu, s1 > s0);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int a = __shfl_sync(0xffffffffu, s0, src), b = __shfl_sync(0xffffffffu, s1, src);
        for (int i = a; i < b; ++i) {
          const float2 p = sm.pts[i];                    // broadcast
          const float d2 = sq_dist(p.x, p.y, gx, gy);
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

// ---------------------------------------------------------------------------------------------------------------
// hierarchical search
// ---------------------------------------------------------------------------------------------------------------
// The warp's nearest (d, u, v): one integer redux on the distance bits (d >= 0 or +inf: the bit patterns order like the
// values), then the lowest lane that holds the minimum hands out its vertex -- every lane ends with the SAME vertex, which
// is all the pruning needs of its reference (any vertex is a valid one; a nearer one prunes more).  (A five-round butterfly
// over (d, u, v) with a lexicographic tie rule was 6.5 % of the kernel's samples.)
__device__ __forceinline__ void warp_argmin(float& d, float& u, float& v) {
  const unsigned bits = __float_as_uint(d);
  const unsigned mn = __reduce_min_sync(0xffffffffu, bits);
  const int src = __ffs(__ballot_sync(0xffffffffu, bits == mn)) - 1;
  d = __uint_as_float(mn);
  u = __shfl_sync(0xffffffffu, u, src);
  v = __shfl_sync(0xffffffffu, v, src);
}

// Upper bound (squared) on the distance from (cx, cy) to its nearest vertex: the nearest non-empty strip is scanned.
__device__ float probe_upper2(const SilSmem& sm, const Grid& gr, float cx, float cy) {
  const int lane = threadIdx.x & 31, nstrips = gr.G * gr.S;
  float blb = CUDART_INF_F;
  int bs = -1;
  {                                                      // the strip that holds the centre has bound 0: if it is not empty, it is the one
    const int row = min(max((int)(cy / (float)gr.B), 0), gr.G - 1), sx = min(max((int)(cx / (float)(kStrip * gr.B)), 0), gr.S - 1);
    int a, b;
    strip_range(sm, gr, row, sx, a, b);
    if (b > a && strip_lb2(gr, row, sx, cx, cy) == 0.f) { blb = 0.f; bs = row * gr.S + sx; }
  }
  const bool centre_strip = bs >= 0;                     // warp-uniform
  for (int t = lane; !centre_strip && t < nstrips; t += 32) {
    const int row = t / gr.S, sx = t - row * gr.S;
    int a, b;
    strip_range(sm, gr, row, sx, a, b);
    const float lb = strip_lb2(gr, row, sx, cx, cy);
    if (b > a && lb < blb) { blb = lb; bs = t; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float l2 = __shfl_xor_sync(0xffffffffu, blb, o);
    const int s2 = __shfl_xor_sync(0xffffffffu, bs, o);
    if (l2 < blb || (l2 == blb && s2 > bs)) { blb = l2; bs = s2; }
  }
  if (bs < 0) return CUDART_INF_F;                       // no vertices at all
  int a, b;
  strip_range(sm, gr, bs / gr.S, bs % gr.S, a, b);
  float u2 = CUDART_INF_F;
  for (int i = a + lane; i < b; i += 32) {
    const float2 p = sm.pts[i];
    u2 = fminf(u2, sq_dist(p.x, p.y, cx, cy));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) u2 = fminf(u2, __shfl_xor_sync(0xffffffffu, u2, o));
  return u2;
}

// Visit every vertex within sqrt(rc2) of (cx, cy) (and some more): f(sorted position, point), lanes strided.  Cells are
// row-major, so the cells of one grid row that the disc can reach are ONE contiguous range of the sorted array: lane = row,
// its chord of the disc -> first and last cell -> [cstart, cstart).  (Whole 8-cell strips, tested one by one against the
// disc, visited about twice the vertices: a strip is 32 pixels wide at 256 x 256, the disc of a 16 x 8 region 36.)
// Vertices outside the image are clamped INTO the border cells, which only moves their cell closer: the bounds hold.
template <typename F>
__device__ __forceinline__ void for_each_candidate(const SilSmem& sm, const Grid& gr, float cx, float cy, float rc2, F f) {
  const int lane = threadIdx.x & 31;
  rc2 = rc2 * 1.0001f + 0.01f;                           // the cell bounds and the vertex distances round differently
  const float rc = sqrtf(rc2) + 1.0f;
  const float fB = (float)gr.B, invB = 1.0f / fB;        // B is a power of two
  const int r_lo = max(0, (int)floorf((cy - rc) * invB)), r_hi = min(gr.G - 1, (int)floorf((cy + rc) * invB));
  for (int base = r_lo; base <= r_hi; base += 32) {
    const int row = base + lane;
    int a = 0, b = 0;
    if (row <= r_hi) {
      const float y0 = (float)row * fB, y1 = y0 + fB;                // the row's band of pixels
      const float dy = fmaxf(fmaxf(y0 - cy, cy - y1), 0.f);
      const float rem = rc2 - dy * dy;
      if (rem >= 0.f) {
        const float dx = sqrtf(rem) * 1.0001f + 0.01f;
        const int c_lo = min(max((int)floorf((cx - dx) * invB), 0), gr.G - 1);
        const int c_hi = min(max((int)floorf((cx + dx) * invB), 0), gr.G - 1);
        a = sm.cstart[row * gr.G + c_lo];
        b = sm.cstart[row * gr.G + c_hi + 1];
      }
    }
    unsigned todo = __ballot_sync(0xffffffffu, b > a);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int sa = __shfl_sync(0xffffffffu, a, src), sb = __shfl_sync(0xffffffffu, b, src);
      for (int i0 = sa; i0 < sb; i0 += 32) {             // warp-uniform trip count: f may use warp collectives
        const int i = i0 + lane;
        const bool in = i < sb;
        const float2 p = in ? sm.pts[i] : make_float2(0.f, 0.f);
        f(i, p, in);
      }
    }
  }
}

// dominance test of vertex p against the reference vertex (u0, v0, its squared distance bd to the region centre)
__device__ __forceinline__ bool survives(float2 p, float cx, float cy, float hw, float hh, float u0, float v0, float bd) {
  const float du = p.x - cx, dv = p.y - cy;
  const float d = du * du + dv * dv;
  const float slack = 2.0f * (fabsf(p.x - u0) * hw + fabsf(p.y - v0) * hh);
  return (d - bd) - slack < kMargin + kRel * d;
}

// in[0..n) -> out: the vertices that can be nearest to some pixel of the region centred (cx, cy), half extents (hw, hh).
// Returns the survivor count, or -1 if it exceeds cap (the caller then works with the unpruned list).
__device__ int prune_list(const SilSmem& sm, const unsigned short* in, int n, float cx, float cy, float hw, float hh,
                          unsigned short* out, int cap) {
  const int lane = threadIdx.x & 31;
  float bd = CUDART_INF_F, u0 = 0.f, v0 = 0.f;
  for (int i = lane; i < n; i += 32) {
    const float2 p = sm.pts[in[i]];
    const float du = p.x - cx, dv = p.y - cy;
    const float d = du * du + dv * dv;
    if (d < bd) { bd = d; u0 = p.x; v0 = p.y; }
  }
  warp_argmin(bd, u0, v0);
  int m = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    unsigned short id = 0;
    bool keep = false;
    if (i < n) {
      id = in[i];
      keep = survives(sm.pts[id], cx, cy, hw, hh, u0, v0, bd);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int pos = m + __popc(bal & ((1u << lane) - 1u));
    if (keep && pos < cap) out[pos] = id;
    m += __popc(bal);
  }
  __syncwarp();
  return m <= cap ? m : -1;
}

// nearest vertex among list[0..n) for NPX pixels of this lane (gx[i], gy[j] as a (NPX == 4 ? 2 x 2 : 1 x 1) block)
template <int NX, int NY>
__device__ __forceinline__ void eval_list(const SilSmem& sm, const unsigned short* list, int n, const float (&gx)[NX],
                                          const float (&gy)[NY], float (&best)[NX * NY], int (&barg)[NX * NY]) {
#pragma unroll
  for (int q = 0; q < NX * NY; ++q) { best[q] = CUDART_INF_F; barg[q] = -1; }
  for (int i = 0; i < n; ++i) {
    const int id = list[i];                              // broadcast
    const float2 p = sm.pts[id];
    float ux[NX], vy[NY];
#pragma unroll
    for (int a = 0; a < NX; ++a) { const float d = __fsub_rn(p.x, gx[a]); ux[a] = __fmul_rn(d, d); }
#pragma unroll
    for (int b = 0; b < NY; ++b) { const float d = __fsub_rn(p.y, gy[b]); vy[b] = __fmul_rn(d, d); }
#pragma unroll
    for (int b = 0; b < NY; ++b)
#pragma unroll
      for (int a = 0; a < NX; ++a) {
        const float d2 = __fadd_rn(ux[a], vy[b]);
        const int q = b * NX + a;
        if (d2 < best[q]) { best[q] = d2; barg[q] = id; }
      }
  }
}

// one pixel: score, store (forward) or gradient contribution (backward)
constexpr unsigned short kSilNone = 0xffffu;   // saved arg-min of a pixel with no vertex at all

template <bool BWD>
__device__ __forceinline__ void finish_pixel(const SilSmem& sm, int n, int wh, int c, int r, bool active, float best,
                                             int barg, const float* __restrict__ g_sil, float* __restrict__ out,
                                             float2& fwd_val, float& cu, float& cv, int& vid) {
  const float d = sqrtf(best);
  const float s = expf(__fdiv_rn(-d, 1.2f));             // tf.exp(tf.negative(norm) / 1.2) (:37)
  fwd_val = make_float2(1.0f - s, s);
  cu = 0.f; cv = 0.f; vid = -1;
  if (!BWD && active && sm.saved)                         // forward: the arg-min vertex, for the search-free backward
    sm.saved[((size_t)n * wh + (wh - 1 - r)) * wh + c] = barg >= 0 ? sm.vid[barg] : kSilNone;
  if (BWD && active && barg >= 0) {
    const size_t o = (((size_t)n * wh + (wh - 1 - r)) * wh + c) * 2;
    const float2 g = *reinterpret_cast<const float2*>(g_sil + o);
    const float2 p = sm.pts[barg];
    const float du = __fsub_rn(p.x, (float)c), dv = __fsub_rn(p.y, (float)r);
    const float coef = (d > 0.f) ? (-(g.y - g.x) * s / 1.2f) / d : 0.f;
    cu = coef * du; cv = coef * dv;
    vid = sm.vid[barg];
  }
}

// backward accumulation of one lane's NPX contributions; pixels of a lane that share the vertex are merged first
template <int NPX>
__device__ __forceinline__ void accumulate(const SilSmem& sm, const int (&vid)[NPX], const float (&cu)[NPX],
                                           const float (&cv)[NPX]) {
  bool done[NPX];
#pragma unroll
  for (int q = 0; q < NPX; ++q) done[q] = vid[q] < 0;
#pragma unroll
  for (int q = 0; q < NPX; ++q) {
    if (!done[q]) {
      float su = cu[q], sv = cv[q];
#pragma unroll
      for (int t = q + 1; t < NPX; ++t)
        if (!done[t] && vid[t] == vid[q]) { su += cu[t]; sv += cv[t]; done[t] = true; }
      atomicAdd(&sm.gacc[vid[q] * 2], su);
      atomicAdd(&sm.gacc[vid[q] * 2 + 1], sv);
    }
  }
}

// 8x4 leaf at pixel origin (lx0, ly0): one pixel per lane, compared against list[0..n) (n < 0: plain ring search)
template <bool BWD>
__device__ __forceinline__ void do_leaf(const SilSmem& sm, const Grid& gr, int n_img, int lx0, int ly0,
                                        const unsigned short* list, int n, const float* __restrict__ g_sil,
                                        float* __restrict__ out) {
  const int lane = threadIdx.x & 31, wh = gr.wh;
  const int c = lx0 + (lane & 7), r = ly0 + (lane >> 3);
  const bool active = c < wh && r < wh;
  float best[1];
  int barg[1];
  if (n >= 0) {
    const float gx[1] = {(float)c}, gy[1] = {(float)r};
    eval_list<1, 1>(sm, list, n, gx, gy, best, barg);
  } else {
    const int c1 = min(lx0 + 8, wh) - 1, r1 = min(ly0 + 4, wh) - 1;
    ring_search(sm, gr, min(lx0 / gr.B, gr.G - 1), min(c1 / gr.B, gr.G - 1), min(ly0 / gr.B, gr.G - 1),
                min(r1 / gr.B, gr.G - 1), active, (float)c, (float)r, best[0], barg[0]);
  }
  float2 val;
  float cu[1], cv[1];
  int vid[1];
  finish_pixel<BWD>(sm, n_img, wh, c, r, active, best[0], barg[0], g_sil, out, val, cu[0], cv[0], vid[0]);
  if (!BWD) {
    if (active) *reinterpret_cast<float2*>(out + (((size_t)n_img * wh + (wh - 1 - r)) * wh + c) * 2) = val;
  } else {
    accumulate<1>(sm, vid, cu, cv);
  }
}

// 16x8 sub-tile at pixel origin (sx0, sy0): a 2x2 block per lane, compared against list[0..n)
template <bool BWD>
__device__ __forceinline__ void do_subtile_direct(const SilSmem& sm, const Grid& gr, int n_img, int sx0, int sy0,
                                                  const unsigned short* list, int n, const float* __restrict__ g_sil,
                                                  float* __restrict__ out) {
  const int lane = threadIdx.x & 31, wh = gr.wh;
  const int c0 = sx0 + (lane & 7) * 2, r0 = sy0 + (lane >> 3) * 2;
  const float gx[2] = {(float)c0, (float)(c0 + 1)}, gy[2] = {(float)r0, (float)(r0 + 1)};
  float best[4];
  int barg[4];
  eval_list<2, 2>(sm, list, n, gx, gy, best, barg);
  float2 val[4];
  float cu[4], cv[4];
  int vid[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int c = c0 + (q & 1), r = r0 + (q >> 1);
    finish_pixel<BWD>(sm, n_img, wh, c, r, c < wh && r < wh, best[q], barg[q], g_sil, out, val[q], cu[q], cv[q], vid[q]);
  }
  if (!BWD) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r = r0 + j;
      if (r < wh) {
        float* o = out + (((size_t)n_img * wh + (wh - 1 - r)) * wh + c0) * 2;
        if ((wh & 1) == 0 && c0 + 1 < wh) {             // even width: (c0, c0+1) is one aligned 16-byte store
          *reinterpret_cast<float4*>(o) = make_float4(val[j * 2].x, val[j * 2].y, val[j * 2 + 1].x, val[j * 2 + 1].y);
        } else {
          if (c0 < wh) *reinterpret_cast<float2*>(o) = val[j * 2];
          if (c0 + 1 < wh) *reinterpret_cast<float2*>(o + 2) = val[j * 2 + 1];
        }
      }
    }
  } else {
    accumulate<4>(sm, vid, cu, cv);
  }
}

__device__ __forceinline__ SilSmem carve_sil(unsigned char* raw, int Vs, int cells, int nwarps, bool bwd) {
  SilSmem sm;
  size_t off = 0;
  sm.pts = reinterpret_cast<float2*>(raw + off); off += (size_t)Vs * 8;
  const size_t sa = (size_t)(cells + 1) * 4, sb = (size_t)nwarps * (kCap1 + kCap2 + kCap3) * 2;
  const size_t scratch = sa > sb ? sa : sb;
  sm.scratch = reinterpret_cast<int*>(raw + off); off += (scratch + 15) & ~(size_t)15;
  sm.gacc = bwd ? reinterpret_cast<float*>(raw + off) : nullptr; off += bwd ? (size_t)Vs * 8 : 0;
  sm.vid = reinterpret_cast<unsigned short*>(raw + off); off += ((size_t)Vs * 2 + 15) & ~(size_t)15;
  sm.cstart = reinterpret_cast<unsigned short*>(raw + off);
  sm.saved = nullptr;
  return sm;
}
size_t sil_smem_bytes(int Vs, int G, int nwarps, bool bwd) {
  const size_t scratch = std::max((size_t)(G * G + 1) * 4, (size_t)nwarps * (kCap1 + kCap2 + kCap3) * 2);
  return (size_t)Vs * 8 + ((scratch + 15) & ~(size_t)15) + (bwd ? (size_t)Vs * 8 : 0) + (((size_t)Vs * 2 + 15) & ~(size_t)15) +
         (size_t)(G * G + 1) * 2 + 16;
}

template <bool BWD>
__global__ void __launch_bounds__(kSilWarps * 32, 1)
sil_kernel(const float* __restrict__ projects, const float* __restrict__ g_sil, int N, int Vs, int wh, int B, int G,
           int dense, float* __restrict__ out, unsigned short* __restrict__ saved) {
  extern __shared__ __align__(16) unsigned char raw[];
  __shared__ int next_tile, n_items;
  __shared__ unsigned short items[8 * 64];                // work items: tile | part << 8 (part 0..7 = an eighth of the tile, 15 = all)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  SilSmem sm = carve_sil(raw, Vs, G * G, nwarps, BWD);
  sm.saved = BWD ? nullptr : saved;
  const int n = blockIdx.x;
  Grid gr;
  gr.B = B; gr.G = G; gr.S = (G + kStrip - 1) / kStrip; gr.wh = wh;
  if (BWD)
    for (int i = threadIdx.x; i < Vs * 2; i += blockDim.x) sm.gacc[i] = 0.f;
  if (threadIdx.x == 0) next_tile = nwarps;               // first item handed out on demand (see the loop); ordered before
                                                          // its first use by the barriers below
  bin_vertices(sm, projects + (size_t)n * Vs * 3, Vs, gr);
  // from here on `scratch` holds the per-warp survivor lists
  unsigned short* L1 = reinterpret_cast<unsigned short*>(sm.scratch) + (size_t)warp * (kCap1 + kCap2 + kCap3);
  unsigned short* L2 = L1 + kCap1;
  unsigned short* L3 = L2 + kCap2;
  const int TT = kStrip * B;                              // top tile side in pixels (a multiple of 16)
  const int ttx = (wh + TT - 1) / TT, ntop = ttx * ttx;
  const int t0 = (int)(((long long)ntop * blockIdx.y) / gridDim.y);
  const int t1 = (int)(((long long)ntop * (blockIdx.y + 1)) / gridDim.y);
  // Work items, handed to the warps on demand.  A top tile is one item, except where the body is: a tile that holds more
  // than kHeavyTile vertices becomes eight items (a quarter of its rows x half its columns: one 16 x 8 sub-tile at 256 x 256),
  // each pruned and searched as a region of its own (its lists are shorter than the whole tile's, so little is redone).  Heavy items come first, then the light tiles,
  // each group from the image's centre outwards.  (History: a static t -> warp map with t = row * ttx + column gave warp w
  // the SAME column in every row -- all of the body to two or three of the sixteen warps, 14.7 % warp occupancy in ncu
  // against the 25 % one block per SM allowed; whole tiles on demand, centre first, left 64 items for 32 warps and the
  // block waiting for its few interior tiles.)
  if (warp == 0) {
#ifndef SIL_HEAVY
#define SIL_HEAVY 10
#endif
    constexpr int kHeavyTile = SIL_HEAVY;
    const bool can_split = !dense && TT >= 32;
    int cnt = 0;
    for (int pass = 0; pass < 2; ++pass) {                // 0: heavy tiles (four items each), 1: the others
      for (int k0 = t0; k0 < t1; k0 += 32) {
        const int k = k0 + lane;
        bool take = false;
        int tile = 0;
        if (k < t1) {
          const int ka = k / ttx, kb = k - ka * ttx;
          const int tyi = (ttx >> 1) + ((ka & 1) ? -((ka + 1) >> 1) : (ka >> 1));
          const int txi = (ttx >> 1) + ((kb & 1) ? -((kb + 1) >> 1) : (kb >> 1));
          tile = tyi * ttx + txi;
          int nv = 0;                                     // vertices binned into the tile's own cells
          const int cx0 = txi * kStrip, cx1 = min(cx0 + kStrip, G);
          for (int row = tyi * kStrip; row < min(tyi * kStrip + kStrip, G); ++row)
            nv += (int)sm.cstart[row * G + cx1] - (int)sm.cstart[row * G + cx0];
          const bool heavy = can_split && nv > kHeavyTile;
          take = heavy == (pass == 0);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, take);
        const int per = pass == 0 ? 8 : 1;
        if (take) {
          const int pos = cnt + per * __popc(bal & ((1u << lane) - 1u));
          if (pass == 0) { for (int q = 0; q < 8; ++q) items[pos + q] = (unsigned short)(tile | (q << 8)); }
          else items[pos] = (unsigned short)(tile | (15 << 8));
        }
        cnt += per * __popc(bal);
      }
    }
    if (lane == 0) n_items = cnt;
  }
  __syncthreads();
  const int nit = n_items;
  for (int k = warp; k < nit;) {
    const int item = items[k];
    {
      int nk = 0;
      if (lane == 0) nk = atomicAdd(&next_tile, 1);
      k = __shfl_sync(0xffffffffu, nk, 0);
    }
    const int tile = item & 0xff, band = item >> 8;
    const int tyi = tile / ttx, txi = tile - tyi * ttx;
    int tx0 = txi * TT;
    int tw = min(TT, wh - tx0);
    int ty0 = tyi * TT, th = min(TT, wh - ty0);                  // the part of the tile inside the image ...
    if (band != 15) {                                            // ... or one eighth of it: a quarter of its rows, half its columns
      const int y_lo = ty0 + (band >> 1) * (TT / 4), y_hi = min(y_lo + TT / 4, ty0 + th);
      const int x_lo = tx0 + (band & 1) * (TT / 2), x_hi = min(x_lo + TT / 2, tx0 + tw);
      if (y_lo >= y_hi || x_lo >= x_hi) continue;
      ty0 = y_lo; th = y_hi - y_lo; tx0 = x_lo; tw = x_hi - x_lo;
    }
    int n1 = -1;
    if (!dense) {
      const float hw = 0.5f * (float)(tw - 1), hh = 0.5f * (float)(th - 1);
      const float cx = (float)tx0 + hw, cy = (float)ty0 + hh;
      const float R = sqrtf(hw * hw + hh * hh);
      const float U2 = probe_upper2(sm, gr, cx, cy);
      if (U2 < CUDART_INF_F) {
        // pass A: the vertex nearest the centre, among everything within U of it
        float bd = CUDART_INF_F, u0 = 0.f, v0 = 0.f;
        for_each_candidate(sm, gr, cx, cy, U2, [&](int, float2 p, bool in) {
          const float du = p.x - cx, dv = p.y - cy;
          const float d = du * du + dv * dv;
          if (in && d < bd) { bd = d; u0 = p.x; v0 = p.y; }
        });
        warp_argmin(bd, u0, v0);
        // pass B: exact pruning of everything within sqrt(bd) + 2R, survivors to L1
        const float rc = sqrtf(bd) + 2.0f * R;
        int m = 0;
        for_each_candidate(sm, gr, cx, cy, rc * rc, [&](int i, float2 p, bool in) {
          const bool keep = in && survives(p, cx, cy, hw, hh, u0, v0, bd);
          const unsigned bal = __ballot_sync(0xffffffffu, keep);
          const int pos = m + __popc(bal & ((1u << lane) - 1u));
          if (keep && pos < kCap1) L1[pos] = (unsigned short)i;
          m += __popc(bal);
        });
        __syncwarp();
        n1 = m <= kCap1 ? m : -1;
      } else {
        n1 = 0;                                            // no vertices: every score is exp(-inf) = 0
      }
    }
    // 16x8 sub-tiles of the top tile
    for (int sy0 = ty0; sy0 < ty0 + th; sy0 += 8) {
      for (int sx0 = tx0; sx0 < tx0 + tw; sx0 += 16) {
        int n2 = -1;
        const unsigned short* list2 = L1;
        if (n1 >= 0 && tw <= 16 && th <= 8) {
          n2 = n1;                                        // the region IS this sub-tile: L1 was pruned for exactly these pixels
        } else if (n1 >= 0) {
          const int sw = min(16, wh - sx0), sh = min(8, wh - sy0);
          const float shw = 0.5f * (float)(sw - 1), shh = 0.5f * (float)(sh - 1);
          n2 = prune_list(sm, L1, n1, (float)sx0 + shw, (float)sy0 + shh, shw, shh, L2, kCap2);
          if (n2 >= 0) list2 = L2; else n2 = n1;          // L2 overflow: keep working from L1
        }
        if (n2 >= 0 && n2 <= kLeafMax) {
          do_subtile_direct<BWD>(sm, gr, n, sx0, sy0, list2, n2, g_sil, out);
        } else {
          for (int ly0 = sy0; ly0 < min(sy0 + 8, wh); ly0 += 4) {
            for (int lx0 = sx0; lx0 < min(sx0 + 16, wh); lx0 += 8) {
              int n3 = -1;
              const unsigned short* list3 = list2;
              if (n2 >= 0) {
                const int lw = min(8, wh - lx0), lh = min(4, wh - ly0);
                const float lhw = 0.5f * (float)(lw - 1), lhh = 0.5f * (float)(lh - 1);
                n3 = prune_list(sm, list2, n2, (float)lx0 + lhw, (float)ly0 + lhh, lhw, lhh, L3, kCap3);
                if (n3 >= 0) list3 = L3; else n3 = n2;
              }
              do_leaf<BWD>(sm, gr, n, lx0, ly0, list3, n3, g_sil, out);
            }
          }
        }
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

// Search-free backward from the forward's saved arg-min map.  One block per (sample, slice of output rows); the sample's
// (u, v) and the per-vertex sums live in shared memory.  A warp takes 32 consecutive pixels of the output; lanes whose
// neighbours hold the same vertex form a run (the background is one run per hull vertex), runs are summed by a segmented
// shuffle reduction and only the run heads touch the accumulators.
__global__ void __launch_bounds__(512)
sil_bwd_saved_kernel(const float* __restrict__ projects, const float* __restrict__ g_sil,
                     const unsigned short* __restrict__ saved, int N, int Vs, int wh, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char raw[];
  float2* pts = reinterpret_cast<float2*>(raw);                    // [Vs]
  float* gacc = reinterpret_cast<float*>(raw + (size_t)Vs * 8);    // [Vs][2]
  const int n = blockIdx.x, lane = threadIdx.x & 31;
  const float* proj = projects + (size_t)n * Vs * 3;
  for (int i = threadIdx.x; i < Vs; i += blockDim.x) {
    pts[i] = make_float2(proj[i * 3], proj[i * 3 + 1]);
    gacc[2 * i] = 0.f; gacc[2 * i + 1] = 0.f;
  }
  __syncthreads();
  const int npx = wh * wh;
  const int p0 = (int)(((long long)npx * blockIdx.y) / gridDim.y) & ~31;                 // whole warps' worth of pixels
  const int p1 = (blockIdx.y + 1 == gridDim.y) ? npx : ((int)(((long long)npx * (blockIdx.y + 1)) / gridDim.y) & ~31);
  const unsigned short* sv = saved + (size_t)n * npx;
  const float2* g2 = reinterpret_cast<const float2*>(g_sil) + (size_t)n * npx;
  // Four 32-pixel groups per trip, their loads (2 + 8 bytes per pixel) issued before any is used: with one group per trip
  // and 32 warps per SM the kernel sat on the latency of these loads (ncu: half of its stall samples on their first use).
  constexpr int kU = 8;
  const float inv_wh = 1.0f / (float)wh;
  for (int base = p0 + (threadIdx.x & ~31); base < p1; base += kU * blockDim.x) {
    unsigned vids[kU];
    float2 gs[kU];
#pragma unroll
    for (int k = 0; k < kU; ++k) {
      const int i = base + k * blockDim.x + lane;
      const bool in = i < p1;
      vids[k] = in ? sv[i] : kSilNone;
      gs[k] = in ? g2[i] : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < kU; ++k) {
      if (base + k * (int)blockDim.x >= p1) break;                   // warp-uniform
      const int i = base + k * blockDim.x + lane;
      const unsigned vid = vids[k];
      float cu = 0.f, cv = 0.f;
      if (vid != kSilNone) {
        const float2 g = gs[k];
        int ro = (int)((float)i * inv_wh), c = i - ro * wh;          // i / wh for i < 2^24: the float quotient is off by at most one
        if (c < 0) { --ro; c += wh; } else if (c >= wh) { ++ro; c -= wh; }
        const int r = wh - 1 - ro;                                   // output row -> grid row (:42)
        const float2 p = pts[vid];
        const float du = __fsub_rn(p.x, (float)c), dv = __fsub_rn(p.y, (float)r);
        const float d = sqrtf(__fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv)));
        const float s = expf(__fdiv_rn(-d, 1.2f));
        const float coef = (d > 0.f) ? (-(g.y - g.x) * s / 1.2f) / d : 0.f;
        cu = coef * du; cv = coef * dv;
      }
      // runs of equal vertex ids among consecutive lanes
      const unsigned prev = __shfl_up_sync(0xffffffffu, vid, 1);
      const bool head = lane == 0 || prev != vid;
      const unsigned heads = __ballot_sync(0xffffffffu, head);
      const unsigned above = heads & ~((2u << lane) - 1u);           // heads strictly after this lane
      const int end = above ? (__ffs(above) - 2) : 31;               // last lane of my run
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float tu = __shfl_down_sync(0xffffffffu, cu, o), tv = __shfl_down_sync(0xffffffffu, cv, o);
        if (lane + o <= end) { cu += tu; cv += tv; }
      }
      if (head && vid != kSilNone) {
        atomicAdd(&gacc[2 * vid], cu);
        atomicAdd(&gacc[2 * vid + 1], cv);
      }
    }
  }
  __syncthreads();
  float* gp = out + (size_t)n * Vs * 3;
  for (int i = threadIdx.x; i < Vs * 2; i += blockDim.x) {
    const int v = i >> 1, k = i & 1;
    if (gridDim.y == 1) gp[v * 3 + k] = gacc[i];
    else if (gacc[i] != 0.f) atomicAdd(&gp[v * 3 + k], gacc[i]);
  }
  if (gridDim.y == 1)
    for (int i = threadIdx.x; i < Vs; i += blockDim.x) gp[i * 3 + 2] = 0.f;
}

void sil_grid(int wh, int& B, int& G) {
  B = 2;
  while ((wh + B - 1) / B > kMaxGrid) B *= 2;
  G = (wh + B - 1) / B;
}

template <bool BWD>
cudaError_t launch_sil(const float* projects, const float* g_sil, int N, int Vs, int wh, float* out, unsigned short* saved,
                       cudaStream_t st) {
  if (Vs >= 65535) return cudaErrorInvalidValue;         // sorted positions and vertex ids are kept as 16 bits
  if (BWD && saved) {                                    // search-free backward
    if (wh > 4096) return cudaErrorInvalidValue;         // pixel indices are split into (row, column) through a float quotient: < 2^24
    const size_t smem = (size_t)Vs * 16;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(sil_bwd_saved_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // enough blocks to fill the GPU (two per SM fit at 6890 vertices): slices of the image when the batch is small
    int split = 1;
    const int npx32 = (wh * wh + 31) / 32;
    if (N < 2 * 148) split = max(1, min(npx32, (2 * 148 + N - 1) / N));
    if (split > 1) {
      e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N * Vs * 3, st);
      if (e != cudaSuccess) return e;
    }
    LaunchScope scope(KID_SIL_BWD, st);
    sil_bwd_saved_kernel<<<dim3(N, split), 512, smem, st>>>(projects, g_sil, saved, N, Vs, wh, out);
    return cudaGetLastError();
  }
  int B, G;
  sil_grid(wh, B, G);
  // Survivor lists only pay where pixels outnumber vertices; denser launches use the plain ring search throughout.
  const int dense = (double)Vs > 0.25 * (double)wh * (double)wh ? 1 : 0;
  int warps = kSilWarps;
  while (warps > 4 && sil_smem_bytes(Vs, G, warps, BWD) > 227 * 1024) warps -= 2;
  const size_t smem = sil_smem_bytes(Vs, G, warps, BWD);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(sil_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int TT = kStrip * B, ttx = (wh + TT - 1) / TT, ntop = ttx * ttx;
  int split = 1;
  if (N < 2 * 148) split = max(1, min(ntop, (2 * 148 + N - 1) / N));
  if (BWD && split > 1) {
    e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N * Vs * 3, st);
    if (e != cudaSuccess) return e;
  }
  dim3 grid(N, split);
  LaunchScope scope(BWD ? KID_SIL_BWD : KID_SIL_FWD, st);
  sil_kernel<BWD><<<grid, warps * 32, smem, st>>>(projects, g_sil, N, Vs, wh, B, G, dense, out, saved);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_sil_fwd(const float* projects, int N, int Vs, int wh, float* sil, unsigned short* saved, cudaStream_t st) {
  return launch_sil<false>(projects, nullptr, N, Vs, wh, sil, saved, st);
}

cudaError_t launch_sil_bwd(const float* projects, const float* g_sil, int N, int Vs, int wh, float* g_projects,
                           const unsigned short* saved, cudaStream_t st) {
  return launch_sil<true>(projects, g_sil, N, Vs, wh, g_projects, const_cast<unsigned short*>(saved), st);
}

}  // namespace smplb200
