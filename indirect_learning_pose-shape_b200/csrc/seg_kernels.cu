// K5: 31-part soft segmentation from projected vertices and the visibility weights, forward and backward.
//
// Reference arithmetic (keras_smpl/projects_to_seg.py:9-69), per part k and pixel g = (column c, row r):
//   s_k[g] = max_i exp(-(||p_i - g||_2 * w_i))        :52-56
//   bg[g]  = 1 - clip(sum_k s_k[g], 0, 1)             :61-64
//   out[n, wh-1-r, c, :] = [bg, s_0 .. s_30]          :66-68   (rows flipped)
//
// exp, sqrt and the multiply by w are monotone, so  max_i exp(-(d_i w_i)) = exp(-min_i (d_i w_i)): both directions are
// an exact weighted-nearest-vertex query: the arg-min over fp32 squared distances, everywhere with tf.norm's own
// roundings, fl(fl(du^2) + fl(dv^2)).  The scalar paths spell that with __fmul_rn / __fadd_rn.  The forward's packed hot
// loop cannot simply write mul.rn.f32x2 then add.rn.f32x2: ptxas 12.9 contracts that pair into one FFMA2 (the scalar .rn
// forms are never contracted, the f32x2 ones are -- a two-line kernel shows it), which is fma(du, du, fl(dv^2)), one
// rounding fewer; rounds 1 and 2 ran like that (round 2's "A/B of both forms", tools/q13_ab.py, timed the same SASS
// twice).  The loop now forms the sum as fma.rn.f32x2(fl(du^2), one, fl(dv^2)) with `one` a kernel parameter that ptxas can
// neither drop nor fuse: the reference's two roundings for one more FMUL2 per candidate (measured: 2.548 -> 2.574 ms; label
// mismatch 0 and scores within 3.6e-7 of the oracle either way, tools/seg_ab.py).  Vertices are split by weight once per sample:
//   light    w == 1        one per occupied z-buffer cell after compute_mask; min over SQUARED distances in the hot loop
//   heavy    w >= 256      d*w > 128 unless d < 0.5, and exp(-128) is exactly 0 in fp32, so a heavy vertex can only
//                           reach the one pixel it rounds to: chained per pixel, visited by that pixel alone
//   generic  anything else evaluated against every pixel (never produced by compute_mask; kept for drop-in inputs)
//
// Forward: a warp owns a 16 x 8 pixel tile (tiles are taken from a shared counter, the image's centre columns first), a
// lane a 2 x 2 block of it, so the per-axis squared offsets du^2 and dv^2 are shared by the block's pixels.  Once per
// tile, lane k prunes part k EXACTLY: with i0 the part's vertex nearest the tile centre c, f(g) = d_j^2(g) - d_i0^2(g) is linear in the pixel g, f(g) >= f(c) - 2(|du_j0| hw + |dv_j0| hh) on the
// tile, so a vertex whose bound clears a margin (far above the fp32 rounding of the squared distances) can never be the
// arg-min inside the tile and is dropped from the part's survivor words; the hot loop visits survivors only (about 3
// per (tile, part) instead of 12).  Eight channels of each pixel are staged per lane and leave as ONE 32-byte store
// (st.global.v8.f32 = a full DRAM sector; partial-sector stores cost a read-fill).  The score epilogue is
// d = sqrt.approx(d2), s = ex2.approx(-d*log2 e): <= 3e-7 absolute from exp(-sqrt(d2)), an order of magnitude inside
// the 1e-5 geometry tolerance that bounds the inputs.
// When a backward will follow, the forward also records per pixel the clip gate and the arg-min of every part as one
// byte (`saved`, 32 B per pixel, in the layout of the output itself: [n][wh-1-r][c][channel]), so the backward never
// searches.
//
// Backward: lane = channel, a warp walks output rows (taken from a shared counter) in groups of four pixels.  The
// pixel's upstream gradient row is one coalesced 128-byte load and its saved row one 32-byte load; s is recomputed at the recorded arg-min and the
// per-vertex sums accumulate in per-warp PRIVATE shared-memory slots indexed by the part's light slot: lane k is the
// only writer of part k's slots, so a plain load/add/store replaces atomics (shared fp32 atomicAdd is a CAS loop on
// sm_100) and no cross-lane reduction is needed.
// Gradient conventions (TF autodiff): the first arg-min takes the whole gradient on exact ties (TF splits
// evenly; measure zero); d == 0 yields 0 where TF yields NaN; the clip gate is inclusive (0 <= sum <= 1).
#include <math_constants.h>
#include "common.cuh"

namespace smplb200 {

namespace {

constexpr float kHeavyMin = 256.0f;
constexpr float kDropX = 110.0f;      // exp(-x) == 0 in fp32 (denormals included) for x > 103.98
constexpr unsigned kNone16 = 0xffffu;
// A row of survivor words (one tile of the forward): word 0 of part k at [k] -- its address does not depend on the part's
// descriptor, so the hot path loads both at once --, words w >= 1 at [kRow0 + woff[k] + w - 1].
constexpr int kRow0 = 32;
constexpr float kLog2e = 1.4426950408889634f;

struct SegSmem {
  float4* ent;     // [E]  light: {u, v, 1, vid}   heavy/generic: {u, v, w, entry | next << 16}
  unsigned short* head;   // [wh*wh] first heavy slot of each pixel, kNone16 none (16 bits: shared memory is what caps occupancy)
  int* lcount;     // [32] light entries per part (packed at the front of the part's CSR segment)
  int* lbase;      // [32] exclusive prefix sum of lcount
  int* pptr;       // [36] the part table's CSR pointers (P+1 used)
  int* woff;       // [36] exclusive prefix sum of max(ceil(lcount / 32) - 1, 0): offsets of the parts' survivor words 1.. (see kRow0)
  int4* pdesc;     // [32] per part {shared address of its first entry, survivor words, word offset, -}: one load in the hot path
  int* ghead;      // [1]  chain of generic slots, -1 none
  int* nheavy;     // [1]  number of chained heavy entries
  int* next_tile;  // [1]  next tile (in visiting order) to hand out
  unsigned char* rest;
};

// push `slot` on the 16-bit chain head of pixel `pix`; returns the previous head (kNone16 = none)
__device__ __forceinline__ unsigned head_push(unsigned short* head, int pix, unsigned slot) {
  unsigned* w = reinterpret_cast<unsigned*>(head) + (pix >> 1);
  const unsigned sh = (pix & 1) * 16;
  unsigned old = *w, assumed;
  do {
    assumed = old;
    old = atomicCAS(w, assumed, (assumed & ~(0xffffu << sh)) | (slot << sh));
  } while (old != assumed);
  return (old >> sh) & 0xffffu;
}

__device__ __forceinline__ float dist2(float u, float v, float gx, float gy) {
  const float du = __fsub_rn(u, gx), dv = __fsub_rn(v, gy);
  return __fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv));
}

// Layout: the fixed-size fields first, so that their shared addresses are the block's base plus a compile-time constant
// (the forward's per-part loop rebuilt `wh * wh * 2 + (E + 2) * 16 + ...` for every descriptor load when they followed
// the variable-size arrays: ~8 of its ~63 fixed instructions per (tile, part)).
constexpr size_t kSegFixed = 32 * 16 + 2 * 32 * 4 + 2 * 36 * 4 + 16;   // pdesc, lcount, lbase, pptr, woff, {ghead, nheavy, next_tile}
__device__ __forceinline__ SegSmem carve(unsigned char* raw, int E, int wh) {
  SegSmem sm;
  size_t off = 0;
  sm.pdesc = reinterpret_cast<int4*>(raw + off); off += 32 * 16;
  sm.lcount = reinterpret_cast<int*>(raw + off); off += 32 * 4;
  sm.lbase = reinterpret_cast<int*>(raw + off); off += 32 * 4;
  sm.pptr = reinterpret_cast<int*>(raw + off); off += 36 * 4;
  sm.woff = reinterpret_cast<int*>(raw + off); off += 36 * 4;
  sm.ghead = reinterpret_cast<int*>(raw + off);
  sm.nheavy = sm.ghead + 1;
  sm.next_tile = sm.ghead + 2; off += 16;
  sm.head = reinterpret_cast<unsigned short*>(raw + off); off += ((size_t)wh * wh * 2 + 15) & ~(size_t)15;
  sm.ent = reinterpret_cast<float4*>(raw + off) + 1; off += (size_t)(E + 2) * 16;   // ent[-1], ent[E]: readable dummies
  sm.rest = raw + off;
  return sm;
}
size_t seg_base_smem(int E, int wh) {
  return (size_t)(E + 2) * 16 + (((size_t)wh * wh * 2 + 15) & ~(size_t)15) + kSegFixed;
}

// Split the sample's part vertices into weight classes (one warp per part, ballot compaction).
__device__ void classify(const SegSmem& sm, const float* __restrict__ proj, const float* __restrict__ mask,
                         const int* __restrict__ ptr, const int* __restrict__ idx, int P, int wh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < (wh * wh + 1) / 2; i += blockDim.x) reinterpret_cast<unsigned*>(sm.head)[i] = 0xffffffffu;
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
              // the vertex wins iff xh < sqrt(min_i d2_i); one light vertex with d2 <= xh^2 already settles it (sqrtf is
              // correctly rounded, so sqrtf(fl(xh*xh)) == xh and the early exit cannot change the comparison)
              const float xh2 = __fmul_rn(xh, xh);
              float best = CUDART_INF_F;
              for (int i = 0; i < nl && best > xh2; ++i) {
                const float4 l = sm.ent[p0 + i];
                best = fminf(best, dist2(l.x, l.y, pu, pv));
              }
              if (xh < sqrtf(best)) {
                const unsigned next = head_push(sm.head, (int)pv * wh + (int)pu, (unsigned)slot);
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
  if (threadIdx.x < 32) {                               // exclusive scans of the light counts and of their word counts
    const int c = sm.lcount[threadIdx.x];
    const int cw = (c + 31) >> 5;
    const int cx = max(cw - 1, 0);                      // words beyond the part's first (they follow the kRow0 first words)
    int s = c, sw = cx;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, o), tw = __shfl_up_sync(0xffffffffu, sw, o);
      if ((int)threadIdx.x >= o) { s += t; sw += tw; }
    }
    sm.lbase[threadIdx.x] = s - c;
    sm.woff[threadIdx.x] = sw - cx;
    sm.pdesc[threadIdx.x] = make_int4((int)((uint32_t)__cvta_generic_to_shared(sm.ent) + (uint32_t)sm.pptr[min((int)threadIdx.x, 35)] * 16u),
                                      cw, kRow0 + sw - cx - 1, 0);   // .z + w = row index of word w >= 1
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

__device__ __forceinline__ void st_global_v8(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// Tile geometry shared by forward and launchers: a warp owns kTW x kTH pixels, lane (lx, ly) = (lane & 7, lane >> 3) a
// 2 x 2 block of it.
constexpr int kTW = 16, kTH = 8, kNB = 4;
constexpr float kTileHW = 0.5f * (kTW - 1), kTileHH = 0.5f * (kTH - 1);
struct SegGeom {
  int tiles_x, tiles_y, ntiles;
};
SegGeom seg_geom(int wh) {
  SegGeom g;
  g.tiles_x = (wh + kTW - 1) / kTW; g.tiles_y = (wh + kTH - 1) / kTH; g.ntiles = g.tiles_x * g.tiles_y;
  return g;
}
// row length: kRow0 + sum_k max(ceil(lcount_k / 32) - 1, 0) <= kRow0 + E / 32
__host__ __device__ __forceinline__ int seg_keep_words(int E) { return kRow0 + E / 32 + 1; }

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
constexpr float kBigD2 = 1e30f;      // "no vertex yet": sqrt/ex2 map it to a score of exactly 0 without a branch
constexpr float kPruneMargin = 0.01f;   // px^2; plus a relative term, see prune_tile
constexpr float kPruneRel = 4e-6f;

__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ unsigned bfind_u32(unsigned x) {           // index of the highest set bit (x != 0)
  unsigned r;
  asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}

// explicit shared-window accesses (a generic pointer into dynamic smem makes the compiler rebuild the window base)
__device__ __forceinline__ float2 lds_f2_nv(uint32_t a) {            // read-only data: free to be scheduled early
  float2 r;
  asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(a));
  return r;
}
__device__ __forceinline__ unsigned long long lds_b64(uint32_t a) {
  unsigned long long r;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ unsigned long long lds_b64_nv(uint32_t a) {
  unsigned long long r;
  asm("ld.shared.b64 %0, [%1];" : "=l"(r) : "r"(a));
  return r;
}
// (volatile: the survivor words are rewritten per tile in the per-warp-row mode and the descriptors by classify; the
// statement must stay behind the __syncwarp / __syncthreads that publishes them and must not be merged across tiles)
__device__ __forceinline__ unsigned lds_u32(uint32_t a) {
  unsigned r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ int4 lds_v4(uint32_t a) {
  int4 r;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
  return r;
}
// the forward's per-lane staging slots: written and read by the same lane, ordered by the volatile qualifier
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float r;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(a) : "memory");
  return r;
}
__device__ __forceinline__ void sts_b64(uint32_t a, unsigned long long v) {
  asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}

// Exact pruning of one tile, lane k working on part k: writes the part's survivor words to kw[woff[k] ..].
// i0 = the part's vertex nearest the tile centre (tcx, tcy).  f(g) = d_j^2(g) - d_i0^2(g) is linear in the pixel g:
// f(g) = f(c) + 2 (p_i0 - p_j).(g - c) >= f(c) - 2 (|u_j - u_i0| hw + |v_j - v_i0| hh) on the tile.  If that bound is at
// least margin = 0.01 + 4e-6 d_j^2(c), vertex j can never be the fp32 arg-min inside the tile (the rounding of the
// squared distances and of this test is below 1e-6 relative) and is dropped.  i0 itself always survives (bound = 0).
__device__ __forceinline__ void prune_tile(const SegSmem& sm, unsigned* kw, int P, int lane, float tcx, float tcy) {
  if (lane < P) {
    const int n0 = sm.lcount[lane];
    if (n0 > 0) {
      const float4* ek = sm.ent + sm.pptr[lane];
      unsigned* kwx = kw + kRow0 + sm.woff[lane] - 1;              // word w >= 1 at kwx[w]
      float bd = CUDART_INF_F, u0 = 0.f, v0 = 0.f;
      for (int v = 0; v < n0; ++v) {
        const float2 e = *reinterpret_cast<const float2*>(&ek[v]);
        const float du = e.x - tcx, dv = e.y - tcy;
        const float d = du * du + dv * dv;
        if (d < bd) { bd = d; u0 = e.x; v0 = e.y; }
      }
      unsigned word = 0u;
      for (int v = 0; v < n0; ++v) {
        const float2 e = *reinterpret_cast<const float2*>(&ek[v]);
        const float du = e.x - tcx, dv = e.y - tcy;
        const float d = du * du + dv * dv;
        const float slack = 2.0f * (fabsf(e.x - u0) * kTileHW + fabsf(e.y - v0) * kTileHH);
        const bool keep = (d - bd) - slack < kPruneMargin + kPruneRel * d;
        word |= (keep ? 1u : 0u) << (v & 31);
        if ((v & 31) == 31) { *((v >> 5) ? kwx + (v >> 5) : kw + lane) = word; word = 0u; }
      }
      if (n0 & 31) *((n0 >> 5) ? kwx + (n0 >> 5) : kw + lane) = word;
    }
  }
  __syncwarp();
}

// The same pruning for ALL tiles of the block at once, lane = tile, warp = part: the part's vertices are walked with a
// warp-uniform trip count and broadcast loads, where prune_tile's lane = part loop runs to the LARGEST part's count with
// most lanes idle (ncu: 11 - 16 % of the forward).  Same arithmetic, so the survivor words are identical.  kwt =
// [tiles of this block][KW] words; tile ti (visiting order, see the kernel) owns row ti - t0.
__device__ void prune_all_tiles(const SegSmem& sm, unsigned* kwt, int KW, int P, int t0, int t1, int tiles_x, int tiles_y) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int tb = t0; tb < t1; tb += 32) {
    const bool act = tb + lane < t1;
    const int ti = act ? tb + lane : t0;
    const int ci = ti / tiles_y, ty = ti - ci * tiles_y;
    const int tx = (tiles_x >> 1) + ((ci & 1) ? -((ci + 1) >> 1) : (ci >> 1));
    const float tcx = (float)(tx * kTW) + kTileHW, tcy = (float)(ty * kTH) + kTileHH;
    unsigned* kwl = kwt + (size_t)(ti - t0) * KW;
    for (int k = warp; k < P; k += nwarps) {
      const int n0 = sm.lcount[k];                                   // same address on every lane: broadcast
      if (n0 == 0) continue;
      const float4* ek = sm.ent + sm.pptr[k];
      unsigned* kwx = kwl + kRow0 + sm.woff[k] - 1;                 // word w >= 1 at kwx[w]
      float bd = CUDART_INF_F, u0 = 0.f, v0 = 0.f;
      for (int v = 0; v < n0; ++v) {
        const float2 e = *reinterpret_cast<const float2*>(&ek[v]);
        const float du = e.x - tcx, dv = e.y - tcy;
        const float d = du * du + dv * dv;
        if (d < bd) { bd = d; u0 = e.x; v0 = e.y; }
      }
      unsigned word = 0u;
      for (int v = 0; v < n0; ++v) {
        const float2 e = *reinterpret_cast<const float2*>(&ek[v]);
        const float du = e.x - tcx, dv = e.y - tcy;
        const float d = du * du + dv * dv;
        const float slack = 2.0f * (fabsf(e.x - u0) * kTileHW + fabsf(e.y - v0) * kTileHH);
        const bool keep = (d - bd) - slack < kPruneMargin + kPruneRel * d;
        word |= (keep ? 1u : 0u) << (v & 31);
        if ((v & 31) == 31) { if (act) *((v >> 5) ? kwx + (v >> 5) : kwl + k) = word; word = 0u; }
      }
      if ((n0 & 31) && act) *((n0 >> 5) ? kwx + (n0 >> 5) : kwl + k) = word;
    }
  }
}

// Packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2: two fp32 lanes per instruction, each rounded to nearest like the
// scalar op).  A pair lives in an aligned 64-bit register; ptxas reads a scalar operand as a broadcast (R.F32).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// Squared distances of one vertex to a lane's 2 x 2 pixel block, GX = (gx0, gx1), GY = (gy0, gy1):
// d2[q] = fl(fl(du^2) + fl(dv^2)): two FADD2, two FMUL2, two FFMA2 for the four pixels.
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ void block_d2(float2 e, f32x2 GX, f32x2 GY, float (&d2)[kNB], f32x2 ONE2) {
  const f32x2 dx = sub2(pk2(e.x, e.x), GX), dy = sub2(pk2(e.y, e.y), GY);
  float vy0, vy1;
  upk2(mul2(dy, dy), vy0, vy1);
  const f32x2 dx2 = mul2(dx, dx);                                  // fl(du^2), kept as a product of its own (see the header)
  // fl(du^2) * 1 + fl(dv^2), ONE2 = (1, 1) from a kernel parameter
  upk2(fma2(dx2, ONE2, pk2(vy0, vy0)), d2[0], d2[1]);
  upk2(fma2(dx2, ONE2, pk2(vy1, vy1)), d2[2], d2[3]);
}

// One survivor against a lane's 2 x 2 pixel block: best[q] = min squared distance, barg[q] = code of its arg-min,
// (index + 1) << sh.  CLAMP: indices >= 254 do not fit a byte and are recorded as 255 (re-queried by the backward);
// only words 7 and up can hold them.  FIRST: the block's first candidate of this part: no comparison needed.
template <bool TRACK, bool CLAMP, bool FIRST>
__device__ __forceinline__ void compare_vertex(float2 e, unsigned b, int w, unsigned sh, unsigned shmul, unsigned wcode,
                                               f32x2 GX, f32x2 GY, float (&best)[kNB], unsigned (&barg)[kNB], f32x2 ONE2) {
  float d2[kNB];
  block_d2(e, GX, GY, d2, ONE2);
  if (FIRST) {
    const unsigned vcode = CLAMP ? (unsigned)min(w * 32 + (int)b + 1, 255) << sh : b * shmul + wcode;
#pragma unroll
    for (int q = 0; q < kNB; ++q) { best[q] = d2[q]; if (TRACK) barg[q] = vcode; }
  } else if (TRACK) {
    const unsigned vcode = CLAMP ? (unsigned)min(w * 32 + (int)b + 1, 255) << sh : b * shmul + wcode;
#pragma unroll
    for (int q = 0; q < kNB; ++q) {
      const bool le = d2[q] <= best[q];
      barg[q] = le ? vcode : barg[q];
      best[q] = le ? d2[q] : best[q];
    }
  } else {
#pragma unroll
    for (int q = 0; q < kNB; ++q) best[q] = fminf(best[q], d2[q]);
  }
}

// FRESH: best / barg hold nothing yet: the word's first survivor initialises them (an empty word leaves kBigD2 / 0).
template <bool TRACK, bool CLAMP, bool FRESH>
__device__ __forceinline__ void scan_word(unsigned m, uint32_t eb, int w, unsigned sh, f32x2 GX, f32x2 GY,
                                          float (&best)[kNB], unsigned (&barg)[kNB], f32x2 ONE2) {
  const unsigned shmul = 1u << sh, wcode = (unsigned)(w * 32 + 1) << sh;
  if (FRESH) {
    if (m == 0u) {
#pragma unroll
      for (int q = 0; q < kNB; ++q) { best[q] = kBigD2; barg[q] = 0u; }
      return;
    }
    const unsigned b0 = bfind_u32(m);
    m ^= 1u << b0;
    compare_vertex<TRACK, CLAMP, true>(lds_f2_nv(eb + b0 * 16u), b0, w, sh, shmul, wcode, GX, GY, best, barg, ONE2);
  }
  // survivors two at a time: both coordinate loads are in flight before either vertex is compared
  while (m) {
    const unsigned b0 = bfind_u32(m);
    m ^= 1u << b0;
    const float2 e0 = lds_f2_nv(eb + b0 * 16u);
    if (m) {
      const unsigned b1 = bfind_u32(m);
      m ^= 1u << b1;
      const float2 e1 = lds_f2_nv(eb + b1 * 16u);
      compare_vertex<TRACK, CLAMP, false>(e0, b0, w, sh, shmul, wcode, GX, GY, best, barg, ONE2);
      compare_vertex<TRACK, CLAMP, false>(e1, b1, w, sh, shmul, wcode, GX, GY, best, barg, ONE2);
    } else {
      compare_vertex<TRACK, CLAMP, false>(e0, b0, w, sh, shmul, wcode, GX, GY, best, barg, ONE2);
    }
  }
}

// Fused Reshape -> softmax -> categorical focal loss (model.py:119-120, focal_loss.py:10-48) with integer labels: what the
// LOSS variants of the two kernels exchange per OUTPUT pixel (16 bytes instead of the 128-byte upstream gradient row).
//   x = al * q_l      al = w_l [gamma (1-p_l)^(gamma-1) log p_l - (1-p_l)^gamma / p_l] where the clip passes, else 0
//   y = x / Z         Z = sum_c exp(s_c): the gradient w.r.t. score c is  g_loss * (x [c == l] - y exp(s_c))
//   z = gate * (x [l == 0] - y exp(s_0))   the background channel's share, routed to every part through 1 - clip(sum)
//   w = the label, as a float
struct SegLossArgs {
  const unsigned char* labels;   // [N][wh][wh] class ids in OUTPUT pixel order (rows flipped, like y_true)
  const float* class_w;          // [32] or null (ones)
  float gamma;
  float* loss;                   // [N][wh*wh]
  float4* aux;                   // [N][wh*wh]
};
constexpr float kKerasEps = 1e-7f;          // K.epsilon(), focal_loss.py:15
__device__ __forceinline__ float pow_gamma(float x, float gamma) { return gamma == 2.0f ? x * x : powf(x, gamma); }

// One pixel's focal loss and backward record from its softmax bookkeeping (Zp = sum of exp(s_k) over the parts, elp =
// exp(score) of the label's channel if it is a part, bg = the background score).  Out of line: once per pixel, and the
// rasteriser's instruction footprint is what its issue rate hangs on (inlined four times per lane with logf / powf it
// grew the kernel from 44 to 72 KB).
__device__ __noinline__ float4 loss_pixel(float Zp, float elp, float bg, int lab, int C, bool gate,
                                          const float* __restrict__ class_w, float gamma, float* __restrict__ loss_out) {
  const float e0 = ex2_approx(bg * kLog2e);
  const float Z = Zp + e0;
  const float elq = (lab == 0) ? e0 : elp;
  const float invZ = 1.0f / Z;
  const float ql = elq * invZ;                                       // softmax of the label's channel (model.py:120)
  const float pl = fminf(fmaxf(ql, kKerasEps), 1.0f - kKerasEps);    // focal_loss.py:16
  const bool inlab = lab < C;
  const float wl = (class_w && inlab) ? class_w[lab] : 1.0f;
  const float om = 1.0f - pl, lg = logf(pl);
  float al = 0.f;
  if (inlab && ql >= kKerasEps && ql <= 1.0f - kKerasEps) {          // TF: the clip's gradient passes on the closed interval
    const float dpow = gamma == 2.0f ? 2.0f * om : gamma * powf(om, gamma - 1.0f);
    al = wl * (dpow * lg - pow_gamma(om, gamma) / pl);
  }
  *loss_out = inlab ? pow_gamma(om, gamma) * ((-lg) * wl) : 0.f;     // focal_loss.py:17, 40, 44-45
  const float x = al * ql, y = x * invZ;
  const float z = gate ? x * ((lab == 0) ? 1.0f : 0.f) - y * e0 : 0.f;
  return make_float4(x, y, z, (float)lab);
}

// saved layout: 32 bytes per OUTPUT pixel, [n][wh-1-r][c][32].  byte 0: bit 0 = clip gate.  byte 1+k (part k): 0 none,
// 1..254 light index + 1, 255 re-query (heavy / generic winner or light index >= 254).
// (The LOSS variant carries twelve more live values through the part loop -- softmax denominators, the label's
// numerator, the labels -- and spilled them at the 80 registers of four blocks per SM: it is compiled without that cap
// and runs three 6-warp blocks per SM.)
// WH: img_wh as a compile-time constant (0 = the runtime argument); C32: 31 parts + background = one 128-byte line per pixel.
// The training resolution (48, a whole number of 16 x 8 tiles) compiles without bounds checks and with constant tile counts.
template <bool TRACK, bool LOSS, int WH, bool C32>
__global__ void __launch_bounds__(LOSS ? 288 : 256, LOSS ? 2 : 3)   // LOSS: 112 registers (three 6-warp blocks per SM), 8-warp launches allowed
seg_fwd_kernel(const float* __restrict__ projects, const float* __restrict__ mask, int N, int Vs,
               const int* __restrict__ ptr, const int* __restrict__ idx, int P, int E, int wh_arg,
               float* __restrict__ seg, unsigned char* __restrict__ saved, const SegLossArgs la, int KW, int kw_rows,
               float one) {
  extern __shared__ __align__(16) unsigned char raw[];
  const f32x2 ONE2 = pk2(one, one);
  const int wh = WH ? WH : wh_arg;
  constexpr bool kWhole = WH != 0 && WH % kTW == 0 && WH % kTH == 0;   // no partial tiles: every pixel of a tile is in the image
  const SegSmem sm = carve(raw, E, wh);
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int C = C32 ? 32 : P + 1;
  const int tiles_x = (wh + kTW - 1) / kTW, tiles_y = (wh + kTH - 1) / kTH, ntiles = tiles_x * tiles_y;
  const int t0 = (int)(((long long)ntiles * blockIdx.y) / gridDim.y);
  const int t1 = (int)(((long long)ntiles * (blockIdx.y + 1)) / gridDim.y);
  const int lx = lane & 7, ly = lane >> 3;
  unsigned char* sv = TRACK ? saved + (size_t)n * wh * wh * 32 : nullptr;
  float* seg_n = seg ? seg + (size_t)n * wh * wh * C : nullptr;
  // survivor words: kw_rows == nwarps: one row of KW words per warp, filled per tile by prune_tile; otherwise one row
  // per tile of this block (kw_rows >= t1 - t0), filled once by prune_all_tiles
  const bool kw_table = kw_rows != nwarps;
  unsigned* const kw_base = reinterpret_cast<unsigned*>(sm.rest);
  // per-lane staging of one chunk: stage[(sub*4 + q)*32 + lane]  (lane-contiguous: conflict-free, no sync needed)
  float* stage = reinterpret_cast<float*>(sm.rest + (((size_t)kw_rows * KW * 4 + 15) & ~(size_t)15)) +
                 (size_t)warp * (8 * kNB * 32) + lane;
  // shared-window addresses of what the per-part loop touches (sm.pdesc sits at the block's base, see carve)
  const uint32_t pdesc_sa = (uint32_t)__cvta_generic_to_shared(raw);
  const uint32_t stage_sa = (uint32_t)__cvta_generic_to_shared(stage);

  // Tiles are handed to the warps on demand (the first nwarps statically), the image's centre columns first: they hold
  // the body and cost the most, and a fixed tile -> warp map gave all of them to the same two warps.
  if (threadIdx.x == 0) *sm.next_tile = t0 + nwarps;                 // ordered before its first use by classify's barriers
  classify(sm, projects + (size_t)n * Vs * 3, mask + (size_t)n * Vs, ptr, idx, P, wh);
  const int ghead = *sm.ghead;
  const bool any_heavy = *sm.nheavy > 0;
  if (kw_table) {
    prune_all_tiles(sm, kw_base, KW, P, t0, t1, tiles_x, tiles_y);
    __syncthreads();
  }
  for (int ti = t0 + warp; ti < t1;) {
    const int ci = ti / tiles_y, ty = ti - ci * tiles_y;
    const int tx = (tiles_x >> 1) + ((ci & 1) ? -((ci + 1) >> 1) : (ci >> 1));
    const int c0 = tx * kTW + lx * 2, r0 = ty * kTH + ly * 2;        // this lane's block origin (grid = (column,row), :26-31)
    const float gx0 = (float)c0, gx1 = (float)(c0 + 1), gy0 = (float)r0, gy1 = (float)(r0 + 1);
    const f32x2 GX = pk2(gx0, gx1), GY = pk2(gy0, gy1);
    const unsigned* kw = kw_base + (size_t)(kw_table ? ti - t0 : warp) * KW;
    const uint32_t kw_sa = (uint32_t)__cvta_generic_to_shared(kw);
    if (!kw_table) prune_tile(sm, kw_base + (size_t)warp * KW, P, lane, (float)(tx * kTW) + kTileHW, (float)(ty * kTH) + kTileHH);

    bool blk_slow = ghead >= 0;                                      // any heavy vertex chained to one of my pixels?
    if (any_heavy) {
#pragma unroll
      for (int q = 0; q < kNB; ++q) {
        const int r = r0 + (q >> 1), c = c0 + (q & 1);
        if (kWhole || (c < wh && r < wh)) blk_slow |= sm.head[r * wh + c] != kNone16;
      }
    }
    float S[kNB];
#pragma unroll
    for (int q = 0; q < kNB; ++q) S[q] = 0.f;
    // LOSS: softmax denominator, exp(score) of the label's channel and the label of each of the lane's pixels
    float Zs[kNB], el[kNB];
    int lab[kNB];
    if (LOSS) {
#pragma unroll
      for (int q = 0; q < kNB; ++q) {
        const int r = r0 + (q >> 1), c = c0 + (q & 1);
        Zs[q] = 0.f; el[q] = 0.f;
        lab[q] = (kWhole || (r < wh && c < wh)) ? (int)la.labels[(size_t)n * wh * wh + (size_t)(wh - 1 - r) * wh + c] : 0;
      }
    }

    // channel chunks 1, 2, 3, then chunk 0 last: its channel 0 (background) needs the sum over all parts
    for (int cc = 1; cc <= 4; ++cc) {
      const int chunk = cc & 3;
      if (!C32 && chunk * 8 >= C) continue;
      unsigned cw[2][kNB];                                           // packed saved bytes, channels 0-3 / 4-7 of the chunk
#pragma unroll
      for (int q = 0; q < kNB; ++q) { cw[0][q] = 0u; cw[1][q] = 0u; }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll 1   // (unrolling by two: 2.67 -> 3.92 ms; a single copy of the body for both halves: 2.675 -> 2.711 ms)
        for (int s4 = 0; s4 < 4; ++s4) {
          const int sub = half * 4 + s4;
          const int ch = chunk * 8 + sub;
          const uint32_t st_sa = stage_sa + (uint32_t)sub * (kNB * 32 * 4);
          if ((half == 0 && ch == 0) || (!C32 && ch >= C)) {         // background is filled in after the loop
#pragma unroll
            for (int q = 0; q < kNB; ++q) sts_f32(st_sa + q * 128, 0.f);
            continue;
          }
          const int4 pd = lds_v4(pdesc_sa + (uint32_t)(ch - 1) * 16u);   // {entry address, survivor words, row index of word 1 - 1}
          const unsigned m0 = lds_u32(kw_sa + (uint32_t)(ch - 1) * 4u);  // the part's first survivor word: independent of pd
          const unsigned sh = 8u * (unsigned)s4;
          float best[kNB];
          unsigned barg[kNB];                                        // arg-min code, already shifted to its byte
          // survivors are visited from the highest index down and replace on <=, so the LOWEST index wins exact ties
          if (pd.y == 1) {                                           // at most 32 visible vertices: the common case
            scan_word<TRACK, false, true>(m0, (uint32_t)pd.x, 0, sh, GX, GY, best, barg, ONE2);
          } else {
#pragma unroll
            for (int q = 0; q < kNB; ++q) { best[q] = kBigD2; barg[q] = 0u; }
            for (int w = pd.y; w-- > 0;) {
              const unsigned m = w ? kw[pd.z + w] : m0;              // same address on every lane: broadcast
              const uint32_t eb = (uint32_t)pd.x + (uint32_t)(w * 32) * 16u;
              if (TRACK && w >= 7) scan_word<TRACK, true, false>(m, eb, w, sh, GX, GY, best, barg, ONE2);
              else scan_word<TRACK, false, false>(m, eb, w, sh, GX, GY, best, barg, ONE2);
            }
          }
          float sc[kNB];
#pragma unroll
          for (int q = 0; q < kNB; ++q)
            sc[q] = ex2_approx(sqrt_approx(best[q]) * (-kLog2e));    // exp(-sqrt(d2)); kBigD2 -> 0, d2 == 0 -> 1
          if (blk_slow) {                                            // rare: heavy / generic vertices
#pragma unroll
            for (int q = 0; q < kNB; ++q) {
              const int r = r0 + (q >> 1), c = c0 + (q & 1);
              if (kWhole || (c < wh && r < wh)) {
                const int hd = (any_heavy && sm.head[r * wh + c] != kNone16) ? (int)sm.head[r * wh + c] : -1;
                if (hd >= 0 || ghead >= 0) {
                  const float ss = slow_pixel_score(sm, sm.pptr[ch - 1], sm.pptr[ch], (q & 1) ? gx1 : gx0, (q >> 1) ? gy1 : gy0, hd, ghead, best[q]);
                  if (ss >= 0.f) { sc[q] = ss; barg[q] = 255u << sh; }
                }
              }
            }
          }
#pragma unroll
          for (int q = 0; q < kNB; ++q) {
            sts_f32(st_sa + q * 128, sc[q]);
            S[q] += sc[q];
            if (TRACK) cw[half][q] |= barg[q];
            if (LOSS) {                                              // softmax numerator of this channel (scores are in [0, 1]: no max shift)
              const float ex = ex2_approx(sc[q] * kLog2e);
              Zs[q] += ex;
              el[q] = (lab[q] == ch) ? ex : el[q];
            }
          }
        }
      }
      if (chunk == 0) {
#pragma unroll
        for (int q = 0; q < kNB; ++q) {
          const float bg = 1.0f - fminf(fmaxf(S[q], 0.f), 1.f);      // :61-64
          sts_f32(stage_sa + q * 128, bg);
          const bool gate = S[q] >= 0.f && S[q] <= 1.f;              // clip gate (inclusive), for the backward
          if (TRACK) cw[0][q] |= gate ? 1u : 0u;
          if (LOSS) {
            const int r = r0 + (q >> 1), c = c0 + (q & 1);
            if (kWhole || (r < wh && c < wh)) {
              const size_t opx = (size_t)n * wh * wh + (size_t)(wh - 1 - r) * wh + c;
              la.aux[opx] = loss_pixel(Zs[q], el[q], bg, lab[q], C, gate, la.class_w, la.gamma, la.loss + opx);
            }
          }
        }
      }
      // one 32-byte sector per pixel; rows flipped (:68)
#pragma unroll
      for (int q = 0; q < kNB; ++q) {
        const int r = r0 + (q >> 1), c = c0 + (q & 1);
        if (kWhole || (r < wh && c < wh)) {
          const size_t opx = (size_t)(wh - 1 - r) * wh + c;
          if (seg_n) {
            float* o = seg_n + opx * C + chunk * 8;
            float v8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v8[e] = lds_f32(stage_sa + (e * kNB + q) * 128);
            if (C32) {
              st_global_v8(o, v8);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (chunk * 8 + e < C) o[e] = v8[e];
            }
          }
          if (TRACK) *reinterpret_cast<uint2*>(sv + opx * 32 + chunk * 8) = make_uint2(cw[0][q], cw[1][q]);
        }
      }
    }
    int nt = 0;
    if (lane == 0) nt = atomicAdd(sm.next_tile, 1);
    ti = __shfl_sync(0xffffffffu, nt, 0);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------
// Shared memory of the backward: only the LIGHT vertices (w == 1) are kept, in a lane-interleaved layout: slot i of part
// k sits at [i][k], so when lane = part every lane's 8-byte access falls in its own bank pair whatever i is (a warp-wide
// 64-bit access is exactly two conflict-free wavefronts).  The private accumulators of each warp use the same layout.
// Parts with more than kIL visible vertices spill to a compact overflow list (shared accumulators, atomics; 0.3 % of
// the entries at vertex_sampling=5).  Heavy / generic winners (code 255) are re-queried from global memory.
constexpr int kIL = 32;            // interleaved light slots per part
constexpr int kAccRows = kIL + 1;  // accumulator rows per warp: row kIL takes the (discarded) sums of "none" and rare codes
constexpr unsigned kNoVid = 0xffffu;
constexpr int kOvPriv = 64;        // private overflow slots per warp; a sample that needs more spills to shared atomics
constexpr size_t kBwdHeader = 512; // head of the block's shared memory: per-warp scratch of the LOSS variant (8 warps x 64 B)

struct BwdSmem {
  float2* lpos;          // [kAccRows][32]   row kIL = zeros: what "none" and the rare codes read
  unsigned short* lvid;  // [kIL][32]
  int* lcount;           // [32]
  int* obase;            // [36]  overflow offsets of the parts (prefix of max(size_k - kIL, 0))
  float2* opos;          // [OV]
  float2* oacc;          // [OV]
  unsigned short* ovid;  // [OV]  kNoVid = unused
  float2* wacc;          // [nwarps][kAccRows][32]
  float2* wov;           // [nwarps][kOvPriv]  private overflow accumulators, slots handed out per sample (odyn)
  int* odyn;             // [32]  exclusive prefix of max(lcount_k - kIL, 0) for THIS sample
  int* next_row;         // [1]   next output row to hand out (ALIGNED schedule)
};
__host__ __device__ __forceinline__ size_t bwd_smem_bytes(int OV, int nwarps) {
  const size_t ov = ((size_t)OV + 7) & ~(size_t)7;
  const size_t need = kBwdHeader + (size_t)kAccRows * 32 * 8 + (size_t)kIL * 32 * 2 + 2 * 32 * 4 + 36 * 4 + 16 + ov * 8 * 2 + ov * 2 +
                      (size_t)nwarps * (kAccRows * 32 + kOvPriv) * 8;
  return need;
}
__device__ __forceinline__ BwdSmem carve_bwd(unsigned char* raw, int OV, int nwarps) {
  BwdSmem b;
  const size_t ov = ((size_t)OV + 7) & ~(size_t)7;
  size_t off = kBwdHeader;
  b.lpos = reinterpret_cast<float2*>(raw + off); off += (size_t)kAccRows * 32 * 8;
  b.wacc = reinterpret_cast<float2*>(raw + off); off += (size_t)nwarps * kAccRows * 32 * 8;
  b.wov = reinterpret_cast<float2*>(raw + off); off += (size_t)nwarps * kOvPriv * 8;
  b.opos = reinterpret_cast<float2*>(raw + off); off += ov * 8;
  b.oacc = reinterpret_cast<float2*>(raw + off); off += ov * 8;
  b.lcount = reinterpret_cast<int*>(raw + off); off += 32 * 4;
  b.odyn = reinterpret_cast<int*>(raw + off); off += 32 * 4;
  b.obase = reinterpret_cast<int*>(raw + off); off += 36 * 4;
  b.next_row = reinterpret_cast<int*>(raw + off); off += 16;
  b.lvid = reinterpret_cast<unsigned short*>(raw + off); off += (size_t)kIL * 32 * 2;
  b.ovid = reinterpret_cast<unsigned short*>(raw + off);
  return b;
}

// The light list of every part, in the forward's order (entries of the part's CSR segment with w == 1, compacted).
__device__ void classify_light(const BwdSmem& b, const float* __restrict__ proj, const float* __restrict__ mask,
                               const int* __restrict__ ptr, const int* __restrict__ idx, const int* __restrict__ obase,
                               int P, int OV) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (threadIdx.x < 32) b.lcount[threadIdx.x] = 0;
  if (threadIdx.x < 36) b.obase[threadIdx.x] = obase[min((int)threadIdx.x, P)];
  for (int i = threadIdx.x; i < OV; i += blockDim.x) { b.ovid[i] = (unsigned short)kNoVid; b.oacc[i] = make_float2(0.f, 0.f); }
  // The initialisation above must be complete before any warp publishes a part's count or overflow entries below.  (This
  // barrier was missing: a warp held up for a few microseconds -- it took the TMA queue behind the L2 bulk prefetch to do
  // it -- zeroed a count another warp had already written, and that part's gradient was dropped for the sample.)
  __syncthreads();
  for (int k = warp; k < P; k += nwarps) {
    const int p0 = ptr[k], p1 = ptr[k + 1], ob = obase[k];
    int nl = 0;
    constexpr int kU = 4;                                  // chunks of 32 entries whose gather chains overlap
    for (int base0 = p0; base0 < p1; base0 += 32 * kU) {
      int vids[kU];
      float us[kU], vv[kU], ws[kU];
#pragma unroll
      for (int c = 0; c < kU; ++c) {
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
        if (base0 + c * 32 >= p1) break;
        const bool light = (base0 + c * 32 + lane < p1) && (ws[c] == 1.0f);
        const unsigned bl = __ballot_sync(0xffffffffu, light);
        if (light) {
          const int pos = nl + __popc(bl & ((1u << lane) - 1u));
          if (pos < kIL) {
            b.lpos[pos * 32 + k + 1] = make_float2(us[c], vv[c]);        // column = channel = part + 1
            b.lvid[pos * 32 + k + 1] = (unsigned short)vids[c];
          } else {
            b.opos[ob + pos - kIL] = make_float2(us[c], vv[c]);
            b.ovid[ob + pos - kIL] = (unsigned short)vids[c];
          }
        }
        nl += __popc(bl);
      }
    }
    if (lane == 0) b.lcount[k] = nl;
  }
}

// rare: the forward's winner at this pixel was a heavy / generic vertex (or a light index that did not fit a byte):
// exact weighted nearest-vertex re-query of part [p0,p1) straight from global memory, with the forward's rule -- the
// light minimum over squared distances (lowest index on ties), beaten only by a strictly smaller d*w of another vertex.
// The upstream gradient of the score is G(s) = ga - gb * exp(s): gb = 0 for a plain g_seg, the fused loss's form otherwise.
__device__ __noinline__ void slow_pixel_grad_global(const float* __restrict__ proj, const float* __restrict__ mask,
                                                    const int* __restrict__ idx, int p0, int p1, float gx, float gy,
                                                    float ga, float gb, float* __restrict__ out) {
  float best = CUDART_INF_F, xo = CUDART_INF_F;
  int lv = -1, ov = -1;
  for (int e = p0; e < p1; ++e) {
    const int vid = idx[e];
    const float w = mask[vid];
    const float d2 = dist2(proj[vid * 3], proj[vid * 3 + 1], gx, gy);
    if (w == 1.0f) {
      if (d2 < best) { best = d2; lv = vid; }
    } else {
      const float xe = __fmul_rn(sqrtf(d2), w);
      if (xe < xo) { xo = xe; ov = vid; }
    }
  }
  const int vid = (ov >= 0 && xo < sqrtf(best)) ? ov : lv;
  if (vid < 0) return;
  const float w = mask[vid];
  const float du = __fsub_rn(proj[vid * 3], gx), dv = __fsub_rn(proj[vid * 3 + 1], gy);
  const float d = sqrtf(__fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv)));
  const float s = expf(-__fmul_rn(d, w));
  const float G = ga - gb * expf(s);
  const float coef = (d > 0.f) ? (-w * s * G) / d : 0.f;           // d(exp(-d w))/dp = -w s (p - g)/d
  atomicAdd(&out[vid * 3], coef * du);
  atomicAdd(&out[vid * 3 + 1], coef * dv);
}

// rare: light winner beyond the interleaved slots (a part with more than kIL visible vertices).  j = li - kIL.  The
// first kOvPriv overflow slots of a sample are private to (warp, lane = part): plain read-modify-write; beyond that,
// shared atomics.
template <bool LOSS>
__device__ __forceinline__ void overflow_pixel_grad(const BwdSmem& b, float2* wov_w, int ob, int od, int j, float gx,
                                                    float gy, float ga, float gb) {
  const float2 e = b.opos[ob + j];
  const float du = __fsub_rn(e.x, gx), dv = __fsub_rn(e.y, gy);
  const float d2 = __fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv));
  const float rs = rsqrt_approx(fmaxf(d2, 1e-30f));
  const float s = ex2_approx((d2 * rs) * (-kLog2e));
  const float G = LOSS ? ga - gb * ex2_approx(s * kLog2e) : ga;
  const float coef = (s * G) * (-rs);
  if (od + j < kOvPriv) {
    float2 a = wov_w[od + j];
    a.x += coef * du; a.y += coef * dv;
    wov_w[od + j] = a;
  } else {
    atomicAdd(&b.oacc[ob + j].x, coef * du);
    atomicAdd(&b.oacc[ob + j].y, coef * dv);
  }
}

// LOSS: instead of an upstream gradient row per pixel the kernel reads the forward's 16-byte `aux` record and the
// upstream gradient of the per-pixel loss (see SegLossArgs): lane (j & 3) of a group loads pixel j's record, the group's
// values cross lanes by shuffle, and the score's own gradient  g (x [c == l] - y exp(s_c) - z)  is formed from the
// recomputed score.
// WH: img_wh as a compile-time constant (0 = the runtime argument).  ncu's source view of the runtime-wh kernel showed
// the per-group bookkeeping re-loading wh from the constant bank (no register to keep it at the 96-register cap) and
// stalling on it: ~8 % of the kernel's samples.
template <bool C32, bool ALIGNED, bool LOSS, int WH>
__global__ void __launch_bounds__(192, LOSS ? 2 : 3)
seg_bwd_kernel(const float* __restrict__ projects, const float* __restrict__ mask, const float* __restrict__ g_seg,
               const unsigned char* __restrict__ saved, int N, int Vs, const int* __restrict__ ptr,
               const int* __restrict__ idx, const int* __restrict__ obase, int P, int OV, int wh_arg,
               float* __restrict__ g_projects, int pf_ok, const float4* __restrict__ aux,
               const float* __restrict__ g_loss) {
  extern __shared__ __align__(16) unsigned char raw[];
  const int wh = WH ? WH : wh_arg;
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const BwdSmem b = carve_bwd(raw, OV, nwarps);
  // ALIGNED (wh % 4 == 0, wh >= 8): output rows (already flipped: grid row = wh-1-row) are handed to the warps on demand
  // (a shared counter): the scheduler does not run the warps of a block evenly, and with a fixed share per warp a
  // quarter of the warp-time went into waiting at the final barrier.  A warp walks a row in groups of 4 consecutive
  // pixels.  The gradient rows (lane 0) and saved rows (lane 1) of a warp's first output row are requested into L2 now,
  // so the classification below overlaps their DRAM latency; taking row r requests row r + nwarps.
  // Otherwise: each warp owns a contiguous range of 4-pixel groups of the row-major pixel list, no prefetch.
  const int npx = wh * wh;
  const int nb = (npx + 3) >> 2;
  const int b0 = (int)(((long long)nb * warp) / nwarps), b1 = (int)(((long long)nb * (warp + 1)) / nwarps);
  // cp.async.bulk.prefetch needs 16-byte aligned addresses: the rows are multiples of 16 bytes here, and the launcher
  // takes the unaligned schedule (no prefetch) when the caller's g_seg / saved base pointers are not (a contiguous
  // autograd view with an odd offset)
  constexpr bool kPrefetch = C32 && ALIGNED;
  (void)pf_ok;
  const unsigned char* pf_base = (lane == 0) ? (LOSS ? reinterpret_cast<const unsigned char*>(aux + (size_t)n * npx)
                                                     : reinterpret_cast<const unsigned char*>(g_seg + (size_t)n * npx * 32))
                                             : saved + (size_t)n * npx * 32;
  const uint32_t pf_row = (uint32_t)wh * ((lane == 0) ? (LOSS ? 16u : 128u) : 32u);   // bytes per output row
  if (kPrefetch && lane < 2 && warp < wh) prefetch_l2_bulk(pf_base + (size_t)warp * pf_row, pf_row);
  const float* proj_n = projects + (size_t)n * Vs * 3;
  const float* mask_n = mask + (size_t)n * Vs;
  float* out = g_projects + (size_t)n * Vs * 3;
  for (int i = threadIdx.x; i < Vs * 3; i += blockDim.x) out[i] = 0.f;   // z and untouched vertices stay 0
  for (int i = threadIdx.x; i < nwarps * (kAccRows * 32 + kOvPriv); i += blockDim.x) b.wacc[i] = make_float2(0.f, 0.f);   // + wov
  for (int i = threadIdx.x; i < 32; i += blockDim.x) b.lpos[kIL * 32 + i] = make_float2(0.f, 0.f);
  if (threadIdx.x == 0) *b.next_row = nwarps;                       // rows 0 .. nwarps-1 are the warps' first rows
  classify_light(b, proj_n, mask_n, ptr, idx, obase, P, OV);
  __syncthreads();
  if (threadIdx.x < 32) {                                           // this sample's overflow slots, handed out in part order
    const int c = max(b.lcount[threadIdx.x] - kIL, 0);
    int sc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, sc, o);
      if ((int)threadIdx.x >= o) sc += t;
    }
    b.odyn[threadIdx.x] = sc - c;
  }
  __syncthreads();
  const int C = C32 ? 32 : P + 1;
  const bool live = lane >= 1 && lane < C;                          // lane = channel; channel 0 carries the gate
  const int k = live ? lane - 1 : 0;
  const int p0 = ptr[k], p1 = ptr[k + 1];
  // column = lane: lane 0 (the gate channel) and dead lanes own harmless columns of their own, so no lane needs masking
  const uint32_t lpos_sa = (uint32_t)__cvta_generic_to_shared(b.lpos) + (uint32_t)lane * 8u;
  const uint32_t wacc_sa = (uint32_t)__cvta_generic_to_shared(b.wacc + (size_t)warp * kAccRows * 32) + (uint32_t)lane * 8u;
  const int ob = b.obase[k], od = b.odyn[k];
  float2* wov_w = b.wov + (size_t)warp * kOvPriv;
  const unsigned char* sv = saved + (size_t)n * npx * 32 + lane;
  const float* g_n = LOSS ? nullptr : g_seg + (size_t)n * npx * C + (lane < C ? lane : 0);
  const bool ld_g = C32 || lane < C;
  // LOSS: this lane fetches the record of pixel (lane & 3) of every group; lanes 0..3 publish {x g - z g, -z g, y g, label}
  // of their pixel in a 64-byte per-warp scratch at the head of the block's shared memory (kBwdHeader), which every lane
  // then reads as a broadcast: one LDS.128 per pixel instead of five shuffles (32 lanes per clock per SM)
  const float4* ax_n = LOSS ? aux + (size_t)n * npx + (lane & 3) : nullptr;
  const float* gl_n = LOSS ? g_loss + (size_t)n * npx + (lane & 3) : nullptr;
  const float lane_f = (float)lane;
  float4* scr = reinterpret_cast<float4*>(raw) + (size_t)warp * 4;   // [4] records of the group being computed

  // Register loads run one group ahead in two named register sets (no copies), pointers advance linearly.
  const int px0 = ALIGNED ? warp * wh : b0 * 4;                     // first pixel of this warp
  const unsigned char* svp = sv + (size_t)px0 * 32;
  const float* gp = LOSS ? nullptr : g_n + (size_t)px0 * C;
  const float4* axp = LOSS ? ax_n + px0 : nullptr;
  const float* glp = LOSS ? gl_n + px0 : nullptr;
  int px = px0;                                                     // first pixel of the group being COMPUTED (!ALIGNED)
  int pxl = px;                                                     // first pixel of the group being LOADED (!ALIGNED)
  int rL = warp, gL = 0;                                            // ALIGNED: load cursor (output row, group in the row)
  int rC = warp, col = 0;                                           // ALIGNED: compute cursor (output row, column)
  const int G = wh >> 2;
  // (column, grid row) of the group's four pixels as packed pairs (rows flipped, :68); ALIGNED: carried across groups
  f32x2 GP[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) GP[j] = pk2((float)j, (float)(wh - 1 - rC));
  const f32x2 step_in = pk2(4.0f, 0.0f);
#define SEG_LOAD4(code, g, ax)                                                                                         \
  do {                                                                                                                 \
    _Pragma("unroll") for (int j = 0; j < 4; ++j) {                                                                    \
      const bool in = ALIGNED || pxl + j < npx;                                                                        \
      code[j] = in ? (int)svp[j * 32] : 0;                                                                             \
      if (!LOSS) g[j] = (in && ld_g) ? gp[j * C] : 0.f;                                                                \
    }                                                                                                                  \
    if (LOSS) {                                                     /* pixel (lane & 3) of the group: record + dL/dloss */ \
      const bool in = ALIGNED || pxl + (lane & 3) < npx;                                                               \
      ax = in ? *axp : make_float4(0.f, 0.f, 0.f, 0.f);                                                                \
      g[0] = in ? *glp : 0.f;                                                                                          \
      axp += 4; glp += 4;                                                                                              \
    } else {                                                                                                           \
      gp += 4 * C;                                                                                                     \
    }                                                                                                                  \
    svp += 4 * 32; pxl += 4;                                                                                           \
  } while (0)
#define SEG_COMPUTE4_CORE(code, g, ax)                                                                                 \
  do {                                                                                                                 \
    if (!ALIGNED) {                                                                                                    \
      _Pragma("unroll") for (int j = 0; j < 4; ++j) {                                                                  \
        const int r = (px + j) / wh;                                                                                   \
        GP[j] = pk2((float)(px + j - r * wh), (float)(wh - 1 - r));                                                    \
      }                                                                                                                \
    }                                                                                                                  \
    f32x2 c2[4];                                                                                                       \
    float Gv[4], Gb[LOSS ? 4 : 1];                                  /* upstream gradient of the score: Gv - Gb exp(s) */ \
    int li[4];                                                                                                         \
    uint32_t row[4];                                                                                                   \
    if (LOSS) {                                                     /* lanes 0..3 publish their pixel's record */      \
      __syncwarp();                                                 /* the previous group's reads are done */          \
      if (lane < 4) {                                                                                                  \
        const float zg = ax.z * g[0];                                                                                  \
        scr[lane] = make_float4(fmaf(ax.x, g[0], -zg), -zg, ax.y * g[0], ax.w);                                        \
      }                                                                                                                \
      __syncwarp();                                                                                                    \
    }                                                                                                                  \
    /* (a) four pixels, mutually independent: the arithmetic of the four chains interleaves */                        \
    _Pragma("unroll") for (int j = 0; j < 4; ++j) {                                                                    \
      if (LOSS) {                                                                                                      \
        const float4 rj = scr[j];                                   /* same address on every lane: one broadcast */    \
        Gb[j] = rj.z;                                                                                                  \
        Gv[j] = (rj.w == lane_f) ? rj.x : rj.y;                                                                        \
      } else {                                                                                                         \
        const float t0 = (code[j] & 1) ? g[j] : 0.f;               /* lane 0: gate * g_bg   (d bg / d s_k = -gate) */  \
        Gv[j] = g[j] - __shfl_sync(0xffffffffu, t0, 0);                                                                \
      }                                                                                                                \
      li[j] = code[j] - 1;                                          /* -1 none, >= kIL overflow, 254 re-query */       \
      row[j] = min((unsigned)li[j], (unsigned)kIL) * 256u;          /* "none" and the rare codes: row kIL */           \
      const f32x2 e = lds_b64_nv(lpos_sa + row[j]);                                                                    \
      const f32x2 d = sub2(e, GP[j]);                               /* (du, dv) */                                      \
      float du2, dv2;                                                                                                  \
      upk2(mul2(d, d), du2, dv2);                                                                                      \
      const float d2 = __fadd_rn(du2, dv2);                                                                            \
      const float rs = rsqrt_approx(fmaxf(d2, 1e-30f));                                                                \
      const float s = ex2_approx((d2 * rs) * (-kLog2e));                                                               \
      const float Gj = LOSS ? fmaf(-Gb[LOSS ? j : 0], ex2_approx(s * kLog2e), Gv[j]) : Gv[j];                          \
      const float coef = (s * Gj) * (-rs);                          /* -s (p - g)/d ; d == 0 -> du = dv = 0 -> 0 */    \
      c2[j] = mul2(pk2(coef, coef), d);                                                                                \
    }                                                                                                                  \
    /* (b) the four read-modify-writes of this lane's private slots (it is their only writer), in order; "none" and  */\
    /* the rare codes add into the discarded row kIL, so the common path has no branch                               */\
    _Pragma("unroll") for (int j = 0; j < 4; ++j) {                                                                    \
      const uint32_t a_sa = wacc_sa + row[j];                                                                          \
      sts_b64(a_sa, add2(lds_b64(a_sa), c2[j]));                                                                       \
    }                                                                                                                  \
    if (__any_sync(0xffffffffu, live && max(max(li[0], li[1]), max(li[2], li[3])) >= kIL)) {                           \
      _Pragma("unroll") for (int j = 0; j < 4; ++j) {                                                                  \
        if (live && li[j] >= kIL) {                                 /* rare: overflow slot, or re-query (code 255) */  \
          float gxj, gyj;                                                                                              \
          upk2(GP[j], gxj, gyj);                                                                                       \
          const float gbj = LOSS ? Gb[LOSS ? j : 0] : 0.f;                                                             \
          if (li[j] == 254) slow_pixel_grad_global(proj_n, mask_n, idx, p0, p1, gxj, gyj, Gv[j], gbj, out);            \
          else overflow_pixel_grad<LOSS>(b, wov_w, ob, od, li[j] - kIL, gxj, gyj, Gv[j], gbj);                         \
        }                                                                                                              \
      }                                                                                                                \
    }                                                                                                                  \
  } while (0)
#define SEG_COMPUTE4(code, g, ax)                                                                                      \
  do {                                                                                                                 \
    SEG_COMPUTE4_CORE(code, g, ax);                                                                                    \
    px += 4;                                                                                                           \
    if (ALIGNED) {                                                  /* next group of the row, or the warp's next row */ \
      col += 4;                                                                                                        \
      const bool wrap = col == wh;                                                                                     \
      col = wrap ? 0 : col;                                                                                            \
      if (wrap) {                                                   /* the load cursor is one group ahead: on my next row */ \
        rC = rL;                                                                                                       \
        _Pragma("unroll") for (int j = 0; j < 4; ++j) GP[j] = pk2((float)j, (float)(wh - 1 - rC));                     \
      } else {                                                                                                         \
        _Pragma("unroll") for (int j = 0; j < 4; ++j) GP[j] = add2(GP[j], step_in);                                    \
      }                                                                                                                \
    }                                                                                                                  \
  } while (0)
// ALIGNED: move the load cursor one group on; at a row's end take the next free row and request the row nwarps further
#define SEG_ADVANCE_LOAD()                                                                                             \
  do {                                                                                                                 \
    if (++gL == G && rL < wh) {                                                                                        \
      gL = 0;                                                                                                          \
      int r = 0;                                                                                                       \
      if (lane == 0) r = atomicAdd(b.next_row, 1);                                                                     \
      rL = __shfl_sync(0xffffffffu, r, 0);                                                                             \
      svp = sv + (size_t)rL * wh * 32;                                                                                 \
      if (LOSS) { axp = ax_n + (size_t)rL * wh; glp = gl_n + (size_t)rL * wh; }                                        \
      else gp = g_n + (size_t)rL * wh * C;                                                                             \
      if (kPrefetch && lane < 2 && rL + nwarps < wh) prefetch_l2_bulk(pf_base + (size_t)(rL + nwarps) * pf_row, pf_row); \
    }                                                                                                                  \
  } while (0)
  {
    int codeA[4], codeB[4];
    float gA[4], gB[4];
    float4 axA = make_float4(0.f, 0.f, 0.f, 0.f), axB = axA;
    if (ALIGNED && WH != 0 && ((WH >> 2) & 1) == 0) {
      // img_wh known at compile time with an even number of groups per row: the two register sets alternate in step with
      // the rows, so a row is a counted loop of G/2 (A, B) pairs and the per-group cursor bookkeeping of the general
      // schedule below (column wrap, load-cursor compare, pointer rebuild: ~60 of ~220 instructions per group) goes away.
      // The warp's next row is taken -- and the row nwarps further requested into L2 -- at the start of a row's last
      // pair, as late as the general schedule takes it.
      constexpr int kPairs = WH >> 3;
      if (rC < wh) {
        if (kPrefetch && lane < 2 && rC + nwarps < wh) prefetch_l2_bulk(pf_base + (size_t)(rC + nwarps) * pf_row, pf_row);
        SEG_LOAD4(codeA, gA, axA);
        for (;;) {
          int rN = wh;
#pragma unroll 1
          for (int pr = 0; pr < kPairs; ++pr) {
            const bool last = pr == kPairs - 1;
            int r = 0;
            if (last && lane == 0) r = atomicAdd(b.next_row, 1);    // consumed after the next compute: its latency is hidden
            SEG_LOAD4(codeB, gB, axB);
            SEG_COMPUTE4_CORE(codeA, gA, axA);
#pragma unroll
            for (int j = 0; j < 4; ++j) GP[j] = add2(GP[j], step_in);
            if (last) {                                             // the next group is the first of the warp's next row
              rN = __shfl_sync(0xffffffffu, r, 0);
              if (kPrefetch && lane < 2 && rN + nwarps < wh) prefetch_l2_bulk(pf_base + (size_t)(rN + nwarps) * pf_row, pf_row);
              svp = sv + (size_t)rN * wh * 32;
              if (LOSS) { axp = ax_n + (size_t)rN * wh; glp = gl_n + (size_t)rN * wh; }
              else gp = g_n + (size_t)rN * wh * C;
            }
            if (!last || rN < wh) SEG_LOAD4(codeA, gA, axA);
            SEG_COMPUTE4_CORE(codeB, gB, axB);
#pragma unroll
            for (int j = 0; j < 4; ++j) GP[j] = add2(GP[j], step_in);
          }
          if (rN >= wh) break;
          rC = rN;
#pragma unroll
          for (int j = 0; j < 4; ++j) GP[j] = pk2((float)j, (float)(wh - 1 - rC));
        }
      }
    } else if (ALIGNED) {
      if (rL < wh) {
        if (kPrefetch && lane < 2 && rL + nwarps < wh) prefetch_l2_bulk(pf_base + (size_t)(rL + nwarps) * pf_row, pf_row);
        SEG_LOAD4(codeA, gA, axA);
        for (;;) {
          SEG_ADVANCE_LOAD();
          if (rL < wh) SEG_LOAD4(codeB, gB, axB);
          SEG_COMPUTE4(codeA, gA, axA);
          if (rC >= wh) break;
          SEG_ADVANCE_LOAD();
          if (rL < wh) SEG_LOAD4(codeA, gA, axA);
          SEG_COMPUTE4(codeB, gB, axB);
          if (rC >= wh) break;
        }
      }
    } else {
      const int nbw = b1 - b0;
      if (nbw > 0) SEG_LOAD4(codeA, gA, axA);
      for (int i = 0; i < nbw; i += 2) {
        const bool hasB = i + 1 < nbw;
        if (hasB) SEG_LOAD4(codeB, gB, axB);
        SEG_COMPUTE4(codeA, gA, axA);
        if (hasB) {
          if (i + 2 < nbw) SEG_LOAD4(codeA, gA, axA);
          SEG_COMPUTE4(codeB, gB, axB);
        }
      }
    }
  }
#undef SEG_ADVANCE_LOAD
#undef SEG_LOAD4
#undef SEG_COMPUTE4
#undef SEG_COMPUTE4_CORE
  __syncthreads();
  // fold the warps' private slots and the overflow list into the output (a vertex may sit in more than one part)
  for (int s = threadIdx.x; s < kIL * 32; s += blockDim.x) {
    const int i = s >> 5, kk = (s & 31) - 1;                       // column = part + 1
    if (kk >= 0 && kk < P && i < b.lcount[kk]) {
      float su = 0.f, sv2 = 0.f;
      for (int w = 0; w < nwarps; ++w) { const float2 a = b.wacc[(size_t)w * kAccRows * 32 + s]; su += a.x; sv2 += a.y; }
      const int vid = b.lvid[s];
      atomicAdd(&out[vid * 3], su); atomicAdd(&out[vid * 3 + 1], sv2);
    }
  }
  for (int kk = threadIdx.x; kk < P; kk += blockDim.x) {
    const int cnt = max(b.lcount[kk] - kIL, 0), obk = b.obase[kk], odk = b.odyn[kk];
    for (int j = 0; j < cnt; ++j) {
      float2 a = b.oacc[obk + j];
      if (odk + j < kOvPriv)
        for (int w = 0; w < nwarps; ++w) { const float2 t = b.wov[(size_t)w * kOvPriv + odk + j]; a.x += t.x; a.y += t.y; }
      const int vid = b.ovid[obk + j];
      atomicAdd(&out[vid * 3], a.x); atomicAdd(&out[vid * 3 + 1], a.y);
    }
  }
}

constexpr size_t kMaxSmem = 227 * 1024;

cudaError_t launch_fwd_impl(const SmplB200Parts* p, const float* projects, const float* mask, int N, int Vs, int wh,
                            float* seg, unsigned char* saved, const SegLossArgs* la, cudaStream_t st) {
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
  // Survivor words: one row per warp (pruned per tile, lane = part), or -- when it costs no resident block -- one row
  // per tile of the block, pruned once with lane = tile (see prune_all_tiles).
  const int KW = p->keep_words > 0 ? p->keep_words : seg_keep_words(p->E);
  const size_t fixed = seg_base_smem(p->E, wh) + (size_t)warps * 8 * kNB * 32 * 4;
  const size_t smem_warp = fixed + (((size_t)warps * KW * 4 + 15) & ~(size_t)15);
  if (smem_warp > kMaxSmem) return cudaErrorInvalidConfiguration;
  const int tiles_blk = (g.ntiles + split - 1) / split;
  const size_t smem_table = fixed + (((size_t)tiles_blk * KW * 4 + 15) & ~(size_t)15);
  constexpr size_t kSmSmem = 228 * 1024, kBlkReserve = 1024;   // per SM, and what the driver adds to every block
  const int max_blocks = la ? 3 : 4;                          // the register cap of the kernel variants
  const int blocks_warp = (int)std::min<size_t>(kSmSmem / (smem_warp + kBlkReserve), (size_t)max_blocks);
  // (measured: the fused-loss variant runs 3.76 ms with per-warp rows and 3.86 ms with the table: it keeps the rows)
  const bool table = !la && tiles_blk != warps && smem_table <= kMaxSmem &&
                     (int)std::min<size_t>(kSmSmem / (smem_table + kBlkReserve), (size_t)max_blocks) >= blocks_warp;
  const size_t smem = table ? smem_table : smem_warp;
  const int kw_rows = table ? tiles_blk : warps;
  dim3 grid(N, split);
  LaunchScope scope(KID_SEG_FWD, st);
  const SegLossArgs none{nullptr, nullptr, 0.f, nullptr, nullptr};
#define SMPL_SEG_FWD_K(TR, LO, W, C3)                                                                                   \
  do {                                                                                                                 \
    cudaError_t e = cudaFuncSetAttribute(seg_fwd_kernel<TR, LO, W, C3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                                    \
    seg_fwd_kernel<TR, LO, W, C3><<<grid, warps * 32, smem, st>>>(projects, mask, N, Vs, p->ptr, p->idx, p->P, p->E, wh, seg, \
                                                                  saved, la ? *la : none, KW, kw_rows, 1.0f);               \
  } while (0)
  // the training resolution with the 31-part table: img_wh and the channel count folded into the code
#define SMPL_SEG_FWD(TR, LO)                                                                                           \
  do {                                                                                                                 \
    if (wh == 48 && p->P == 31) SMPL_SEG_FWD_K(TR, LO, 48, true);                                                      \
    else SMPL_SEG_FWD_K(TR, LO, 0, false);                                                                             \
  } while (0)
  if (la) { if (saved) SMPL_SEG_FWD(true, true); else SMPL_SEG_FWD(false, true); }
  else { if (saved) SMPL_SEG_FWD(true, false); else SMPL_SEG_FWD(false, false); }
#undef SMPL_SEG_FWD_K
#undef SMPL_SEG_FWD
  return cudaGetLastError();
}

}  // namespace

size_t seg_saved_bytes(int N, int wh) { return (size_t)N * wh * wh * 32; }

cudaError_t launch_seg_fwd(const SmplB200Parts* p, const float* projects, const float* mask, int N, int Vs, int wh,
                           float* seg, unsigned char* saved, cudaStream_t st) {
  return launch_fwd_impl(p, projects, mask, N, Vs, wh, seg, saved, nullptr, st);
}

cudaError_t launch_seg_loss_fwd(const SmplB200Parts* p, const float* projects, const float* mask, int N, int Vs, int wh,
                                const unsigned char* labels, float gamma, const float* class_w, float* seg, float* loss,
                                unsigned char* saved, void* aux, cudaStream_t st) {
  const SegLossArgs la{labels, class_w, gamma, loss, reinterpret_cast<float4*>(aux)};
  return launch_fwd_impl(p, projects, mask, N, Vs, wh, seg, saved, &la, st);
}

static cudaError_t launch_bwd_impl(const SmplB200Parts* p, const float* projects, const float* mask, const float* g_seg,
                                   const unsigned char* saved, int N, int Vs, int wh, float* g_projects,
                                   const void* aux, const float* g_loss, cudaStream_t st) {
  if (Vs >= (int)kNoVid) return cudaErrorInvalidValue;              // vertex ids are kept as 16 bits
  // the largest warp count whose blocks still fit three to an SM; at least 4
  int warps = 6;                     // measured: 4 x 5 warps and 4 x 4 warps per SM are slower than 3 x 6; 7 warps spill
  while (warps > 4 && 3 * (bwd_smem_bytes(p->ovf, warps) + 1024) > kMaxSmem + 1024) --warps;
  const size_t smem = bwd_smem_bytes(p->ovf, warps);
  if (smem > kMaxSmem) return cudaErrorInvalidConfiguration;
  LaunchScope scope(KID_SEG_BWD, st);
#define SMPL_SEG_BWD_W(C32, AL, LO, W)                                                                                 \
  do {                                                                                                                 \
    cudaError_t e = cudaFuncSetAttribute(seg_bwd_kernel<C32, AL, LO, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                                    \
    seg_bwd_kernel<C32, AL, LO, W><<<N, warps * 32, smem, st>>>(projects, mask, g_seg, saved, N, Vs, p->ptr, p->idx,    \
                                                                 p->obase, p->P, p->ovf, wh, g_projects, pf_ok,        \
                                                                 reinterpret_cast<const float4*>(aux), g_loss);        \
  } while (0)
#define SMPL_SEG_BWD(C32, AL, LO) SMPL_SEG_BWD_W(C32, AL, LO, 0)
  const void* row0 = aux ? aux : (const void*)g_seg;
  const int pf_ok = (reinterpret_cast<uintptr_t>(row0) % 16 == 0 && reinterpret_cast<uintptr_t>(saved) % 16 == 0) ? 1 : 0;
  const bool c32 = p->P == 31, al = wh % 4 == 0 && wh >= 8 && (pf_ok || p->P != 31);
  if (aux) {
    if (c32 && al && wh == 48) SMPL_SEG_BWD_W(true, true, true, 48);   // the training resolution: wh folded into the code
    else if (c32 && al) SMPL_SEG_BWD(true, true, true);
    else if (c32) SMPL_SEG_BWD(true, false, true);
    else if (al) SMPL_SEG_BWD(false, true, true);
    else SMPL_SEG_BWD(false, false, true);
  } else {
    if (c32 && al && wh == 48) SMPL_SEG_BWD_W(true, true, false, 48);
    else if (c32 && al) SMPL_SEG_BWD(true, true, false);
    else if (c32) SMPL_SEG_BWD(true, false, false);
    else if (al) SMPL_SEG_BWD(false, true, false);
    else SMPL_SEG_BWD(false, false, false);
  }
#undef SMPL_SEG_BWD
#undef SMPL_SEG_BWD_W
  return cudaGetLastError();
}

cudaError_t launch_seg_bwd(const SmplB200Parts* p, const float* projects, const float* mask, const float* g_seg,
                           const unsigned char* saved, int N, int Vs, int wh, float* g_projects, cudaStream_t st) {
  return launch_bwd_impl(p, projects, mask, g_seg, saved, N, Vs, wh, g_projects, nullptr, nullptr, st);
}

cudaError_t launch_seg_loss_bwd(const SmplB200Parts* p, const float* projects, const float* mask, const float* g_loss,
                                const unsigned char* saved, const void* aux, int N, int Vs, int wh, float* g_projects,
                                cudaStream_t st) {
  return launch_bwd_impl(p, projects, mask, nullptr, saved, N, Vs, wh, g_projects, aux, g_loss, st);
}

}  // namespace smplb200
