// Mesh visualiser: the GPU replacement of the OpenDR path of renderer.py (SMPLRenderer.__call__ :34-85, simple_renderer
// :146-197, render_model :221-256).  OpenDR itself is a third-party dependency that is absent from the reference tree
// (opendr.renderer.ColoredRenderer / opendr.lighting.LambertianPointLight / opendr.camera.ProjectPoints, unpinned); what
// is restated here is its published algorithm at the reference's call sites:
//   * ProjectPoints(f, rt = 0, t = 0, k = 0, c)           x = f X / Z + cx,  y = f Y / Z + cy           (renderer.py:57-62)
//   * VertNormals                                          n_v = normalise(sum over the faces of v of (v1 - v0) x (v2 - v0))
//   * LambertianPointLight(light_pos, vc, light_color)     max(n_v . normalise(light_pos - v), 0) * vc * light_color,
//                                                          three lights summed                           (renderer.py:169-195)
//   * ColoredRenderer: z-buffered triangles with the lit vertex colours interpolated across each face (perspective
//     correct, as the GL pipeline does), colours clamped to [0, 1] per vertex and stored in an 8-bit frame buffer over a
//     white background or the caller's image; pixel (row r, column c) is sampled at the projected position (c, r)
//     (OpenDR shifts the GL frustum by its half-pixel `pixel_center_offset`), fragments outside [near, far] are clipped.
// Not restated: OpenDR's `overdraw` anti-aliasing of silhouette edges (GL line rasterisation) and GL's fixed-point vertex
// snapping -- a visualiser's boundary pixels may differ from a GL driver's by one pixel.
//
// Three kernels.  render_vertex: thread = (image, vertex): normal from the vertex's faces (CSR built at create time, face
// order, so the sum is deterministic), lighting, projection.  render_face: thread = (image, face): the range of 16 x 16
// tiles the face's bounding box can meet, packed in four bytes (an empty range for faces behind the camera or without
// area).  render_raster: block = (tile, image): the faces' packed ranges are streamed in chunks of 256 (one coalesced
// 4-byte load per face and tile; the three corner gathers only for the faces whose range holds the tile -- before, every
// tile gathered all 13776 faces' corners: 661 KB through L2 per tile, 130 MB per image), the ones whose bounding box meets
// the tile (the exact test, as before) are compacted IN FACE ORDER into a shared-memory list (ballot + block scan), and
// every thread resolves its pixel against the list; ties in depth keep the lowest face index (GL_LESS, draw order).
#include <math_constants.h>
#include "common.cuh"

namespace smplb200 {

namespace {

constexpr int kTile = 16;
constexpr int kRasterThreads = kTile * kTile;
constexpr int kListCap = 768;          // faces per shared-memory round (56 B each); a round is rasterised when fewer than 256 slots are left

struct RFace {
  float bx0, bx1, by0, by1;            // bounding box of the corners: a pixel outside it is outside the triangle
  float x0, y0, x1, y1, x2, y2;        // projected corners
  float iz0, iz1, iz2;                 // 1 / Z of the corners (linear in screen space)
  int f;
};

__global__ void __launch_bounds__(256)
render_vertex_kernel(const float* __restrict__ verts, const float* __restrict__ cam, int N, int V,
                     const int* __restrict__ faces, const int* __restrict__ adj_ptr, const int* __restrict__ adj_face,
                     const float* __restrict__ albedo, int albedo_per_vertex, RenderLights lights,
                     float4* __restrict__ vscreen, float4* __restrict__ vcolor) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)N * V) return;
  const int n = (int)(t / V), v = (int)(t - (long long)n * V);
  const float* vn = verts + (size_t)n * V * 3;
  const float X = vn[v * 3], Y = vn[v * 3 + 1], Z = vn[v * 3 + 2];
  float r, g, b;
  if (albedo_per_vertex) { r = albedo[v * 3]; g = albedo[v * 3 + 1]; b = albedo[v * 3 + 2]; }
  else { r = albedo[0]; g = albedo[1]; b = albedo[2]; }
  if (lights.count > 0) {
    float nx = 0.f, ny = 0.f, nz = 0.f;
    for (int e = adj_ptr[v]; e < adj_ptr[v + 1]; ++e) {
      const int f = adj_face[e];
      const int i0 = faces[f * 3], i1 = faces[f * 3 + 1], i2 = faces[f * 3 + 2];
      const float ax = vn[i1 * 3] - vn[i0 * 3], ay = vn[i1 * 3 + 1] - vn[i0 * 3 + 1], az = vn[i1 * 3 + 2] - vn[i0 * 3 + 2];
      const float bx = vn[i2 * 3] - vn[i0 * 3], by = vn[i2 * 3 + 1] - vn[i0 * 3 + 1], bz = vn[i2 * 3 + 2] - vn[i0 * 3 + 2];
      nx += ay * bz - az * by; ny += az * bx - ax * bz; nz += ax * by - ay * bx;
    }
    const float nn = sqrtf(nx * nx + ny * ny + nz * nz);
    const float inv = nn > 0.f ? 1.0f / nn : 0.f;
    nx *= inv; ny *= inv; nz *= inv;
    float sr = 0.f, sg = 0.f, sb = 0.f;
    for (int l = 0; l < lights.count; ++l) {
      const float lx = lights.pos[l][0] - X, ly = lights.pos[l][1] - Y, lz = lights.pos[l][2] - Z;
      const float ll = sqrtf(lx * lx + ly * ly + lz * lz);
      const float d = ll > 0.f ? fmaxf((nx * lx + ny * ly + nz * lz) / ll, 0.f) : 0.f;
      sr += d * r * lights.color[l][0]; sg += d * g * lights.color[l][1]; sb += d * b * lights.color[l][2];
    }
    r = sr; g = sg; b = sb;
  }
  const float f = cam[n * 3], cx = cam[n * 3 + 1], cy = cam[n * 3 + 2];
  const float iz = Z > 0.f ? 1.0f / Z : 0.f;                       // Z <= 0: behind the camera, the face is dropped
  vscreen[t] = make_float4(f * X * iz + cx, f * Y * iz + cy, iz, Z);
  vcolor[t] = make_float4(fminf(fmaxf(r, 0.f), 1.f), fminf(fmaxf(g, 0.f), 1.f), fminf(fmaxf(b, 0.f), 1.f), 0.f);
}

// edge (a -> b) evaluated at p, in coordinates relative to a (small differences: no cancellation of image-sized terms)
__device__ __forceinline__ float edge_fn(float ax, float ay, float bx, float by, float px, float py) {
  return (bx - ax) * (py - ay) - (by - ay) * (px - ax);
}
// fill rule for a sample exactly on an edge (orientation already normalised to positive area in the y-down image
// frame): the edge owns the sample when it runs up the image (dy < 0) or is horizontal running towards -x
__device__ __forceinline__ bool edge_owns(float ax, float ay, float bx, float by) {
  const float dx = bx - ax, dy = by - ay;
  return dy < 0.f || (dy == 0.f && dx < 0.f);
}

// Packed tile range of a face: bytes {tx_lo, tx_hi, ty_lo, ty_hi}; a superset of the tiles its bounding box meets (the
// rasteriser applies the exact test to the faces it keeps), lo > hi for a face that can never be drawn.
constexpr uint32_t kNoTiles = 0x00010001u;   // tx 1..0, ty 1..0
__global__ void __launch_bounds__(256)
render_face_kernel(const float4* __restrict__ vscreen, int N, int V, const int* __restrict__ faces, int F, int tiles_x,
                   int tiles_y, uint32_t* __restrict__ fbox) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)N * F) return;
  const int n = (int)(t / F), f = (int)(t - (long long)n * F);
  const float4* vs = vscreen + (size_t)n * V;
  const float4 a = vs[faces[f * 3]], b = vs[faces[f * 3 + 1]], c = vs[faces[f * 3 + 2]];
  const float xmin = fminf(a.x, fminf(b.x, c.x)), xmax = fmaxf(a.x, fmaxf(b.x, c.x));
  const float ymin = fminf(a.y, fminf(b.y, c.y)), ymax = fmaxf(a.y, fmaxf(b.y, c.y));
  uint32_t box = kNoTiles;
  const bool finite = fabsf(xmin) <= 1e30f && fabsf(xmax) <= 1e30f && fabsf(ymin) <= 1e30f && fabsf(ymax) <= 1e30f;
  if (finite && a.w > 0.f && b.w > 0.f && c.w > 0.f && edge_fn(a.x, a.y, b.x, b.y, c.x, c.y) != 0.f) {
    // tile tx holds pixels 16 tx .. 16 tx + 15: the box meets it iff xmin <= 16 tx + 15 and xmax >= 16 tx; widened by a
    // thousandth of a tile so that rounding here can only add tiles
    const float inv = 1.0f / (float)kTile;
    const int x_lo = max(0, (int)ceilf(fmaxf((xmin - (float)(kTile - 1)) * inv - 1e-3f, -1.0f)));
    const int x_hi = min(tiles_x - 1, (int)floorf(fminf(xmax * inv + 1e-3f, (float)tiles_x)));
    const int y_lo = max(0, (int)ceilf(fmaxf((ymin - (float)(kTile - 1)) * inv - 1e-3f, -1.0f)));
    const int y_hi = min(tiles_y - 1, (int)floorf(fminf(ymax * inv + 1e-3f, (float)tiles_y)));
    if (x_lo <= x_hi && y_lo <= y_hi) box = (uint32_t)x_lo | ((uint32_t)x_hi << 8) | ((uint32_t)y_lo << 16) | ((uint32_t)y_hi << 24);
  }
  fbox[t] = box;
}

__global__ void __launch_bounds__(kRasterThreads)
render_raster_kernel(const float4* __restrict__ vscreen, const float4* __restrict__ vcolor, int N, int V,
                     const int* __restrict__ faces, int F, int h, int w, const float* __restrict__ near_far,
                     const unsigned char* __restrict__ background, int bg_per_image, const unsigned char* __restrict__ q8,
                     int channels, const uint32_t* __restrict__ fbox, unsigned char* __restrict__ out) {
  __shared__ __align__(8) RFace list[kListCap];
  __shared__ int warp_cnt[kRasterThreads / 32];
  __shared__ int list_n;
  const int n = blockIdx.y;
  const int tiles_x = (w + kTile - 1) / kTile;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pc = tx * kTile + (threadIdx.x & (kTile - 1)), pr = ty * kTile + (threadIdx.x >> 4);
  const float px = (float)pc, py = (float)pr;
  const float tx0 = (float)(tx * kTile), tx1 = (float)(tx * kTile + kTile - 1);
  const float ty0 = (float)(ty * kTile), ty1 = (float)(ty * kTile + kTile - 1);
  const float4* vs = vscreen + (size_t)n * V;
  const float zn = near_far[n * 2], zf = near_far[n * 2 + 1];
  // Z in [near, far] <=> 1/Z in [1/far, 1/near]; a non-positive near plane (invalid for glFrustum) clips at Z > 0 only
  const float iz_hi = zn > 0.f ? 1.0f / zn : CUDART_INF_F, iz_lo = zf > 0.f ? 1.0f / zf : 0.f;
  float best_iz = -1.0f, bw0 = 0.f, bw1 = 0.f, bw2 = 0.f;
  int best_f = -1;
  if (threadIdx.x == 0) list_n = 0;
  __syncthreads();

  auto rasterise = [&](const RFace* lst, int count) {
    for (int e = 0; e < count; ++e) {
      // same address on every thread: broadcast.  The box first: SMPL's faces cover a few pixels at 224 x 224, so ~95 % of
      // the (pixel, face) pairs of a tile end here instead of in three edge functions (same result: inside => in the box)
      const float2 bx = *reinterpret_cast<const float2*>(&lst[e].bx0), by = *reinterpret_cast<const float2*>(&lst[e].by0);
      if (px < bx.x || px > bx.y || py < by.x || py > by.y) continue;
      const RFace t = lst[e];
      float w0 = edge_fn(t.x1, t.y1, t.x2, t.y2, px, py);
      float w1 = edge_fn(t.x2, t.y2, t.x0, t.y0, px, py);
      float w2 = edge_fn(t.x0, t.y0, t.x1, t.y1, px, py);
      const float area = edge_fn(t.x0, t.y0, t.x1, t.y1, t.x2, t.y2);
      bool in;
      if (area > 0.f) {
        in = (w0 > 0.f || (w0 == 0.f && edge_owns(t.x1, t.y1, t.x2, t.y2))) &&
             (w1 > 0.f || (w1 == 0.f && edge_owns(t.x2, t.y2, t.x0, t.y0))) &&
             (w2 > 0.f || (w2 == 0.f && edge_owns(t.x0, t.y0, t.x1, t.y1)));
      } else {                                                     // the other winding: the same test on the reversed edges
        in = (w0 < 0.f || (w0 == 0.f && edge_owns(t.x2, t.y2, t.x1, t.y1))) &&
             (w1 < 0.f || (w1 == 0.f && edge_owns(t.x0, t.y0, t.x2, t.y2))) &&
             (w2 < 0.f || (w2 == 0.f && edge_owns(t.x1, t.y1, t.x0, t.y0)));
      }
      if (in) {
        const float s = w0 + w1 + w2;                              // the three edge functions sum to the area (up to rounding)
        const float b0 = w0 / s, b1 = w1 / s, b2 = w2 / s;
        const float iz = b0 * t.iz0 + b1 * t.iz1 + b2 * t.iz2;
        if (iz >= iz_lo && iz <= iz_hi && iz > best_iz) {
          best_iz = iz; best_f = t.f;
          bw0 = b0 * t.iz0; bw1 = b1 * t.iz1; bw2 = b2 * t.iz2;    // perspective-correct weights (before division by iz)
        }
      }
    }
  };

  // With the packed tile ranges: no barrier per chunk.  Each warp owns a contiguous eighth of the faces; it counts the
  // faces whose range holds this tile, the eight counts are scanned once, and each warp then writes its exact survivors
  // at its own offset -- the concatenation is in face order.  (More candidates than the list holds: the streaming path.)
  bool done = false;
  if (fbox) {
    constexpr int kW = kRasterThreads / 32;
    __shared__ int woff[kW + 1];
    const uint32_t* fb = fbox + (size_t)n * F;
    const int per = (F + kW - 1) / kW, f0 = warp * per, f1 = min(F, f0 + per);
    auto in_range = [&](int f) {
      const uint32_t bb = fb[f];
      return (uint32_t)tx >= (bb & 0xffu) && (uint32_t)tx <= ((bb >> 8) & 0xffu) && (uint32_t)ty >= ((bb >> 16) & 0xffu) &&
             (uint32_t)ty <= (bb >> 24);
    };
    int cnt = 0;
    for (int base = f0; base < f1; base += 32) {
      const int f = base + lane;
      cnt += __popc(__ballot_sync(0xffffffffu, f < f1 && in_range(f)));
    }
    if (lane == 0) warp_cnt[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int q = 0; q < kW; ++q) { woff[q] = tot; tot += warp_cnt[q]; }
      woff[kW] = tot;
    }
    __syncthreads();
    if (woff[kW] <= kListCap) {                                    // uniform
      int pos = woff[warp];
      for (int base = f0; base < f1; base += 32) {
        const int f = base + lane;
        bool keep = false;
        RFace t;
        if (f < f1 && in_range(f)) {
          const int i0 = faces[f * 3], i1 = faces[f * 3 + 1], i2 = faces[f * 3 + 2];
          const float4 a = vs[i0], b = vs[i1], c = vs[i2];
          t.x0 = a.x; t.y0 = a.y; t.x1 = b.x; t.y1 = b.y; t.x2 = c.x; t.y2 = c.y;
          t.iz0 = a.z; t.iz1 = b.z; t.iz2 = c.z; t.f = f;
          const float xmin = fminf(a.x, fminf(b.x, c.x)), xmax = fmaxf(a.x, fmaxf(b.x, c.x));
          const float ymin = fminf(a.y, fminf(b.y, c.y)), ymax = fmaxf(a.y, fmaxf(b.y, c.y));
          t.bx0 = xmin; t.bx1 = xmax; t.by0 = ymin; t.by1 = ymax;
          keep = a.w > 0.f && b.w > 0.f && c.w > 0.f && xmin <= tx1 && xmax >= tx0 && ymin <= ty1 && ymax >= ty0 &&
                 edge_fn(a.x, a.y, b.x, b.y, c.x, c.y) != 0.f;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) list[pos + __popc(bal & ((1u << lane) - 1u))] = t;
        pos += __popc(bal);
      }
      __syncthreads();                                             // every warp's counts have been read: reuse them
      if (lane == 0) warp_cnt[warp] = pos - woff[warp];
      __syncthreads();
      for (int q = 0; q < kW; ++q) rasterise(list + woff[q], warp_cnt[q]);
      done = true;
    }
  }
  for (int base = 0; !done && base < F; base += kRasterThreads) {
    if (list_n > kListCap - kRasterThreads) {                      // uniform: list_n is read after a barrier
      rasterise(list, list_n);
      __syncthreads();
      if (threadIdx.x == 0) list_n = 0;
      __syncthreads();
    }
    const int f = base + threadIdx.x;
    bool keep = false;
    RFace t;
    bool maybe = f < F;
    if (maybe && fbox) {                                           // the packed range first: one coalesced load per face
      const uint32_t bb = fbox[(size_t)n * F + f];
      maybe = (uint32_t)tx >= (bb & 0xffu) && (uint32_t)tx <= ((bb >> 8) & 0xffu) && (uint32_t)ty >= ((bb >> 16) & 0xffu) &&
              (uint32_t)ty <= (bb >> 24);
    }
    if (maybe) {
      const int i0 = faces[f * 3], i1 = faces[f * 3 + 1], i2 = faces[f * 3 + 2];
      const float4 a = vs[i0], b = vs[i1], c = vs[i2];
      t.x0 = a.x; t.y0 = a.y; t.x1 = b.x; t.y1 = b.y; t.x2 = c.x; t.y2 = c.y;
      t.iz0 = a.z; t.iz1 = b.z; t.iz2 = c.z; t.f = f;
      const float xmin = fminf(a.x, fminf(b.x, c.x)), xmax = fmaxf(a.x, fmaxf(b.x, c.x));
      const float ymin = fminf(a.y, fminf(b.y, c.y)), ymax = fmaxf(a.y, fmaxf(b.y, c.y));
      t.bx0 = xmin; t.bx1 = xmax; t.by0 = ymin; t.by1 = ymax;
      keep = a.w > 0.f && b.w > 0.f && c.w > 0.f && xmin <= tx1 && xmax >= tx0 && ymin <= ty1 && ymax >= ty0 &&
             edge_fn(a.x, a.y, b.x, b.y, c.x, c.y) != 0.f;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = list_n;
    for (int q = 0; q < warp; ++q) off += warp_cnt[q];
    if (keep) list[off + __popc(bal & ((1u << lane) - 1u))] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int q = 0; q < kRasterThreads / 32; ++q) tot += warp_cnt[q];
      list_n += tot;
    }
    __syncthreads();
  }
  if (!done) rasterise(list, list_n);

  if (pc < w && pr < h) {
    const size_t opx = ((size_t)n * h + pr) * w + pc;
    unsigned char* o = out + opx * channels;
    unsigned char rgb[3];
    if (best_f >= 0) {
      const float4* vc = vcolor + (size_t)n * V;
      const float4 c0 = vc[faces[best_f * 3]], c1 = vc[faces[best_f * 3 + 1]], c2 = vc[faces[best_f * 3 + 2]];
      const float inv = 1.0f / (bw0 + bw1 + bw2);
      const float cr = (bw0 * c0.x + bw1 * c1.x + bw2 * c2.x) * inv;
      const float cg = (bw0 * c0.y + bw1 * c1.y + bw2 * c2.y) * inv;
      const float cb = (bw0 * c0.z + bw1 * c1.z + bw2 * c2.z) * inv;
      // 8-bit frame buffer (round to nearest), then the reference's (k / 255.) * 255 -> uint8 truncation (table q8)
      rgb[0] = q8[(int)rintf(fminf(fmaxf(cr, 0.f), 1.f) * 255.0f)];
      rgb[1] = q8[(int)rintf(fminf(fmaxf(cg, 0.f), 1.f) * 255.0f)];
      rgb[2] = q8[(int)rintf(fminf(fmaxf(cb, 0.f), 1.f) * 255.0f)];
    } else if (background) {
      const unsigned char* bgp = background + ((bg_per_image ? (size_t)n * h * w : 0) + (size_t)pr * w + pc) * 3;
      rgb[0] = q8[bgp[0]]; rgb[1] = q8[bgp[1]]; rgb[2] = q8[bgp[2]];
    } else {
      rgb[0] = rgb[1] = rgb[2] = 255;                              // bgcolor = ones(3), renderer.py:153
    }
    o[0] = rgb[0]; o[1] = rgb[1]; o[2] = rgb[2];
    if (channels == 4) {
      // get_alpha (renderer.py:200-209): transparent where the white-background render is exactly white;
      // append_alpha (:212-218): opaque everywhere when a background image was given
      o[3] = (background || !(rgb[0] == 255 && rgb[1] == 255 && rgb[2] == 255)) ? 255 : 0;
    }
  }
}

}  // namespace

cudaError_t launch_render(const SmplB200Renderer* r, const float* verts, const float* cam, const float* near_far, int N,
                          int h, int w, const float* albedo, int albedo_per_vertex, const RenderLights& lights,
                          const unsigned char* background, int bg_per_image, int channels, float4* vscreen, float4* vcolor,
                          uint32_t* fbox, unsigned char* out, cudaStream_t st) {
  {
    LaunchScope scope(KID_RENDER_VERTEX, st);
    const long long total = (long long)N * r->V;
    render_vertex_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(verts, cam, N, r->V, r->faces, r->adj_ptr,
                                                                         r->adj_face, albedo, albedo_per_vertex, lights,
                                                                         vscreen, vcolor);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  const int tiles_x = (w + kTile - 1) / kTile, tiles_y = (h + kTile - 1) / kTile;
  const bool boxed = fbox && tiles_x <= 255 && tiles_y <= 255;     // a tile index is one byte of the packed range
  if (boxed) {
    LaunchScope scope(KID_RENDER_FACE, st);
    const long long total = (long long)N * r->F;
    render_face_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(vscreen, N, r->V, r->faces, r->F, tiles_x, tiles_y, fbox);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  LaunchScope scope(KID_RENDER_RASTER, st);
  render_raster_kernel<<<dim3(tiles_x * tiles_y, N), kRasterThreads, 0, st>>>(vscreen, vcolor, N, r->V, r->faces, r->F, h, w,
                                                                             near_far, background, bg_per_image, r->q8, channels,
                                                                             boxed ? fbox : nullptr, out);
  return cudaGetLastError();
}

}  // namespace smplb200
