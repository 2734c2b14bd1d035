// K3: linear-blend skinning fused with the weak-perspective projection, forward and backward, plus the
// stand-alone projection kernels and the optional keypoint regression.
//
// Reference arithmetic (file:line relative to the reference tree):
//   skinning  T = W*A, verts = T*[v_posed;1] ... keras_smpl/batch_smpl.py:135-145
//   orthographic_project ....................... keras_smpl/projection.py:54-81
//   (commented) keypoint regression ............ keras_smpl/batch_smpl.py:147-151
//
// Layout: one block owns a chunk of 256 vertices and a group of kGroup samples.  Skin weights (ELL, <= KW
// non-zeros per vertex) stay in registers for the whole group, the group's 24 bone transforms sit in shared
// memory, and each sample's 768-float slice of v_posed / verts moves through a shared staging tile with
// 16-byte coalesced loads and 8-byte coalesced stores (verts rows are only 8-byte aligned: 3*6890*4 % 16 == 8).
#include "common.cuh"

namespace smplb200 {

namespace {

constexpr int kChunk = 256;     // vertices per block
constexpr int kGroup = 8;       // samples per block
constexpr int kARow = kJ * 12;  // floats of bone transforms per sample

template <int KW>
struct Skin {
  float w[KW];
  int j[KW];        // byte offset of the joint's 3x4 transform inside a sample's block of 24 (joint * 48)
};

template <int KW>
__device__ __forceinline__ Skin<KW> load_skin(const uint8_t* __restrict__ idx, const float* __restrict__ w, int v) {
  Skin<KW> s;
#pragma unroll
  for (int k = 0; k < KW; k += 4) {
    const uchar4 i4 = *reinterpret_cast<const uchar4*>(idx + (size_t)v * KW + k);
    const float4 w4 = *reinterpret_cast<const float4*>(w + (size_t)v * KW + k);
    s.j[k] = i4.x * 48; s.j[k + 1] = i4.y * 48; s.j[k + 2] = i4.z * 48; s.j[k + 3] = i4.w * 48;
    s.w[k] = w4.x; s.w[k + 1] = w4.y; s.w[k + 2] = w4.z; s.w[k + 3] = w4.w;
  }
  return s;
}

// explicit shared-window load: through a generic pointer the compiler rebuilds the window base (S2UR + ULEA) in front of
// every access of the loop -- six times per (vertex, sample), a tenth of the forward's stall samples
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 r;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a));
  return r;
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// T (3x4, row-major) = sum_k w_k * A[j_k]   (batch_smpl.py:138-140; zero-weight joints contribute exact zeros)
// As_sa: shared-window address of the sample's 24 transforms
template <int KW>
__device__ __forceinline__ void blend_T(const Skin<KW>& s, uint32_t As_sa, float T[12]) {
#pragma unroll
  for (int e = 0; e < 12; ++e) T[e] = 0.f;
#pragma unroll
  for (int k = 0; k < KW; ++k) {
    const float w = s.w[k];
    if (w == 0.f) continue;      // ELL padding (2.9 of 4 entries are real on average): skipping an exact zero changes nothing
    const uint32_t a = As_sa + (uint32_t)s.j[k];
    const float4 r0 = lds_f4(a), r1 = lds_f4(a + 16), r2 = lds_f4(a + 32);
    T[0] = fmaf(w, r0.x, T[0]); T[1] = fmaf(w, r0.y, T[1]); T[2] = fmaf(w, r0.z, T[2]); T[3] = fmaf(w, r0.w, T[3]);
    T[4] = fmaf(w, r1.x, T[4]); T[5] = fmaf(w, r1.y, T[5]); T[6] = fmaf(w, r1.z, T[6]); T[7] = fmaf(w, r1.w, T[7]);
    T[8] = fmaf(w, r2.x, T[8]); T[9] = fmaf(w, r2.y, T[9]); T[10] = fmaf(w, r2.z, T[10]); T[11] = fmaf(w, r2.w, T[11]);
  }
}

__device__ __forceinline__ void load_group_A(float* As, float* cam, const float* __restrict__ A,
                                             const float* __restrict__ params, int n0, int rows) {
  const float4* src = reinterpret_cast<const float4*>(A + (size_t)n0 * kARow);
  float4* dst = reinterpret_cast<float4*>(As);
  for (int i = threadIdx.x; i < rows * (kARow / 4); i += blockDim.x) dst[i] = src[i];
  if (threadIdx.x < rows * 4) cam[threadIdx.x] = params[(size_t)(n0 + (threadIdx.x >> 2)) * kParams + (threadIdx.x & 3)];
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kLbsStages = 4;   // v_posed slices in flight per block (cp.async ring)

// Forward: each warp owns 32 vertices (96 floats = 384 B per sample) and runs its own cp.async ring, so the only
// block-wide barrier is the one that publishes the group's bone transforms.  (A block-synchronous predecessor spent 2.4
// of ~11 warp-issue-slots at barriers: three per sample and 3 KB moved between them.)
constexpr int kWG = 16;          // samples per block
template <int KW>
__global__ void __launch_bounds__(kChunk)
lbs_fwd_warp_kernel(const float* __restrict__ vp, int LD, const float* __restrict__ A, const float* __restrict__ params,
                    int N, int V, const uint8_t* __restrict__ lbs_idx, const float* __restrict__ lbs_w,
                    float* __restrict__ verts, float* __restrict__ projects, int vs, int Vs,
                    float* __restrict__ vps, int vps_ld) {
  __shared__ __align__(16) float As[kWG * kARow];
  __shared__ float cam[kWG * 4];
  __shared__ __align__(16) float ring[kChunk / 32][kLbsStages][96];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.y * kWG;
  const int rows = min(kWG, N - n0);
  const int v = blockIdx.x * kChunk + tid;
  const bool valid = v < V;
  const int colw = (blockIdx.x * kChunk + warp * 32) * 3;          // this warp's first column of v_posed / verts
  auto issue = [&](int s) {
    if (s < rows && lane < 24) cp_async16(&ring[warp][s % kLbsStages][lane * 4], vp + (size_t)(n0 + s) * LD + colw + lane * 4);
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < kLbsStages - 1; ++s) issue(s);
  {
    const float4* src = reinterpret_cast<const float4*>(A + (size_t)n0 * kARow);
    float4* dst = reinterpret_cast<float4*>(As);
    for (int i = tid; i < rows * (kARow / 4); i += kChunk) dst[i] = src[i];
    if (tid < rows * 4) cam[tid] = params[(size_t)(n0 + (tid >> 2)) * kParams + (tid & 3)];
  }
  const Skin<KW> skin = load_skin<KW>(lbs_idx, lbs_w, valid ? v : 0);
  const bool sampled = valid && projects && (v % vs == 0);
  const bool need = valid && (verts || sampled);
  const int q = v / vs;
  const int lim = (V * 3 - colw) / 2;                              // float2 slots of this warp's slice that exist (V*3, colw even)
  __syncthreads();                                                 // As / cam published; the only block-wide barrier
  const uint32_t As_sa = smem_addr(As);
  for (int s = 0; s < rows; ++s) {
    const int n = n0 + s;
    issue(s + kLbsStages - 1);                                     // keeps kLbsStages-1 slices in flight behind this one
    cp_async_wait<kLbsStages - 1>();
    __syncwarp();                                                  // slice s has landed for the whole warp
    float* st = ring[warp][s % kLbsStages];
    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (need) {
      const float x = st[lane * 3], y = st[lane * 3 + 1], z = st[lane * 3 + 2];
      float T[12];
      blend_T<KW>(skin, As_sa + (uint32_t)(s * kARow * 4), T);
      ox = fmaf(T[0], x, fmaf(T[1], y, fmaf(T[2], z, T[3])));
      oy = fmaf(T[4], x, fmaf(T[5], y, fmaf(T[6], z, T[7])));
      oz = fmaf(T[8], x, fmaf(T[9], y, fmaf(T[10], z, T[11])));
      if (sampled) {                                               // projection.py:77-79: multiply, then add (two roundings)
        float* p = projects + ((size_t)n * Vs + q) * 3;
        p[0] = __fadd_rn(cam[s * 4 + 2], __fmul_rn(ox, cam[s * 4 + 0]));
        p[1] = __fadd_rn(cam[s * 4 + 3], __fmul_rn(oy, cam[s * 4 + 1]));
        p[2] = oz;
        if (vps) {                                                 // compact rest-pose copy: all the sampled backward reads
          float* o = vps + (size_t)n * vps_ld + (size_t)q * 3;
          o[0] = x; o[1] = y; o[2] = z;
        }
      }
    }
    if (verts) {
      __syncwarp();                                                // every lane has read its input before the slot is reused
      st[lane * 3] = ox; st[lane * 3 + 1] = oy; st[lane * 3 + 2] = oz;
      __syncwarp();
      float2* dst = reinterpret_cast<float2*>(verts + (size_t)n * V * 3 + colw);
      if (lane < lim) dst[lane] = reinterpret_cast<const float2*>(st)[lane];
      if (lane < 16 && lane + 32 < lim) dst[lane + 32] = reinterpret_cast<const float2*>(st)[lane + 32];
    }
    __syncwarp();                                                  // the issue() of the next iteration overwrites the slot read 1 ring-turn ago
  }
}

// Gradient arriving at vertex v of sample n: dense g_verts plus the projection's adjoint at sampled vertices.
__device__ __forceinline__ void vertex_grad(const float* __restrict__ g_verts, const float* __restrict__ g_projects,
                                            int n, int v, int V, int vs_proj, int Vs_proj, float ku, float kv,
                                            float& gx, float& gy, float& gz, float& gu, float& gv) {
  gx = gy = gz = gu = gv = 0.f;
  if (g_verts) {
    const float* g = g_verts + ((size_t)n * V + v) * 3;
    gx = g[0]; gy = g[1]; gz = g[2];
  }
  if (g_projects && v % vs_proj == 0) {
    const float* g = g_projects + ((size_t)n * Vs_proj + v / vs_proj) * 3;
    gu = g[0]; gv = g[1];
    gx = fmaf(gu, ku, gx); gy = fmaf(gv, kv, gy); gz += g[2];
  }
}

// Vertex-parallel half of the backward: g_vp = R_T^T g_vert (compact rows of gvp_ld floats, processed vertices only)
// and per-chunk partial sums of the camera gradient.
template <int KW>
__global__ void __launch_bounds__(kChunk)
lbs_bwd_vertex_kernel(const float* __restrict__ vp, int LD, const float* __restrict__ A,
                      const float* __restrict__ params, int N, int V, const uint8_t* __restrict__ lbs_idx,
                      const float* __restrict__ lbs_w, const float* __restrict__ g_verts,
                      const float* __restrict__ g_projects, int vs_proc, int Vp, int vs_proj, int Vs_proj,
                      float* __restrict__ g_vp, float* __restrict__ g_vp_lo, int gvp_ld, int Kp,
                      float* __restrict__ g_cam) {
  __shared__ __align__(16) float As[kGroup * kARow];
  __shared__ float cam[kGroup * 4];
  const int tid = threadIdx.x;
  const int n0 = blockIdx.y * kGroup;
  const int rows = min(kGroup, N - n0);
  const int q = blockIdx.x * kChunk + tid;          // processed-vertex index
  const bool valid = q < Vp;
  const int v = valid ? q * vs_proc : 0;
  load_group_A(As, cam, A, params, n0, rows);
  const Skin<KW> skin = load_skin<KW>(lbs_idx, lbs_w, v);
  __syncthreads();
  for (int s = 0; s < rows; ++s) {
    const int n = n0 + s;
    float c4[4] = {0.f, 0.f, 0.f, 0.f};
    if (valid) {
      const float ku = cam[s * 4], kv = cam[s * 4 + 1];
      float gx, gy, gz, gu, gv;
      vertex_grad(g_verts, g_projects, n, v, V, vs_proj, Vs_proj, ku, kv, gx, gy, gz, gu, gv);
      float T[12];
      blend_T<KW>(skin, smem_addr(As) + (uint32_t)(s * kARow * 4), T);
      float* o = g_vp + (size_t)n * gvp_ld + (size_t)q * 3;
      float r3[3];
      r3[0] = fmaf(T[0], gx, fmaf(T[4], gy, T[8] * gz));
      r3[1] = fmaf(T[1], gx, fmaf(T[5], gy, T[9] * gz));
      r3[2] = fmaf(T[2], gx, fmaf(T[6], gy, T[10] * gz));
      if (g_vp_lo) {                                // tensor-core blend backward: exact TF32 split
        float* ol = g_vp_lo + (size_t)n * gvp_ld + (size_t)q * 3;
#pragma unroll
        for (int e = 0; e < 3; ++e) { const float h = tf32_hi(r3[e]); o[e] = h; ol[e] = r3[e] - h; }
      } else {
        o[0] = r3[0]; o[1] = r3[1]; o[2] = r3[2];
      }
      if (g_projects) {
        const float* p = vp + (size_t)n * LD + (size_t)v * 3;
        const float x = p[0], y = p[1], z = p[2];
        const float ox = fmaf(T[0], x, fmaf(T[1], y, fmaf(T[2], z, T[3])));
        const float oy = fmaf(T[4], x, fmaf(T[5], y, fmaf(T[6], z, T[7])));
        c4[0] = gu * ox; c4[1] = gv * oy; c4[2] = gu; c4[3] = gv;   // u = u0 + x*k_u, v = v0 + y*k_v
      }
    }
    if (blockIdx.x == gridDim.x - 1) {              // zero the K padding the blend backward reads
      const int c = Vp * 3 + tid;
      if (c < Kp) {
        g_vp[(size_t)n * gvp_ld + c] = 0.f;
        if (g_vp_lo) g_vp_lo[(size_t)n * gvp_ld + c] = 0.f;
      }
    }
    // per-warp partial sums of the camera gradient go straight to their own plane: no block barrier inside the sample
    // loop (pose_bwd adds the planes)
#pragma unroll
    for (int k = 0; k < 4; ++k) c4[k] = warp_sum(c4[k]);
    if ((tid & 31) == 0)
      *reinterpret_cast<float4*>(g_cam + (((size_t)blockIdx.x * (kChunk / 32) + (tid >> 5)) * N + n) * 4) =
          make_float4(c4[0], c4[1], c4[2], c4[3]);
  }
}

// Warp sums of TWELVE per-lane values in 13 shuffles instead of 60: at every step a lane keeps half of its values and
// receives its partner's partial sums of that half (12 -> 6 -> 3 -> 2 -> 1), the last step adds the pair.  Twelve even lanes
// end up owning one total each: returns it, with its index in `idx` (-1 on the other lanes).  Fixed order: deterministic.
// (ncu, sampled backward: the 12 x 5 shuffle + add rounds per joint were 27 % of the kernel's instructions.)
__device__ __forceinline__ float warp_sum12(const float (&a)[12], int lane, int& idx) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
  float s6[6], s3[3];
#pragma unroll
  for (int i = 0; i < 6; ++i) s6[i] = (b4 ? a[i + 6] : a[i]) + __shfl_xor_sync(0xffffffffu, b4 ? a[i] : a[i + 6], 16);
#pragma unroll
  for (int i = 0; i < 3; ++i) s3[i] = (b3 ? s6[i + 3] : s6[i]) + __shfl_xor_sync(0xffffffffu, b3 ? s6[i] : s6[i + 3], 8);
  const float t0 = (b2 ? s3[2] : s3[0]) + __shfl_xor_sync(0xffffffffu, b2 ? s3[0] : s3[2], 4);
  const float t1 = (b2 ? 0.f : s3[1]) + __shfl_xor_sync(0xffffffffu, b2 ? s3[1] : 0.f, 4);
  float r = (b1 ? t1 : t0) + __shfl_xor_sync(0xffffffffu, b1 ? t0 : t1, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  idx = ((lane & 1) || (b2 && b1)) ? -1 : (b4 ? 6 : 0) + (b3 ? 3 : 0) + (b2 ? 2 : 0) + (b1 ? 1 : 0);
  return r;
}

// Joint-parallel half: one warp per (sample, joint) walks the joint's vertex list (CSC of the skin weights) and
// reduces g_A[j] = sum_v w_vj * g_vert_v (x) [v_posed_v ; 1].  No atomics: the result is order-deterministic.
__global__ void __launch_bounds__(256)
lbs_bwd_joint_kernel(const float* __restrict__ vp, int LD, const float* __restrict__ params, int N, int V,
                     const int* __restrict__ csc_ptr, const int* __restrict__ csc_vert,
                     const float* __restrict__ csc_w, const float* __restrict__ g_verts,
                     const float* __restrict__ g_projects, int vs_proj, int Vs_proj, float* __restrict__ g_A) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (long long)N * kJ) return;
  const int n = (int)(wid / kJ), j = (int)(wid % kJ);
  const float ku = params[(size_t)n * kParams], kv = params[(size_t)n * kParams + 1];
  float acc[12];
#pragma unroll
  for (int e = 0; e < 12; ++e) acc[e] = 0.f;
  const int e1 = csc_ptr[j + 1];
  for (int e = csc_ptr[j] + lane; e < e1; e += 32) {
    const int v = csc_vert[e];
    const float w = csc_w[e];
    float gx, gy, gz, gu, gv;
    vertex_grad(g_verts, g_projects, n, v, V, vs_proj, Vs_proj, ku, kv, gx, gy, gz, gu, gv);
    const float* p = vp + (size_t)n * LD + (size_t)v * 3;
    const float x = p[0], y = p[1], z = p[2];
    gx *= w; gy *= w; gz *= w;
    acc[0] = fmaf(gx, x, acc[0]); acc[1] = fmaf(gx, y, acc[1]); acc[2] = fmaf(gx, z, acc[2]); acc[3] += gx;
    acc[4] = fmaf(gy, x, acc[4]); acc[5] = fmaf(gy, y, acc[5]); acc[6] = fmaf(gy, z, acc[6]); acc[7] += gy;
    acc[8] = fmaf(gz, x, acc[8]); acc[9] = fmaf(gz, y, acc[9]); acc[10] = fmaf(gz, z, acc[10]); acc[11] += gz;
  }
#pragma unroll
  for (int e = 0; e < 12; ++e) acc[e] = warp_sum(acc[e]);
  if (lane < 3) {
    float4 r;
    if (lane == 0) r = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else if (lane == 1) r = make_float4(acc[4], acc[5], acc[6], acc[7]);
    else r = make_float4(acc[8], acc[9], acc[10], acc[11]);
    reinterpret_cast<float4*>(g_A + ((size_t)n * kJ + j) * 12)[lane] = r;
  }
}

// Sampled backward (the training configuration: the gradient arrives only through the projection of every vs-th
// vertex).  One block per sample, one pass over global memory:
//   phase 1 (thread = sampled vertex): g_vert = (k_u g_u, k_v g_v, g_z), T = sum_k w_k A[j_k], g_vp = R_T^T g_vert (written
//           as an exact TF32 hi/lo pair for the tensor-core blend backward), camera-gradient partials; g_vert and
//           v_posed of the sampled vertices are parked in shared memory;
//   phase 2 (warp = joint, 3 joints per warp): g_A[j] = sum_{v in joint j} w_vj g_vert_v (x) [v_posed_v; 1] from shared
//           memory through the joint's CSC list, warp-reduced: no atomics, deterministic.
template <int KW, bool STAGE>
__global__ void __launch_bounds__(256)
lbs_bwd_sampled_kernel(const float* __restrict__ vp, int LD, const float* __restrict__ A,
                       const float* __restrict__ params, int N, int vs, int Vs, const uint8_t* __restrict__ sidx,
                       const float* __restrict__ sw, const int* __restrict__ csc_ptr, const int* __restrict__ csc_q,
                       const float* __restrict__ csc_w, const float* __restrict__ g_projects,
                       float* __restrict__ g_vp, float* __restrict__ g_vp_lo, int gvp_ld, int Kp,
                       float* __restrict__ g_A, float* __restrict__ g_cam, const float* __restrict__ vps, int vps_ld) {
  extern __shared__ __align__(16) float sm[];
  float* As = sm;                         // [24][12]
  float* sg = As + kARow;                 // [Vs][3] gradient at the sampled vertices   (STAGE only; otherwise phase 2
  float* sp = sg + Vs * 3;                // [Vs][3] their rest-pose positions           re-reads them through L1/L2)
  __shared__ float red[8][4];
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // The sampled vertices sit 12 bytes in every 12 * vs of the v_posed row: the gather below touches every 64-byte DRAM
  // burst of the row anyway, so request the row (and the gradient row) into L2 as two streaming transfers first
  // (0.52 -> 0.50 ms).  With the forward's compact copy `vps` (rows of Vs*3 floats) the gather is a plain coalesced read of
  // 12*Vs bytes instead: ncu showed 2.46x the algorithmic DRAM bytes for the strided gather.
  const int step3 = vps ? 3 : vs * 3;                               // floats between consecutive sampled vertices
  const float* vrow = vps ? vps + (size_t)n * vps_ld : vp + (size_t)n * LD;
  if (tid == 0) prefetch_l2_inner(vrow, vps ? (size_t)Vs * 12 : (size_t)min(LD, Vs * vs * 3) * 4);
  else if (tid == 32) prefetch_l2_inner(g_projects + (size_t)n * Vs * 3, (size_t)Vs * 12);
  for (int i = tid; i < kARow / 4; i += blockDim.x)
    reinterpret_cast<float4*>(As)[i] = reinterpret_cast<const float4*>(A + (size_t)n * kARow)[i];
  const float ku = params[(size_t)n * kParams], kv = params[(size_t)n * kParams + 1];
  __syncthreads();
  const float* gp = g_projects + (size_t)n * Vs * 3;
  const uint32_t As_sa = smem_addr(As);
  float* orow = g_vp + (size_t)n * gvp_ld;
  float* lrow = g_vp_lo ? g_vp_lo + (size_t)n * gvp_ld : nullptr;
  float c4[4] = {0.f, 0.f, 0.f, 0.f};
  for (int q = tid; q < Vs; q += blockDim.x) {
    const Skin<KW> skin = load_skin<KW>(sidx, sw, q);
    const float gu = gp[q * 3], gv = gp[q * 3 + 1], gz = gp[q * 3 + 2];
    const float x = vrow[(size_t)q * step3], y = vrow[(size_t)q * step3 + 1], z = vrow[(size_t)q * step3 + 2];
    const float gx = gu * ku, gy = gv * kv;                       // u = u0 + x k_u, v = v0 + y k_v (projection.py:77-78)
    float T[12];
    blend_T<KW>(skin, As_sa, T);
    float r3[3];
    r3[0] = fmaf(T[0], gx, fmaf(T[4], gy, T[8] * gz));
    r3[1] = fmaf(T[1], gx, fmaf(T[5], gy, T[9] * gz));
    r3[2] = fmaf(T[2], gx, fmaf(T[6], gy, T[10] * gz));
#pragma unroll
    for (int e = 0; e < 3; ++e) {
      if (lrow) { const float h = tf32_hi(r3[e]); orow[q * 3 + e] = h; lrow[q * 3 + e] = r3[e] - h; }
      else orow[q * 3 + e] = r3[e];
    }
    const float ox = fmaf(T[0], x, fmaf(T[1], y, fmaf(T[2], z, T[3])));
    const float oy = fmaf(T[4], x, fmaf(T[5], y, fmaf(T[6], z, T[7])));
    c4[0] = fmaf(gu, ox, c4[0]); c4[1] = fmaf(gv, oy, c4[1]); c4[2] += gu; c4[3] += gv;
    if (STAGE) {
      sg[q * 3] = gx; sg[q * 3 + 1] = gy; sg[q * 3 + 2] = gz;
      sp[q * 3] = x; sp[q * 3 + 1] = y; sp[q * 3 + 2] = z;
    }
  }
  for (int c = Vs * 3 + tid; c < Kp; c += blockDim.x) {           // zero the K padding the blend backward reads
    orow[c] = 0.f;
    if (lrow) lrow[c] = 0.f;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) c4[k] = warp_sum(c4[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) red[warp][k] = c4[k];
  }
  __syncthreads();
  if (tid < 4) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w][tid];
    g_cam[(size_t)n * 4 + tid] = t;
  }
  for (int j = warp; j < kJ; j += (int)(blockDim.x >> 5)) {
    float acc[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) acc[e] = 0.f;
    const int e1 = csc_ptr[j + 1];
    for (int e = csc_ptr[j] + lane; e < e1; e += 32) {
      const int q = csc_q[e];
      const float w = csc_w[e];
      float gx, gy, gz, x, y, z;
      if (STAGE) {
        gx = sg[q * 3] * w; gy = sg[q * 3 + 1] * w; gz = sg[q * 3 + 2] * w;
        x = sp[q * 3]; y = sp[q * 3 + 1]; z = sp[q * 3 + 2];
      } else {
        gx = gp[q * 3] * ku * w; gy = gp[q * 3 + 1] * kv * w; gz = gp[q * 3 + 2] * w;
        x = vrow[(size_t)q * step3]; y = vrow[(size_t)q * step3 + 1]; z = vrow[(size_t)q * step3 + 2];
      }
      acc[0] = fmaf(gx, x, acc[0]); acc[1] = fmaf(gx, y, acc[1]); acc[2] = fmaf(gx, z, acc[2]); acc[3] += gx;
      acc[4] = fmaf(gy, x, acc[4]); acc[5] = fmaf(gy, y, acc[5]); acc[6] = fmaf(gy, z, acc[6]); acc[7] += gy;
      acc[8] = fmaf(gz, x, acc[8]); acc[9] = fmaf(gz, y, acc[9]); acc[10] = fmaf(gz, z, acc[10]); acc[11] += gz;
    }
    if (STAGE) {                       // measured: 0.462 -> 0.421 ms at vertex_sampling = 5 (C5) ...
      int idx;
      const float tot = warp_sum12(acc, lane, idx);
      if (idx >= 0) g_A[((size_t)n * kJ + j) * 12 + idx] = tot;
    } else {                           // ... but 0.586 -> 0.660 ms at full resolution (C3): the plain rounds stay there
#pragma unroll
      for (int e = 0; e < 12; ++e) acc[e] = warp_sum(acc[e]);
      if (lane < 3) {
        float4 r;
        if (lane == 0) r = make_float4(acc[0], acc[1], acc[2], acc[3]);
        else if (lane == 1) r = make_float4(acc[4], acc[5], acc[6], acc[7]);
        else r = make_float4(acc[8], acc[9], acc[10], acc[11]);
        reinterpret_cast<float4*>(g_A + ((size_t)n * kJ + j) * 12)[lane] = r;
      }
    }
  }
}

// ---- stand-alone projection (projection.py:54-81) ---------------------------------------------------------------
__global__ void __launch_bounds__(256)
project_fwd_kernel(const float* __restrict__ verts, const float* __restrict__ params, int N, int V, int vs, int Vs,
                   float* __restrict__ projects) {
  const int n = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Vs) return;
  const float* prm = params + (size_t)n * kParams;
  const float* p = verts + ((size_t)n * V + (size_t)q * vs) * 3;
  float* o = projects + ((size_t)n * Vs + q) * 3;
  o[0] = __fadd_rn(prm[2], __fmul_rn(p[0], prm[0]));
  o[1] = __fadd_rn(prm[3], __fmul_rn(p[1], prm[1]));
  o[2] = p[2];
}

// One block per sample: writes the whole g_verts row (zeros at unsampled vertices) and the whole g_params row.
__global__ void __launch_bounds__(256)
project_bwd_kernel(const float* __restrict__ verts, const float* __restrict__ params,
                   const float* __restrict__ g_projects, int N, int V, int vs, int Vs, float* __restrict__ g_verts,
                   float* __restrict__ g_params) {
  __shared__ float red[8][4];
  const int n = blockIdx.x, tid = threadIdx.x;
  const float ku = params[(size_t)n * kParams], kv = params[(size_t)n * kParams + 1];
  float c4[4] = {0.f, 0.f, 0.f, 0.f};
  for (int v = tid; v < V; v += blockDim.x) {
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (v % vs == 0) {
      const float* g = g_projects + ((size_t)n * Vs + v / vs) * 3;
      const float* p = verts + ((size_t)n * V + v) * 3;
      gx = g[0] * ku; gy = g[1] * kv; gz = g[2];
      c4[0] = fmaf(g[0], p[0], c4[0]); c4[1] = fmaf(g[1], p[1], c4[1]); c4[2] += g[0]; c4[3] += g[1];
    }
    float* o = g_verts + ((size_t)n * V + v) * 3;
    o[0] = gx; o[1] = gy; o[2] = gz;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) c4[k] = warp_sum(c4[k]);
  if ((tid & 31) == 0)
    for (int k = 0; k < 4; ++k) red[tid >> 5][k] = c4[k];
  __syncthreads();
  if (tid < kParams) {
    float t = 0.f;
    if (tid < 4)
      for (int w = 0; w < 8; ++w) t += red[w][tid];
    g_params[(size_t)n * kParams + tid] = t;
  }
}

// ---- optional keypoint regression (the commented lines batch_smpl.py:147-151) -----------------------------------
__global__ void __launch_bounds__(256)
joints_reg_kernel(const float* __restrict__ verts, int N, int V, int R_used, const int* __restrict__ ptr,
                  const int* __restrict__ vert, const float* __restrict__ w, float* __restrict__ joints) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (long long)N * R_used) return;
  const int n = (int)(wid / R_used), r = (int)(wid % R_used);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int e = ptr[r] + lane; e < ptr[r + 1]; e += 32) {
    const float* p = verts + ((size_t)n * V + vert[e]) * 3;
    const float ww = w[e];
    a0 = fmaf(ww, p[0], a0); a1 = fmaf(ww, p[1], a1); a2 = fmaf(ww, p[2], a2);
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
  if (lane == 0) {
    float* o = joints + ((size_t)n * R_used + r) * 3;
    o[0] = a0; o[1] = a1; o[2] = a2;
  }
}

}  // namespace

int lbs_bwd_cam_chunks(int Vp) { return ((Vp + kChunk - 1) / kChunk) * (kChunk / 32); }   // one plane per warp of the vertex kernel

cudaError_t launch_lbs_fwd(const SmplB200Model* m, const float* v_posed, const float* A, const float* params, int N,
                           float* verts, float* projects, int vs, float* vps, int vps_ld, cudaStream_t st) {
  const int V = m->V, Vs = (V + vs - 1) / vs;
  dim3 grid((V + kChunk - 1) / kChunk, (N + kWG - 1) / kWG);
  LaunchScope scope(KID_LBS_FWD, st);
#define SMPL_LBS_FWD(KW)                                                                                          \
  lbs_fwd_warp_kernel<KW><<<grid, kChunk, 0, st>>>(v_posed, m->LD, A, params, N, V, m->lbs_idx, m->lbs_w, verts,    \
                                                   projects, vs, Vs, projects ? vps : nullptr, vps_ld)
  if (m->KW == 4) SMPL_LBS_FWD(4);
  else if (m->KW == 8) SMPL_LBS_FWD(8);
  else SMPL_LBS_FWD(24);
#undef SMPL_LBS_FWD
  return cudaGetLastError();
}

cudaError_t launch_lbs_bwd(const SmplB200Model* m, const VsTables* t, int vs_proj, const float* v_posed,
                           const float* A, const float* params, const float* g_verts, const float* g_projects, int N,
                           float* g_vp, float* g_vp_lo, size_t gvp_ld, float* g_A, float* g_cam, int* cam_chunks,
                           const float* vps, int vps_ld, cudaStream_t st) {
  const int V = m->V, Vp = t->Vs, Vs_proj = (V + vs_proj - 1) / vs_proj;
  if (!g_verts && g_projects && t->vs == vs_proj) {
    // gradient arrives through the projection only: one fused kernel, one block per sample
    const bool stage = (size_t)6 * Vp * sizeof(float) <= 48 * 1024;   // park g_vert / v_posed in smem only when small
    const size_t smem = (size_t)(kARow + (stage ? 6 * Vp : 0)) * sizeof(float);
    LaunchScope scope(KID_LBS_BWD_VERTEX, st);
    *cam_chunks = 1;
#define SMPL_LBS_BWD_S(KW, ST)                                                                                      \
  do {                                                                                                             \
    cudaError_t e = cudaFuncSetAttribute(lbs_bwd_sampled_kernel<KW, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                                \
    lbs_bwd_sampled_kernel<KW, ST><<<N, 256, smem, st>>>(v_posed, m->LD, A, params, N, t->vs, Vp, t->lbs_idx_s,      \
                                                         t->lbs_w_s, t->csc_ptr, t->csc_q, t->csc_w, g_projects, g_vp, \
                                                         g_vp_lo, (int)gvp_ld, t->Kp, g_A, g_cam, vps, vps_ld);     \
  } while (0)
    if (m->KW == 4) { if (stage) SMPL_LBS_BWD_S(4, true); else SMPL_LBS_BWD_S(4, false); }
    else if (m->KW == 8) { if (stage) SMPL_LBS_BWD_S(8, true); else SMPL_LBS_BWD_S(8, false); }
    else { if (stage) SMPL_LBS_BWD_S(24, true); else SMPL_LBS_BWD_S(24, false); }
#undef SMPL_LBS_BWD_S
    return cudaGetLastError();
  }
  *cam_chunks = lbs_bwd_cam_chunks(Vp);
  dim3 grid((Vp + kChunk - 1) / kChunk, (N + kGroup - 1) / kGroup);
  cudaError_t e;
  {
    LaunchScope scope(KID_LBS_BWD_VERTEX, st);
#define SMPL_LBS_BWD(KW)                                                                                           \
  lbs_bwd_vertex_kernel<KW><<<grid, kChunk, 0, st>>>(v_posed, m->LD, A, params, N, V, m->lbs_idx, m->lbs_w, g_verts,   \
                                                     g_projects, t->vs, Vp, vs_proj, Vs_proj, g_vp, g_vp_lo, (int)gvp_ld, \
                                                     t->Kp, g_cam)
  if (m->KW == 4) SMPL_LBS_BWD(4);
  else if (m->KW == 8) SMPL_LBS_BWD(8);
  else SMPL_LBS_BWD(24);
#undef SMPL_LBS_BWD
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) return e;
  const long long warps = (long long)N * kJ;
  LaunchScope scope(KID_LBS_BWD_JOINT, st);
  lbs_bwd_joint_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(v_posed, m->LD, params, N, V, t->csc_ptr, t->csc_vert,
                                                                   t->csc_w, g_verts, g_projects, vs_proj, Vs_proj, g_A);
  return cudaGetLastError();
}

cudaError_t launch_project_fwd(const float* verts, const float* params, int N, int V, int vs, float* projects,
                               cudaStream_t st) {
  const int Vs = (V + vs - 1) / vs;
  for (int n0 = 0; n0 < N; n0 += 65535) {     // gridDim.y limit
    const int nn = min(65535, N - n0);
    dim3 grid((Vs + 255) / 256, nn);
    LaunchScope scope(KID_PROJECT_FWD, st);
    project_fwd_kernel<<<grid, 256, 0, st>>>(verts + (size_t)n0 * V * 3, params + (size_t)n0 * kParams, nn, V, vs, Vs,
                                            projects + (size_t)n0 * Vs * 3);
  }
  return cudaGetLastError();
}

cudaError_t launch_project_bwd(const float* verts, const float* params, const float* g_projects, int N, int V, int vs,
                               float* g_verts, float* g_params, cudaStream_t st) {
  const int Vs = (V + vs - 1) / vs;
  LaunchScope scope(KID_PROJECT_BWD, st);
  project_bwd_kernel<<<N, 256, 0, st>>>(verts, params, g_projects, N, V, vs, Vs, g_verts, g_params);
  return cudaGetLastError();
}

cudaError_t launch_joints_reg_fwd(const SmplB200Model* m, const float* verts, int N, int R_used, float* joints,
                                  cudaStream_t st) {
  const long long warps = (long long)N * R_used;
  LaunchScope scope(KID_JOINTS_REG, st);
  joints_reg_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(verts, N, m->V, R_used, m->jr_ptr, m->jr_vert, m->jr_w,
                                                                joints);
  return cudaGetLastError();
}

}  // namespace smplb200
