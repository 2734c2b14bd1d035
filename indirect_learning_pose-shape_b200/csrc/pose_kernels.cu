// K2: Rodrigues + folded joint regression + 24-joint kinematic chain, one warp per sample (lane j = joint j),
// parent exchange by __shfl_sync over the tree levels.  Forward and backward.
//
// Reference arithmetic (file:line relative to the reference tree):
//   batch_rodrigues .................... keras_smpl/batch_smpl.py:255-276
//   batch_skew ......................... keras_smpl/batch_smpl.py:230-253
//   joint regression ................... keras_smpl/batch_smpl.py:106-115  (folded: J = Jt + Jd*beta, exact algebra)
//   pose_feature ....................... keras_smpl/batch_smpl.py:122
//   batch_global_rigid_transformation .. keras_smpl/batch_smpl.py:168-228
#include <cuda_fp16.h>
#include "common.cuh"

namespace smplb200 {

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr unsigned kFull = 0xffffffffu;

struct M3 { float m[9]; };
struct V3 { float x, y, z; };

__device__ __forceinline__ M3 shfl_m3(const M3& a, int src) {
  M3 r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.m[i] = __shfl_sync(kFull, a.m[i], src);
  return r;
}
__device__ __forceinline__ V3 shfl_v3(const V3& a, int src) {
  V3 r;
  r.x = __shfl_sync(kFull, a.x, src);
  r.y = __shfl_sync(kFull, a.y, src);
  r.z = __shfl_sync(kFull, a.z, src);
  return r;
}
// C = A * B
__device__ __forceinline__ M3 mul(const M3& a, const M3& b) {
  M3 c;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q) c.m[r * 3 + q] = a.m[r * 3] * b.m[q] + a.m[r * 3 + 1] * b.m[3 + q] + a.m[r * 3 + 2] * b.m[6 + q];
  return c;
}
// C = A * B^T
__device__ __forceinline__ M3 mul_nt(const M3& a, const M3& b) {
  M3 c;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q) c.m[r * 3 + q] = a.m[r * 3] * b.m[q * 3] + a.m[r * 3 + 1] * b.m[q * 3 + 1] + a.m[r * 3 + 2] * b.m[q * 3 + 2];
  return c;
}
// C = A^T * B
__device__ __forceinline__ M3 mul_tn(const M3& a, const M3& b) {
  M3 c;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q) c.m[r * 3 + q] = a.m[r] * b.m[q] + a.m[3 + r] * b.m[3 + q] + a.m[6 + r] * b.m[6 + q];
  return c;
}
__device__ __forceinline__ V3 mulv(const M3& a, const V3& v) {
  return V3{a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z,
            a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z};
}
__device__ __forceinline__ V3 mulv_t(const M3& a, const V3& v) {
  return V3{a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z,
            a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z};
}

struct Rod {   // forward values of one Rodrigues evaluation that the backward needs again
  M3 R;
  V3 r;        // theta / angle
  float angle, c, s;
};

// batch_smpl.py:265-275.  The 1e-8 is added inside the norm only (:265); r divides the un-shifted theta (:266).
// tf.norm = sqrt(sum(x*x)) with separate multiply and add roundings, hence the explicit _rn intrinsics.
__device__ __forceinline__ Rod rodrigues(float tx, float ty, float tz) {
  Rod o;
  const float px = __fadd_rn(tx, 1e-8f), py = __fadd_rn(ty, 1e-8f), pz = __fadd_rn(tz, 1e-8f);
  o.angle = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
  o.r = V3{__fdiv_rn(tx, o.angle), __fdiv_rn(ty, o.angle), __fdiv_rn(tz, o.angle)};
  sincosf(o.angle, &o.s, &o.c);
  const float omc = 1.0f - o.c;
  const float rx = o.r.x, ry = o.r.y, rz = o.r.z;
  // cos*I + (1-cos)*r r^T + sin*skew(r), skew = [[0,-rz,ry],[rz,0,-rx],[-ry,rx,0]] (:238-251)
  o.R.m[0] = __fadd_rn(o.c, __fmul_rn(omc, __fmul_rn(rx, rx)));
  o.R.m[1] = __fadd_rn(__fmul_rn(omc, __fmul_rn(rx, ry)), __fmul_rn(o.s, -rz));
  o.R.m[2] = __fadd_rn(__fmul_rn(omc, __fmul_rn(rx, rz)), __fmul_rn(o.s, ry));
  o.R.m[3] = __fadd_rn(__fmul_rn(omc, __fmul_rn(ry, rx)), __fmul_rn(o.s, rz));
  o.R.m[4] = __fadd_rn(o.c, __fmul_rn(omc, __fmul_rn(ry, ry)));
  o.R.m[5] = __fadd_rn(__fmul_rn(omc, __fmul_rn(ry, rz)), __fmul_rn(o.s, -rx));
  o.R.m[6] = __fadd_rn(__fmul_rn(omc, __fmul_rn(rz, rx)), __fmul_rn(o.s, -ry));
  o.R.m[7] = __fadd_rn(__fmul_rn(omc, __fmul_rn(rz, ry)), __fmul_rn(o.s, rx));
  o.R.m[8] = __fadd_rn(o.c, __fmul_rn(omc, __fmul_rn(rz, rz)));
  return o;
}

// State every lane holds after the forward chain.
struct Chain {
  Rod rod;      // local rotation
  V3 J;         // rest joint
  V3 tl;        // local translation J_j - J_parent (J_0 for the root)
  M3 RG;        // global rotation
  V3 tG;        // global translation (= J_transformed)
  M3 RGp;       // parent's global rotation (identity for the root)
};

__device__ __forceinline__ Chain forward_chain(const float* __restrict__ prm, const float* __restrict__ Jt,
                                               const float* __restrict__ Jd, const TreeInfo& tree, int lane) {
  Chain ch;
  const int j = lane < kJ ? lane : 0;           // lanes 24..31 shadow joint 0 and are never stored
  ch.rod = rodrigues(prm[4 + 3 * j], prm[5 + 3 * j], prm[6 + 3 * j]);
  float beta[kBetas];
#pragma unroll
  for (int k = 0; k < kBetas; ++k) beta[k] = prm[76 + k];
  float jc[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float acc = Jt[j * 3 + c];
#pragma unroll
    for (int k = 0; k < kBetas; ++k) acc = fmaf(Jd[(j * 3 + c) * kBetas + k], beta[k], acc);
    jc[c] = acc;
  }
  ch.J = V3{jc[0], jc[1], jc[2]};
  const int par = tree.parent[j] < 0 ? 0 : tree.parent[j];
  const int dep = tree.depth[j];
  const V3 Jp = shfl_v3(ch.J, par);
  ch.tl = (j == 0) ? ch.J : V3{ch.J.x - Jp.x, ch.J.y - Jp.y, ch.J.z - Jp.z};   // :204, :207
  ch.RG = ch.rod.R;
  ch.tG = ch.tl;
#pragma unroll
  for (int i = 0; i < 9; ++i) ch.RGp.m[i] = (i % 4 == 0) ? 1.f : 0.f;
  for (int d = 1; d <= tree.max_depth; ++d) {   // results[i] = results[parent[i]] * A_here (:209-211), level by level
    const M3 pR = shfl_m3(ch.RG, par);
    const V3 pt = shfl_v3(ch.tG, par);
    if (dep == d) {
      ch.RGp = pR;
      ch.RG = mul(pR, ch.rod.R);
      const V3 t = mulv(pR, ch.tl);
      ch.tG = V3{t.x + pt.x, t.y + pt.y, t.z + pt.z};
    }
  }
  return ch;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
pose_fwd_kernel(const float* __restrict__ params, int N, const float* __restrict__ Jt, const float* __restrict__ Jd,
                const TreeInfo tree, float* __restrict__ X, float* __restrict__ Xlo, int x16, float* __restrict__ A,
                float* __restrict__ Jtr) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (n >= N) return;
  const float* prm = params + (size_t)n * kParams;
  const Chain ch = forward_chain(prm, Jt, Jd, tree, lane);
  float* x = X + (size_t)n * kKPad;
  float* xl = Xlo ? Xlo + (size_t)n * kKPad : nullptr;
  __half* x16h = reinterpret_cast<__half*>(X) + (size_t)n * kKPad;
  __half* x16l = reinterpret_cast<__half*>(Xlo) + (size_t)n * kKPad;
  auto put = [&](int i, float v) {              // tensor-core path: exact split x = hi + lo (TF32, or fp16 of 64 x)
    if (x16) {
      // |64 x| beyond fp16's range (|beta| > 1023: a diverged regressor) saturates to a finite value instead of
      // overflowing to inf / NaN vertices; the supported range is documented in smpl_b200.h
      const float sv = fminf(fmaxf(v * kXScale16, -65504.0f), 65504.0f);
      const __half h = __float2half_rn(sv);
      x16h[i] = h; x16l[i] = __float2half_rn(sv - __half2float(h));
    } else if (xl) { const float h = tf32_hi(v); x[i] = h; xl[i] = v - h; }
    else x[i] = v;
  };
  if (lane < kBetas) put(lane, prm[76 + lane]);
  if (lane < kKPad - kK) put(kK + lane, 0.f);
  if (lane >= 1 && lane < kJ) {                 // pose_feature = Rs[:,1:] - I (:122)
#pragma unroll
    for (int i = 0; i < 9; ++i) put(kBetas + (lane - 1) * 9 + i, ch.rod.R.m[i] - ((i % 4 == 0) ? 1.f : 0.f));
  }
  if (lane < kJ) {
    // A = results - pad(results * [J;0]) (:222-226): rotation block unchanged, translation tG - RG*J
    const V3 b = mulv(ch.RG, ch.J);
    float4* a4 = reinterpret_cast<float4*>(A + ((size_t)n * kJ + lane) * 12);
    a4[0] = make_float4(ch.RG.m[0], ch.RG.m[1], ch.RG.m[2], ch.tG.x - b.x);
    a4[1] = make_float4(ch.RG.m[3], ch.RG.m[4], ch.RG.m[5], ch.tG.y - b.y);
    a4[2] = make_float4(ch.RG.m[6], ch.RG.m[7], ch.RG.m[8], ch.tG.z - b.z);
    if (Jtr) {
      float* jt = Jtr + ((size_t)n * kJ + lane) * 3;  // new_J = results[:,:,:3,3] (:216)
      jt[0] = ch.tG.x; jt[1] = ch.tG.y; jt[2] = ch.tG.z;
    }
  }
}

// Backward of the above: (g_A, g_X, g_Jtr, g_cam partials) -> g_params.
//   g_cam is [cam_chunks][N][4] partial sums written by the LBS backward (summed here in fixed order), or null.
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
pose_bwd_kernel(const float* __restrict__ params, int N, const float* __restrict__ Jt, const float* __restrict__ Jd,
                const TreeInfo tree, const float* __restrict__ gA, const float* __restrict__ gX, int gx_planes,
                const float* __restrict__ gJtr, const float* __restrict__ gcam, int cam_chunks,
                float* __restrict__ gparams) {
  const size_t gx_stride = (size_t)N * kKPad;            // between the K-split planes of gX (added in fixed order)
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (n >= N) return;
  const float* prm = params + (size_t)n * kParams;
  const Chain ch = forward_chain(prm, Jt, Jd, tree, lane);
  const bool live = lane < kJ;
  const int j = live ? lane : 0;
  const int dep = tree.depth[j];

  // ---- seeds from A and J_transformed ---------------------------------------------------------------------
  M3 gRG;
  V3 gtG = V3{0.f, 0.f, 0.f}, gJ = V3{0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 9; ++i) gRG.m[i] = 0.f;
  if (live && gA) {
    const float4* a4 = reinterpret_cast<const float4*>(gA + ((size_t)n * kJ + j) * 12);
    const float4 r0 = a4[0], r1 = a4[1], r2 = a4[2];
    const V3 gAt = V3{r0.w, r1.w, r2.w};
    // A.t = tG - RG*J :  gRG -= gAt (x) J,  gtG += gAt,  gJ -= RG^T gAt
    gRG.m[0] = r0.x - gAt.x * ch.J.x; gRG.m[1] = r0.y - gAt.x * ch.J.y; gRG.m[2] = r0.z - gAt.x * ch.J.z;
    gRG.m[3] = r1.x - gAt.y * ch.J.x; gRG.m[4] = r1.y - gAt.y * ch.J.y; gRG.m[5] = r1.z - gAt.y * ch.J.z;
    gRG.m[6] = r2.x - gAt.z * ch.J.x; gRG.m[7] = r2.y - gAt.z * ch.J.y; gRG.m[8] = r2.z - gAt.z * ch.J.z;
    gtG = gAt;
    const V3 t = mulv_t(ch.RG, gAt);
    gJ = V3{-t.x, -t.y, -t.z};
  }
  if (live && gJtr) {
    const float* g = gJtr + ((size_t)n * kJ + j) * 3;
    gtG.x += g[0]; gtG.y += g[1]; gtG.z += g[2];
  }

  // ---- leaves -> root: parents pull from their children, one tree level at a time -------------------------------
  // A child's whole contribution to its parent, gRG_c R_c^T + gtG_c (x) tl_c  (RG_c = RG_p R_c ; tG_c = RG_p tl_c + tG_p),
  // is formed in the child's lane, so 12 values cross lanes instead of 24, and a (level, child slot) pair that no lane
  // uses is skipped altogether (the standard tree uses 12 of the 36).
  for (int d = tree.max_depth; d >= 1; --d) {
    const M3 a = mul_nt(gRG, ch.rod.R);
    M3 up;
    up.m[0] = a.m[0] + gtG.x * ch.tl.x; up.m[1] = a.m[1] + gtG.x * ch.tl.y; up.m[2] = a.m[2] + gtG.x * ch.tl.z;
    up.m[3] = a.m[3] + gtG.y * ch.tl.x; up.m[4] = a.m[4] + gtG.y * ch.tl.y; up.m[5] = a.m[5] + gtG.y * ch.tl.z;
    up.m[6] = a.m[6] + gtG.z * ch.tl.x; up.m[7] = a.m[7] + gtG.z * ch.tl.y; up.m[8] = a.m[8] + gtG.z * ch.tl.z;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int c = tree.child[j][s];
      const bool take = live && c >= 0 && tree.depth[c < 0 ? 0 : c] == d;
      if (!__any_sync(0xffffffffu, take)) continue;
      const int src = c < 0 ? 0 : c;
      const M3 cu = shfl_m3(up, src);
      const V3 ct = shfl_v3(gtG, src);
      if (take) {
#pragma unroll
        for (int i = 0; i < 9; ++i) gRG.m[i] += cu.m[i];
        gtG.x += ct.x; gtG.y += ct.y; gtG.z += ct.z;
      }
    }
  }
  (void)dep;

  // ---- local rotation / translation gradients --------------------------------------------------------------------
  M3 gR = (j == 0) ? gRG : mul_tn(ch.RGp, gRG);
  V3 gtl = (j == 0) ? gtG : mulv_t(ch.RGp, gtG);
  if (!live) gtl = V3{0.f, 0.f, 0.f};
  gJ.x += gtl.x; gJ.y += gtl.y; gJ.z += gtl.z;          // tl_j = J_j - J_parent  (root: tl_0 = J_0)
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int c = tree.child[j][s];
    const V3 cl = shfl_v3(gtl, c < 0 ? 0 : c);
    if (live && c >= 0) { gJ.x -= cl.x; gJ.y -= cl.y; gJ.z -= cl.z; }
  }
  if (live && j >= 1 && gX) {                             // pose_feature = R_j - I
    const float* g = gX + (size_t)n * kKPad + kBetas + (j - 1) * 9;
    float t9[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) t9[i] = g[i];
    for (int pl = 1; pl < gx_planes; ++pl) {
#pragma unroll
      for (int i = 0; i < 9; ++i) t9[i] += g[(size_t)pl * gx_stride + i];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) gR.m[i] += t9[i];
  }

  // ---- Rodrigues backward (autodiff of batch_smpl.py:265-275) --------------------------------------------------
  {
    const Rod& o = ch.rod;
    const float rx = o.r.x, ry = o.r.y, rz = o.r.z;
    const float omc = 1.0f - o.c;
    const float tr = gR.m[0] + gR.m[4] + gR.m[8];
    // sum_ab gR_ab r_a r_b
    const V3 gr_r = mulv(gR, o.r), gtr_r = mulv_t(gR, o.r);
    const float quad = rx * gr_r.x + ry * gr_r.y + rz * gr_r.z;
    const float gcos = tr - quad;
    const V3 sk = V3{gR.m[7] - gR.m[5], gR.m[2] - gR.m[6], gR.m[3] - gR.m[1]};     // d<gR,skew(r)>/dr
    const float gsin = rx * sk.x + ry * sk.y + rz * sk.z;
    V3 g_r = V3{omc * (gr_r.x + gtr_r.x) + o.s * sk.x, omc * (gr_r.y + gtr_r.y) + o.s * sk.y,
                omc * (gr_r.z + gtr_r.z) + o.s * sk.z};
    const float tx = prm[4 + 3 * j], ty = prm[5 + 3 * j], tz = prm[6 + 3 * j];
    const float inv_a = 1.0f / o.angle;
    // r = theta/angle : g_theta += g_r/angle ; g_angle -= (g_r . theta)/angle^2
    float g_a = -o.s * gcos + o.c * gsin - (g_r.x * tx + g_r.y * ty + g_r.z * tz) * inv_a * inv_a;
    // angle = ||theta + 1e-8|| : d angle / d theta = (theta + 1e-8)/angle
    const float gx = g_r.x * inv_a + g_a * (tx + 1e-8f) * inv_a;
    const float gy = g_r.y * inv_a + g_a * (ty + 1e-8f) * inv_a;
    const float gz = g_r.z * inv_a + g_a * (tz + 1e-8f) * inv_a;
    if (live) {
      float* g = gparams + (size_t)n * kParams + 4 + 3 * j;
      g[0] = gx; g[1] = gy; g[2] = gz;
    }
  }

  // ---- shape: g_beta = gX[0:10] + Jd^T gJ ---------------------------------------------------------------------
  float gb[kBetas];
#pragma unroll
  for (int k = 0; k < kBetas; ++k) {
    float v = 0.f;
    if (live)
      v = Jd[(j * 3 + 0) * kBetas + k] * gJ.x + Jd[(j * 3 + 1) * kBetas + k] * gJ.y + Jd[(j * 3 + 2) * kBetas + k] * gJ.z;
    gb[k] = warp_sum(v);
  }
  if (lane < kBetas) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < kBetas; ++k) if (k == lane) v = gb[k];
    if (gX) {
      float t = gX[(size_t)n * kKPad + lane];
      for (int pl = 1; pl < gx_planes; ++pl) t += gX[(size_t)pl * gx_stride + (size_t)n * kKPad + lane];
      v += t;
    }
    gparams[(size_t)n * kParams + 76 + lane] = v;
  }
  if (lane < 4) {
    float v = 0.f;
    if (gcam)
      for (int c = 0; c < cam_chunks; ++c) v += gcam[((size_t)c * N + n) * 4 + lane];
    gparams[(size_t)n * kParams + lane] = v;
  }
}

}  // namespace

cudaError_t launch_pose_fwd(const SmplB200Model* m, const float* params, int N, float* X, float* X_lo, float* A,
                            float* Jtr, cudaStream_t st) {
  const int blocks = (N + kWarpsPerBlock - 1) / kWarpsPerBlock;
  LaunchScope scope(KID_POSE_FWD, st);
  pose_fwd_kernel<<<blocks, kWarpsPerBlock * 32, 0, st>>>(params, N, m->Jt, m->Jd, m->tree, X, X_lo,
                                                         (X_lo && m->BT16_hi) ? 1 : 0, A, Jtr);
  return cudaGetLastError();
}

cudaError_t launch_pose_bwd(const SmplB200Model* m, const float* params, const float* g_A, const float* g_X,
                            int gx_planes, const float* g_Jtr, const float* g_cam, int cam_chunks, int N,
                            float* g_params, cudaStream_t st) {
  const int blocks = (N + kWarpsPerBlock - 1) / kWarpsPerBlock;
  LaunchScope scope(KID_POSE_BWD, st);
  pose_bwd_kernel<<<blocks, kWarpsPerBlock * 32, 0, st>>>(params, N, m->Jt, m->Jd, m->tree, g_A, g_X, max(gx_planes, 1),
                                                         g_Jtr, g_cam, cam_chunks, g_params);
  return cudaGetLastError();
}

}  // namespace smplb200
