// Shared declarations for the smpl_b200 library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/smpl_b200.h"

namespace smplb200 {

constexpr int kJ = 24;            // joints
constexpr int kBetas = 10;
constexpr int kPoseBasis = 207;
constexpr int kK = 217;           // blend rows actually used: 10 shape + 207 pose
constexpr int kKPad = 224;        // padded blend depth (multiple of 16)
constexpr int kParams = 86;
constexpr int kMaskGrid = 64;     // compute_mask.py:44
constexpr float kMaskInvisible = 500.0f;  // compute_mask.py:68
constexpr int kMaxVsCache = 8;

void set_error(const char* fmt, ...);
void count_launch();

// Kernel ids for the built-in profiler (smpl_b200_profile_*): one per __global__ function of the library.
enum KernelId {
  KID_POSE_FWD = 0, KID_BLEND_FWD, KID_LBS_FWD, KID_JOINTS_REG, KID_LBS_BWD_VERTEX, KID_LBS_BWD_JOINT, KID_BLEND_BWD,
  KID_POSE_BWD, KID_PROJECT_FWD, KID_PROJECT_BWD, KID_MASK, KID_SEG_FWD, KID_SEG_BWD, KID_SIL_FWD, KID_SIL_BWD,
  KID_FOCAL_FWD, KID_FOCAL_BWD, KID_DENSE, KID_RENDER_VERTEX, KID_RENDER_RASTER, KID_RENDER_FACE, KID_COUNT
};
// RAII scope around one kernel launch: counts it and, while profiling is enabled, brackets it with CUDA events
// recorded on the launching stream.
struct LaunchScope {
  LaunchScope(int kid, cudaStream_t st);
  ~LaunchScope();
  int slot;
  cudaStream_t st;
};

struct VsTables {       // per vertex_sampling derived tables (device pointers)
  int vs = 0;           // 0 = unused slot
  int Vs = 0;           // sampled vertex count
  int ncols = 0;        // Vs*3
  int Kp = 0;           // ncols rounded up to a multiple of 32: depth of the backward blend product
  float* BmT = nullptr;       // [Kp][kKPad] : BmT[c][k] = Bm[k][col(c)],  col(c) = 3*vs*(c/3) + c%3 ; zero rows past ncols
  float* Bs_hi = nullptr;     // [kKPad][Kp] : the same matrix K-major for the tensor-core backward, TF32 hi part
  float* Bs_lo = nullptr;     // [kKPad][Kp] : ... and the exact remainder (x = hi + lo)
  int* csc_ptr = nullptr;     // [kJ+1]  joint -> entries (sampled vertices only)
  int* csc_vert = nullptr;    // [nnz]   ORIGINAL vertex id
  int* csc_q = nullptr;       // [nnz]   sampled-space vertex index (vertex id / vs)
  float* csc_w = nullptr;     // [nnz]
  uint8_t* lbs_idx_s = nullptr;  // [Vs][KW] skin-weight joints of the sampled vertices (compact copy of lbs_idx)
  float* lbs_w_s = nullptr;      // [Vs][KW]
};

constexpr int kMaxLights = 8;
struct RenderLights {   // passed by value to the visualiser's vertex kernel; count == 0: unlit (albedo as is)
  int count;
  float pos[kMaxLights][3];
  float color[kMaxLights][3];
};

struct TreeInfo {       // passed by value to the pose kernels
  int parent[kJ];
  int depth[kJ];
  int child[kJ][4];     // up to 4 children, -1 padded (general trees with more children are rejected)
  int max_depth;
};

}  // namespace smplb200

struct SmplB200Model {
  int device = 0;
  int V = 0;
  int LD = 0;            // padded column count of the blend matrix / v_posed rows
  int KW = 0;            // max non-zeros per LBS row
  int R = 0;             // regressed joints
  float* vt_pad = nullptr;   // [LD]
  float* Bm = nullptr;       // [kKPad][LD]
  float* BT_hi = nullptr;    // [LD][kKPad]  Bm transposed (K-major) for the tensor-core forward, TF32 hi part
  float* BT_lo = nullptr;    // [LD][kKPad]  ... and the exact remainder
  uint16_t* BT16_hi = nullptr;  // [LD][kKPad] halfs: fp16(B * bt16_scale)            (fp16-split forward; null = use TF32)
  uint16_t* BT16_lo = nullptr;  // [LD][kKPad] halfs: fp16(B * bt16_scale - hi)
  float bt16_scale = 1.0f;      // power of two that puts max |B| near 2^14
  int num_sms = 148;
  float* Jt = nullptr;       // [kJ*3]
  float* Jd = nullptr;       // [kJ*3][kBetas]
  uint8_t* lbs_idx = nullptr;  // [V][KW]
  float* lbs_w = nullptr;      // [V][KW]
  int* jr_ptr = nullptr;       // [R+1]
  int* jr_vert = nullptr;
  float* jr_w = nullptr;
  smplb200::TreeInfo tree;
  smplb200::VsTables vst[smplb200::kMaxVsCache];
  // host copies kept for lazily building VsTables
  float* h_Bm = nullptr;     // [kK][3V] (host)
  float* h_W = nullptr;      // [V][kJ]  (host)
};

struct SmplB200Parts {
  int device = 0;
  int P = 0;           // parts (31)
  int E = 0;           // total entries
  int Vs = 0;
  int* ptr = nullptr;        // [P+1] device
  int* idx = nullptr;        // [E]   device, sampled-space vertex index
  uint8_t* part_of = nullptr;  // [E] device, part id of each entry
  int max_part = 0;
  int ovf = 0;               // sum_k max(size_k - 32, 0): overflow slots of the seg backward's interleaved light lists
  int* obase = nullptr;      // [P+1] device, exclusive prefix of max(size_k - 32, 0)
  int keep_words = 0;        // 32 + sum_k max(ceil(size_k / 32) - 1, 0): survivor words one tile of the seg forward needs at most
};

struct SmplB200Renderer {   // mesh topology of the visualiser (renderer.py:27), immutable
  int device = 0;
  int V = 0, F = 0;
  int* faces = nullptr;       // [F][3] device
  int* adj_ptr = nullptr;     // [V+1]  device: vertex -> incident faces (ascending face index)
  int* adj_face = nullptr;    // [3F]
  unsigned char* q8 = nullptr;  // [256] device: uint8((k / 255.) * 255.) -- the reference's float image -> uint8 round trip
};

namespace smplb200 {

const VsTables* get_vs_tables(const SmplB200Model* m, int vs);

// kernels launchers (implemented in the *_kernels.cu files). All return cudaError_t of the launch.
// X_lo == null: X receives the blend coefficients.  Otherwise X receives their TF32 hi part and X_lo the remainder.
// X_lo == null: X receives the blend coefficients.  Otherwise the tensor-core split: TF32 hi / remainder as fp32, or
// (model built with the fp16 tables) [N][kKPad] halfs fp16(64 x) / fp16(64 x - hi).
cudaError_t launch_pose_fwd(const SmplB200Model* m, const float* params, int N, float* X, float* X_lo, float* A,
                            float* Jtr, cudaStream_t st);
cudaError_t launch_blend_fwd_tc(const SmplB200Model* m, const float* Xh, const float* Xl, int N, float* v_posed,
                                cudaStream_t st);
constexpr int kMaxBlendBwdSplit = 8;   // K slices of the tensor-core backward blend product = planes of g_X (workspace size)
// g_X: [*planes][N][kKPad] partial sums (planes <= kMaxBlendBwdSplit), added in fixed order by launch_pose_bwd
cudaError_t launch_blend_bwd_tc(const SmplB200Model* m, const VsTables* t, const float* gvp_hi, const float* gvp_lo,
                                size_t gvp_ld, int N, float* g_X, int* planes, cudaStream_t st);
constexpr float kXScale16 = 64.0f;     // blend coefficients are scaled by 2^6 before their fp16 split: |x| < 1023 (pose
                                       // features are bounded by 2, betas by a few units); lo's quantisation is 1e-9 absolute
constexpr int kDenseBatch = 64;   // from this batch on the blend products run on the tensor cores

// exact split of an fp32 value for the 3xTF32 product: hi keeps the 10 mantissa bits TF32 has, lo = x - hi
__host__ __device__ __forceinline__ float tf32_hi(float x) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(__float_as_uint(x) & 0xffffe000u);
#else
  union { float f; unsigned u; } c; c.f = x; c.u &= 0xffffe000u; return c.f;
#endif
}
cudaError_t launch_blend_fwd(const SmplB200Model* m, const float* X, int N, float* v_posed, cudaStream_t st);
// vps (nullable): compact copy [N][vps_ld] of the rest-pose positions of the sampled vertices (the sampled backward's input)
cudaError_t launch_lbs_fwd(const SmplB200Model* m, const float* v_posed, const float* A, const float* params, int N,
                           float* verts, float* projects, int vs, float* vps, int vps_ld, cudaStream_t st);
cudaError_t launch_joints_reg_fwd(const SmplB200Model* m, const float* verts, int N, int R_used, float* joints,
                                  cudaStream_t st);
// t = tables of the PROCESSING stride (1 = every vertex, g_vp rows [LD]; >1 = sampled vertices only, g_vp rows
// [gvp_ld] compact); vs_proj = sampling stride of g_projects.
// g_vp_lo == null: g_vp receives the gradient.  Otherwise g_vp receives its TF32 hi part and g_vp_lo the remainder.
cudaError_t launch_lbs_bwd(const SmplB200Model* m, const VsTables* t, int vs_proj, const float* v_posed,
                           const float* A, const float* params, const float* g_verts, const float* g_projects, int N,
                           float* g_vp, float* g_vp_lo, size_t gvp_ld, float* g_A, float* g_cam, int* cam_chunks,
                           const float* vps, int vps_ld,
                           cudaStream_t st);   // cam_chunks: how many [N][4] partial-sum planes of g_cam were written
                                               // vps (nullable): launch_lbs_fwd's compact copy; v_posed may then be null
cudaError_t launch_blend_bwd(const SmplB200Model* m, const VsTables* t, const float* g_vp, size_t gvp_ld, int N,
                             float* g_X, cudaStream_t st);
// g_cam = [cam_chunks][N][4] partial camera-gradient sums from launch_lbs_bwd (or null)
// g_X = [gx_planes][N][kKPad] partial sums of the blend backward (1 plane unless the tensor-core product was K-split)
cudaError_t launch_pose_bwd(const SmplB200Model* m, const float* params, const float* g_A, const float* g_X,
                            int gx_planes, const float* g_Jtr, const float* g_cam, int cam_chunks, int N,
                            float* g_params, cudaStream_t st);
int lbs_bwd_cam_chunks(int num_processed_verts);
cudaError_t launch_project_fwd(const float* verts, const float* params, int N, int V, int vs, float* projects,
                               cudaStream_t st);
cudaError_t launch_project_bwd(const float* verts, const float* params, const float* g_projects, int N, int V, int vs,
                               float* g_verts, float* g_params, cudaStream_t st);
cudaError_t launch_mask_fwd(const float* projects, int N, int Vs, float* mask, cudaStream_t st);
size_t seg_saved_bytes(int N, int wh);
cudaError_t launch_seg_fwd(const SmplB200Parts* p, const float* projects, const float* mask, int N, int Vs, int wh,
                           float* seg, unsigned char* saved, cudaStream_t st);
cudaError_t launch_seg_bwd(const SmplB200Parts* p, const float* projects, const float* mask, const float* g_seg,
                           const unsigned char* saved, int N, int Vs, int wh, float* g_projects, cudaStream_t st);
// projects_to_seg fused with Reshape -> softmax -> categorical focal loss on integer labels (model.py:119-120,
// focal_loss.py:10-48).  labels [N][wh][wh] in the output's pixel order; aux = [N][wh*wh] float4 (see seg_kernels.cu);
// seg may be null (training never materialises it).
cudaError_t launch_seg_loss_fwd(const SmplB200Parts* p, const float* projects, const float* mask, int N, int Vs, int wh,
                                const unsigned char* labels, float gamma, const float* class_w, float* seg, float* loss,
                                unsigned char* saved, void* aux, cudaStream_t st);
cudaError_t launch_seg_loss_bwd(const SmplB200Parts* p, const float* projects, const float* mask, const float* g_loss,
                                const unsigned char* saved, const void* aux, int N, int Vs, int wh, float* g_projects,
                                cudaStream_t st);
// saved (nullable): [N][wh][wh] uint16 arg-min vertex of every output pixel; given to both, the backward does not search
cudaError_t launch_sil_fwd(const float* projects, int N, int Vs, int wh, float* sil, unsigned short* saved, cudaStream_t st);
cudaError_t launch_sil_bwd(const float* projects, const float* g_sil, int N, int Vs, int wh, float* g_projects,
                           const unsigned short* saved, cudaStream_t st);

cudaError_t launch_focal_loss_fwd(const float* seg, const float* y_true, const uint8_t* labels, long long npix, int C,
                                  float gamma, const float* class_w, int from_logits, float* loss, cudaStream_t st);
cudaError_t launch_focal_loss_bwd(const float* seg, const float* y_true, const uint8_t* labels, const float* g_loss,
                                  long long npix, int C, float gamma, const float* class_w, int from_logits, float* g_seg,
                                  cudaStream_t st);

// ---- the regression module ahead of the decoder (model.py:63-105): Dense layers as 3xTF32 tcgen05 GEMMs ---------------
cudaError_t launch_dense_gemm(const float* Ah, const float* Al, const float* Bh, const float* Bl, int ldk, float* D, int ldd,
                              const float* bias, int M, int Np, int ncols, int K, bool relu, bool accumulate, int num_sms,
                              cudaStream_t st);
int dense_np(int n);
size_t dense_gemm_ws(int M, int N, int K);
cudaError_t launch_dense_fwd(const float* X, int ldx, const float* W, const float* b, int M, int in, int out, bool relu,
                             float* Y, int ldy, void* ws, int num_sms, cudaStream_t st);
cudaError_t launch_dense_bwd(const float* X, int ldx, const float* W, const float* Y, int ldy, const float* gY, int ldg, int M,
                             int in, int out, bool relu, float* gX, int ldgx, float* gW, float* gb, bool accumulate, void* ws,
                             int num_sms, cudaStream_t st);
cudaError_t launch_axpy_cols(const float* a, int lda, const float* d, int ldd, float scale, int rows, int cols, float* out,
                             int ldo, cudaStream_t st);

// mesh visualiser (renderer.py): vscreen / vcolor = [N][V] float4 scratch, fbox = [N][F] packed tile ranges of the faces
cudaError_t launch_render(const SmplB200Renderer* r, const float* verts, const float* cam, const float* near_far, int N,
                          int h, int w, const float* albedo, int albedo_per_vertex, const RenderLights& lights,
                          const unsigned char* background, int bg_per_image, int channels, float4* vscreen, float4* vcolor,
                          uint32_t* fbox, unsigned char* out, cudaStream_t st);

// small device helpers
// Asynchronous request of [p, p + bytes) into L2 (cp.async.bulk.prefetch: no register, no scoreboard; one thread moves a
// whole row).  p 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// the same for any byte range: shrunk to its 16-byte aligned interior, so it never leaves the caller's buffer
__device__ __forceinline__ void prefetch_l2_inner(const void* p, size_t bytes) {
  const uintptr_t a = (reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15;
  const uintptr_t e = (reinterpret_cast<uintptr_t>(p) + bytes) & ~(uintptr_t)15;
  if (e > a) prefetch_l2_bulk(reinterpret_cast<const void*>(a), (uint32_t)(e - a));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace smplb200
