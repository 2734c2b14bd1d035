/*
 * smpl_b200.h — C ABI of the B200-native SMPL decode -> project -> mask -> part-seg / silhouette path.
 *
 * The reference (akashsengupta1997/indirect_learning_pose-shape) has no native code: its hot path is a
 * chain of stock TensorFlow ops driven from Python (keras_smpl/).  This header is therefore the boundary a
 * Python/ctypes (or cffi / pybind) binding of that path binds; every entry point names the reference
 * function (file:line, relative to the reference tree) whose arithmetic it replaces.
 *
 * Conventions
 *  - plain C: opaque handles, raw DEVICE pointers, sizes; no C++ types, no exceptions, no torch types.
 *  - every function returns 0 on success or a negative SmplB200Status; smpl_b200_last_error() returns a
 *    thread-local human-readable message for the last failure on the calling thread.
 *  - all compute entry points are asynchronous on the given cudaStream_t (passed as void*), never allocate,
 *    never synchronise; the caller owns inputs, outputs and the workspace (size from smpl_b200_workspace_bytes).
 *  - there is no CPU fallback: without a CUDA device every create/compute call fails.
 *  - tensors are dense row-major fp32 unless stated.  N = batch, V = 6890 vertices,
 *    Vs = ceil(V / vertex_sampling) sampled vertices (projection.py:67-68), wh = img_wh.
 *  - params layout (N,86): [k_u, k_v, u0, v0 | theta 24x3 axis-angle | beta 10]
 *    (model.py:33-35, batch_smpl.py:98-99, projection.py:62-65).
 */
#ifndef SMPL_B200_H_
#define SMPL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMPL_B200_ABI_VERSION 3
#define SMPL_B200_NUM_PARAMS 86
#define SMPL_B200_NUM_JOINTS 24
#define SMPL_B200_VPOSED_LD(num_verts) ((((num_verts) + 255) / 256) * 768) /* row stride (floats) of the saved v_posed: whole 256-vertex chunks */

#define SMPL_B200_VPS_LD(num_sampled_verts) ((((num_sampled_verts) * 3 + 3) / 4) * 4) /* row stride (floats) of the compact sampled v_posed: 16-byte rows */

typedef enum SmplB200Status {
  SMPL_B200_OK = 0,
  SMPL_B200_ERR_BAD_ARG = -1,      /* null pointer, negative size, misaligned buffer */
  SMPL_B200_ERR_UNSUPPORTED = -2,  /* size outside what the kernels were built for */
  SMPL_B200_ERR_CUDA = -3,         /* a CUDA runtime call failed; message carries cudaGetErrorString */
  SMPL_B200_ERR_WORKSPACE = -4,    /* workspace too small */
  SMPL_B200_ERR_NO_DEVICE = -5     /* no usable CUDA device (there is no CPU path) */
} SmplB200Status;

/* Host-side description of the SMPL constants, in the layouts SMPLLayer.build produces
 * (keras_smpl/batch_smpl.py:31-94).  All pointers are HOST pointers, read during model_create only. */
typedef struct SmplB200HostModel {
  int32_t num_verts;            /* V, 6890 */
  int32_t num_joints;           /* must be 24 */
  int32_t num_betas;            /* must be 10 */
  int32_t num_pose_basis;       /* must be 207 */
  int32_t num_reg_joints;       /* columns of joint_regressor (19 cocoplus / 14 lsp), 0 if absent */
  const float* v_template;      /* [V][3]                      batch_smpl.py:38-41 */
  const float* shapedirs;       /* [10][V*3]                   batch_smpl.py:50-55 */
  const float* posedirs;        /* [207][V*3]                  batch_smpl.py:64-68 */
  const float* J_regressor;     /* [V][24]                     batch_smpl.py:58-61 */
  const float* lbs_weights;     /* [V][24]                     batch_smpl.py:76-79 */
  const int32_t* parents;       /* [24], parents[0] ignored    batch_smpl.py:71 */
  const float* joint_regressor; /* [V][num_reg_joints] or NULL batch_smpl.py:82-87 */
} SmplB200HostModel;

typedef struct SmplB200Model SmplB200Model; /* immutable after create; one per device */
typedef struct SmplB200Parts SmplB200Parts; /* immutable part->vertex table on one device */
typedef struct SmplB200Renderer SmplB200Renderer; /* immutable mesh topology of the visualiser on one device */

/* ---- library -------------------------------------------------------------------------------------- */
int smpl_b200_abi_version(void);
const char* smpl_b200_last_error(void);
/* number of kernel launches issued through this library by the calling process since load (all threads) */
uint64_t smpl_b200_launch_count(void);

/* ---- built-in profiler ------------------------------------------------------------------------------ */
/* While enabled, every kernel launch of the library is bracketed by a cudaEvent pair recorded on the launching
 * stream (the reference's counterpart is the TF FULL_TRACE timeline of profiling_renderer.py:39-50).
 * collect() synchronises the recorded events, returns per-kernel launch counts and summed durations, and resets. */
typedef struct SmplB200KernelStat {
  const char* name;   /* static string, e.g. "seg_fwd" */
  long long launches;
  double total_ms;
} SmplB200KernelStat;
int smpl_b200_profile_enable(int on);
int smpl_b200_profile_collect(SmplB200KernelStat* out, int max_stats, int* num_stats);

/* ---- handles ---------------------------------------------------------------------------------------- */
/* SMPLLayer.__init__/build (batch_smpl.py:24-94): uploads and repacks the constants on `device`.
 * Supported parameter range: for batches of 64 and more the blend products run as fp16-split tensor-core products with
 * the coefficients scaled by 2^6, exact for |beta| <= 1023 (SMPL shape parameters are O(1)); beyond that the coefficient
 * saturates at +-1023.5 (finite output) where the reference would extrapolate linearly.  On any error the half-built
 * handle is released and the caller's current device is restored. */
int smpl_b200_model_create(const SmplB200HostModel* host, int device, SmplB200Model** out);
void smpl_b200_model_destroy(SmplB200Model* model);
int smpl_b200_model_num_verts(const SmplB200Model* model);
int smpl_b200_model_lbs_width(const SmplB200Model* model); /* max non-zeros per LBS weight row */

/* The unpickled part list of projects_to_seg.py:18-24 with indices already divided by vertex_sampling
 * (projects_to_seg.py:36-37), as CSR: part k owns part_idx[part_ptr[k] .. part_ptr[k+1]).  HOST pointers. */
int smpl_b200_parts_create(int device, int num_parts, const int32_t* part_ptr, const int32_t* part_idx,
                           int num_sampled_verts, SmplB200Parts** out);
void smpl_b200_parts_destroy(SmplB200Parts* parts);

/* ---- workspace ---------------------------------------------------------------------------------------- */
typedef enum SmplB200Op {
  SMPL_B200_OP_DECODE_FWD = 0,
  SMPL_B200_OP_DECODE_BWD = 1,
  SMPL_B200_OP_SILHOUETTE_FWD = 2,
  SMPL_B200_OP_SILHOUETTE_BWD = 3,
  SMPL_B200_OP_FULL_FWD = 4,
  SMPL_B200_OP_FULL_BWD = 5
} SmplB200Op;
/* bytes of device scratch the op needs for batch N (16-byte aligned buffer); vertex_sampling <= 1 = none */
size_t smpl_b200_workspace_bytes(const SmplB200Model* model, int op, int N, int img_wh, int vertex_sampling);

/* ---- SMPLLayer.call (batch_smpl.py:96-153) ------------------------------------------------------------- */
/* params (N,86) -> verts (N,V,3).  Optional outputs (NULL to skip):
 *   joints24     (N,24,3)  J_transformed, the side attribute of batch_smpl.py:131
 *   joints_reg   (N,R,3)   the commented-out cocoplus/LSP regression of batch_smpl.py:147-151 (R = num_reg_joints)
 *   v_posed_save (N,LD)    rest-pose vertices after both blend shapes, row stride LD = SMPL_B200_VPOSED_LD(V);
 *                          the tensor decode_bwd needs when a dense vertex gradient may arrive (saved-for-backward).
 *                          NULL: v_posed lives in the workspace only (add N*LD*4 bytes, rounded up to 256, to it)
 *   v_posed_sampled (N, SMPL_B200_VPS_LD(Vs))  compact copy of the rest-pose positions of the SAMPLED vertices (needs
 *                          `projects`): all decode_bwd needs when the gradient arrives through g_projects only -- 16.5
 *                          KB instead of 82.9 KB per sample at vertex_sampling = 5, and a coalesced read
 *   projects     (N,Vs,3)  fused orthographic_project (projection.py:54-81) with `vertex_sampling` */
int smpl_b200_decode_fwd(const SmplB200Model* model, const float* params, int N, float* verts, float* joints24,
                         float* joints_reg, int num_reg_joints_used, float* v_posed_save, float* v_posed_sampled,
                         float* projects, int vertex_sampling, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the above (TF autodiff of batch_smpl.py:96-153 and projection.py:54-81).
 *   g_verts    (N,V,3)  or NULL
 *   g_projects (N,Vs,3) or NULL (gradient w.r.t. the fused projection output, sampled with vertex_sampling)
 *   g_joints24 (N,24,3) or NULL
 *   g_params   (N,86)   written (not accumulated): camera, pose and shape gradients
 * With g_verts == NULL only the sampled vertices carry gradient and the kernels skip the rest; then v_posed_sampled
 * (if given) is read instead of v_posed_save, which may be NULL. */
int smpl_b200_decode_bwd(const SmplB200Model* model, const float* params, int N, const float* v_posed_save,
                         const float* v_posed_sampled, const float* g_verts, const float* g_projects,
                         int vertex_sampling, const float* g_joints24, float* g_params, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ---- orthographic_project (projection.py:54-81), stand-alone ---------------------------------------- */
int smpl_b200_project_fwd(const float* verts, const float* params, int N, int V, int vertex_sampling,
                          float* projects, void* stream);
/* g_verts (N,V,3) is fully written (zeros at unsampled vertices); g_params (N,86) is fully written
 * (zeros outside the four camera slots). */
int smpl_b200_project_bwd(const float* verts, const float* params, const float* g_projects, int N, int V,
                          int vertex_sampling, float* g_verts, float* g_params, void* stream);

/* ---- compute_mask (compute_mask.py:12-108) ----------------------------------------------------------- */
/* projects (N,Vs,3) -> mask (N,Vs) in {1,500}; 64x64 grid, round-half-even, largest z wins, first index on
 * ties, vertex 1 always visible when any grid cell is empty.  Stateless per sample.  No gradient. */
int smpl_b200_mask_fwd(const float* projects, int N, int Vs, float* mask, void* stream);

/* ---- projects_to_seg (projects_to_seg.py:9-69) -------------------------------------------------------- */
/* projects (N,Vs,3), mask (N,Vs) -> seg (N,wh,wh,num_parts+1): channel 0 background, rows flipped.
 * `saved` (nullable) receives what the backward needs instead of a second search: 32 bytes per OUTPUT pixel, laid
 * out like the segmentation itself ([n][row][col][32]) -- byte 0: bit 0 = the clip gate (0 <= sum_k s_k <= 1);
 * byte 1+k: the arg-min of part k as (index into the part's visible-vertex list) + 1, 0 = none, 255 = re-query
 * (heavy / generic winner or index >= 254).  Size: smpl_b200_seg_saved_bytes(); 16-byte aligned.
 * `seg` may be NULL when only `saved` is wanted. */
size_t smpl_b200_seg_saved_bytes(int N, int img_wh);
int smpl_b200_seg_fwd(const SmplB200Parts* parts, const float* projects, const float* mask, int N, int Vs,
                      int img_wh, float* seg, void* saved, void* stream);
/* g_seg (N,wh,wh,P+1) + the forward's `saved` -> g_projects (N,Vs,3), fully written (z column and untouched
 * vertices are 0).  projects and mask must be the forward's inputs. */
int smpl_b200_seg_bwd(const SmplB200Parts* parts, const float* projects, const float* mask, const float* g_seg,
                      const void* saved, int N, int Vs, int img_wh, float* g_projects, void* stream);

/* ---- projects_to_seg fused with the op after it: Reshape -> softmax -> categorical focal loss -------------------------
 * (model.py:119-120, focal_loss.py:10-48.)  Integer labels only: labels (N,wh,wh) uint8 class ids in the OUTPUT's pixel
 * order (rows flipped, i.e. the order of y_true).  loss (N, wh*wh) per pixel, as the reference's loss function returns
 * it.  `seg` may be NULL: a training step never materialises the 128-byte score row per pixel, and the backward reads a
 * 16-byte record per pixel from `state` (smpl_b200_seg_loss_state_bytes, 16-byte aligned) instead of an upstream
 * gradient row.  class_weights: num_parts+1 device floats or NULL (ones).  g_loss (N, wh*wh): upstream gradient of the
 * per-pixel loss.  g_projects (N,Vs,3) fully written. */
size_t smpl_b200_seg_loss_state_bytes(int N, int img_wh);
int smpl_b200_seg_loss_fwd(const SmplB200Parts* parts, const float* projects, const float* mask, int N, int Vs,
                           int img_wh, const uint8_t* labels, float gamma, const float* class_weights, float* seg,
                           float* loss, void* state, void* stream);
int smpl_b200_seg_loss_bwd(const SmplB200Parts* parts, const float* projects, const float* mask, const float* g_loss,
                           const void* state, int N, int Vs, int img_wh, float* g_projects, void* stream);

/* ---- the whole path in one call (model.py:108-118: SMPLLayer -> orthographic_project -> compute_mask -> projects_to_seg)
 * params (N,86) -> projects (N,Vs,3), mask (N,Vs), seg (N,wh,wh,P+1) [+ verts (N,V,3), joints24 (N,24,3); NULL to skip].
 * `state` (nullable; smpl_b200_full_state_bytes, 16-byte aligned) receives what full_bwd needs: the compact sampled
 * v_posed, the 24 bone transforms per sample and the seg arg-min bytes.  NULL = inference: nothing is saved and the rasteriser skips the arg-min tracking.
 * Workspace: smpl_b200_workspace_bytes(model, SMPL_B200_OP_FULL_FWD / _BWD, N, img_wh, vertex_sampling).
 * full_bwd: g_seg (N,wh,wh,P+1) -> g_params (N,86), written; projects and mask are full_fwd's outputs. */
size_t smpl_b200_full_state_bytes(const SmplB200Model* model, int N, int img_wh, int vertex_sampling);
int smpl_b200_full_fwd(const SmplB200Model* model, const SmplB200Parts* parts, const float* params, int N, int img_wh,
                       int vertex_sampling, float* verts, float* joints24, float* projects, float* mask, float* seg,
                       void* state, void* workspace, size_t workspace_bytes, void* stream);
int smpl_b200_full_bwd(const SmplB200Model* model, const SmplB200Parts* parts, const float* params, int N, int img_wh,
                       int vertex_sampling, const float* projects, const float* mask, const float* g_seg,
                       const void* state, float* g_params, void* workspace, size_t workspace_bytes, void* stream);

/* ---- projects_to_silhouette (projects_to_silhouette.py:14-44) ------------------------------------------ */
/* projects (N,Vs,3) -> sil (N,wh,wh,2): channel 0 = 1-s, channel 1 = s, rows flipped.
 * workspace (nullable; smpl_b200_workspace_bytes(model, SMPL_B200_OP_SILHOUETTE_FWD, N, wh, 0) = 2 bytes per pixel): the
 * forward records every pixel's arg-min vertex there; handed to silhouette_bwd with the same projections, the backward
 * streams instead of repeating the nearest-vertex search.  NULL on either side: the backward searches. */
int smpl_b200_silhouette_fwd(const float* projects, int N, int Vs, int img_wh, float* sil, void* workspace,
                             size_t workspace_bytes, void* stream);
int smpl_b200_silhouette_bwd(const float* projects, const float* g_sil, int N, int Vs, int img_wh,
                             float* g_projects, void* workspace, size_t workspace_bytes, void* stream);

/* ---- the op right after the path: Reshape -> softmax -> categorical focal loss ------------------------------ */
/* (model.py:119-120, focal_loss.py:10-48)  seg (num_pixels, C) scores, C <= 32; exactly one of
 * y_true (num_pixels, C) float one-hot / soft labels and labels (num_pixels) uint8 class ids.
 * loss[i] = sum_c w_c (1 - p_c)^gamma (-y_c log p_c), p = clip(softmax(seg_i), 1e-7, 1 - 1e-7);  class_weights: C device
 * floats or NULL (ones; focal_loss.py:19-41 when weight_classes).  from_logits = 0: seg already holds probabilities.
 * bwd: g_seg = g_loss[i] * d loss[i] / d seg (TF autodiff: closed clip interval). */
int smpl_b200_focal_loss_fwd(const float* seg, const float* y_true, const uint8_t* labels, long long num_pixels,
                             int num_classes, float gamma, const float* class_weights, int from_logits, float* loss,
                             void* stream);
int smpl_b200_focal_loss_bwd(const float* seg, const float* y_true, const uint8_t* labels, const float* g_loss,
                             long long num_pixels, int num_classes, float gamma, const float* class_weights,
                             int from_logits, float* g_seg, void* stream);

/* ---- the op right before the path: the regression module's Dense layers (model.py:63-105) --------
 * Keras Dense in fp32: Y (M,out) = act(X (M,in) W (in,out) + bias (out)), act = ReLU (relu != 0) or linear; W is the Keras
 * kernel as stored (in, out), row-major.  Row strides (ldx, ldy, ...) in floats, so a layer can read and write column
 * blocks of a wider state row (the IEF state [features | params], model.py:69,83).  The products run as 3xTF32 split GEMMs on
 * the tensor cores (fp32 accuracy).  Workspace: smpl_b200_dense_workspace_bytes(M, in, out), 256-byte aligned.
 * bwd (TF autodiff): gZ = gY [Y > 0] (ReLU) ; gX (M,in) = gZ W^T (NULL to skip) ; gW (in,out) = X^T gZ and gb (out) = column
 * sums of gZ (NULL to skip), ADDED to their previous contents when accumulate != 0 (the IEF loop shares its three layers
 * across three iterations). */
size_t smpl_b200_dense_workspace_bytes(int M, int in, int out);
int smpl_b200_dense_fwd(const float* X, int ldx, const float* W, const float* bias, int M, int in, int out, int relu, float* Y,
                        int ldy, void* workspace, size_t workspace_bytes, void* stream);
int smpl_b200_dense_bwd(const float* X, int ldx, const float* W, const float* Y, int ldy, const float* gY, int ldg, int M,
                        int in, int out, int relu, float* gX, int ldgx, float* gW, float* gb, int accumulate, void* workspace,
                        size_t workspace_bytes, void* stream);
/* out[r][c] = a[r][c] + scale * d[r][c] over a (rows x cols) block with independent row strides; a or d may be NULL
 * (param_{k+1} = param_k + scaledown * delta_k, model.py:80-82; the column copies that assemble the IEF state). */
int smpl_b200_axpy_cols(const float* a, int lda, const float* d, int ldd, float scale, int rows, int cols, float* out, int ldo,
                        void* stream);

/* ---- the consumer of `verts`: the mesh visualiser (renderer.py:23-115,146-197) ---------------------
 * Replaces SMPLRenderer.__call__ -> render_model -> simple_renderer, i.e. OpenDR's ProjectPoints + LambertianPointLight +
 * ColoredRenderer (third-party, absent from the reference tree; the algorithm restated is listed in
 * csrc/render_kernels.cu and oracle/np_oracle.py:render_mesh).
 * renderer_create: faces (num_faces,3) HOST int32 vertex indices (keras_smpl/smpl_faces.npy, renderer.py:27).
 * render: verts (N,V,3) camera-frame vertices (x right, y down, z forward); cam (N,3) = [f, cx, cy] (renderer.py:54-62);
 * near_far (N,2) clip planes (:64-67; near <= 0 clips at z > 0 only); albedo (3) or (V,3) floats in [0,1]
 * (albedo_per_vertex != 0: the PLY part colours of render_seg, :158-168); lights: num_lights x [x,y,z,r,g,b] HOST floats
 * (:171-195; num_lights = 0 renders the albedo unlit, as render_seg does); background: NULL = white (:153), else
 * (height,width,3) uint8 shared by every image or (N,height,width,3) when background_per_image != 0 (:231-232);
 * image (N,height,width,channels) uint8, channels 3, or 4 with the alpha of get_alpha / append_alpha (:200-218, :252-255).
 * All pointers but faces / lights are DEVICE pointers.  Workspace: smpl_b200_render_workspace_bytes(r, N), 16-byte aligned. */
int smpl_b200_renderer_create(int device, const int32_t* faces, int num_faces, int num_verts, SmplB200Renderer** out);
void smpl_b200_renderer_destroy(SmplB200Renderer* renderer);
size_t smpl_b200_render_workspace_bytes(const SmplB200Renderer* renderer, int N);
int smpl_b200_render(const SmplB200Renderer* renderer, const float* verts, const float* cam, const float* near_far, int N,
                     int height, int width, const float* albedo, int albedo_per_vertex, const float* lights,
                     int num_lights, const uint8_t* background, int background_per_image, int channels, uint8_t* image,
                     void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SMPL_B200_H_ */
