// C-ABI entry points, handles and host-side repacking (see include/smpl_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>
#include <algorithm>
#include <cmath>

#include <cuda_fp16.h>
#include "common.cuh"

namespace smplb200 {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
static std::mutex g_vs_mutex;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- built-in profiler: CUDA-event pairs around every launch, on the launching stream ---------------------------
static const char* const kKernelNames[KID_COUNT] = {
    "pose_fwd", "blend_fwd", "lbs_fwd", "joints_reg", "lbs_bwd_vertex", "lbs_bwd_joint", "blend_bwd", "pose_bwd",
    "project_fwd", "project_bwd", "mask", "seg_fwd", "seg_bwd", "sil_fwd", "sil_bwd", "focal_fwd", "focal_bwd", "dense", "render_vertex", "render_raster", "render_face"};
struct ProfRecord { int kid; cudaEvent_t a, b; };
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mutex;
static std::vector<ProfRecord> g_prof_records;
static std::vector<cudaEvent_t> g_prof_pool;

static cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

LaunchScope::LaunchScope(int kid, cudaStream_t s) : slot(-1), st(s) {
  count_launch();
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  ProfRecord r{kid, prof_event(), prof_event()};
  if (!r.a || !r.b) return;
  cudaEventRecord(r.a, st);
  slot = (int)g_prof_records.size();
  g_prof_records.push_back(r);
}
LaunchScope::~LaunchScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  cudaEventRecord(g_prof_records[slot].b, st);
}

#define CU_TRY(expr)                                                                         \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return SMPL_B200_ERR_CUDA;                                                             \
    }                                                                                        \
  } while (0)

template <typename T>
static cudaError_t upload(T** dptr, const std::vector<T>& h) {
  size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
  cudaError_t e = cudaMalloc((void**)dptr, bytes);
  if (e != cudaSuccess) return e;
  if (!h.empty()) e = cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  return e;
}

static void free_vs_tables(VsTables* t) {
  cudaFree(t->BmT); cudaFree(t->Bs_hi); cudaFree(t->Bs_lo); cudaFree(t->csc_ptr); cudaFree(t->csc_vert); cudaFree(t->csc_w);
  cudaFree(t->csc_q); cudaFree(t->lbs_idx_s); cudaFree(t->lbs_w_s);
  *t = VsTables();
}

static int build_vs_tables(const SmplB200Model* m, int vs, VsTables* t) {
  const int V = m->V;
  const int Vs = (V + vs - 1) / vs;
  const int ncols = Vs * 3;
  const int Kp = (ncols + 31) / 32 * 32;
  std::vector<float> bt((size_t)Kp * kKPad, 0.f), bsh((size_t)kKPad * Kp, 0.f), bsl((size_t)kKPad * Kp, 0.f);
  for (int c = 0; c < ncols; ++c) {
    const int col = 3 * vs * (c / 3) + (c % 3);
    for (int k = 0; k < kK; ++k) {
      const float x = m->h_Bm[(size_t)k * (3 * V) + col];
      bt[(size_t)c * kKPad + k] = x;
      const float h = tf32_hi(x);
      bsh[(size_t)k * Kp + c] = h;
      bsl[(size_t)k * Kp + c] = x - h;
    }
  }
  std::vector<int> ptr(kJ + 1, 0), vert, vq;
  std::vector<float> w;
  for (int j = 0; j < kJ; ++j) {
    for (int v = 0; v < V; v += vs) {
      float x = m->h_W[(size_t)v * kJ + j];
      if (x != 0.f) { vert.push_back(v); vq.push_back(v / vs); w.push_back(x); }
    }
    ptr[j + 1] = (int)vert.size();
  }
  // compact ELL skin weights of the sampled vertices (same packing as model_create)
  std::vector<uint8_t> sidx((size_t)Vs * m->KW, 0);
  std::vector<float> sw((size_t)Vs * m->KW, 0.f);
  for (int q = 0; q < Vs; ++q) {
    int nn = 0;
    for (int j = 0; j < kJ; ++j) {
      float x = m->h_W[(size_t)(q * vs) * kJ + j];
      if (x != 0.f) { sidx[(size_t)q * m->KW + nn] = (uint8_t)j; sw[(size_t)q * m->KW + nn] = x; ++nn; }
    }
  }
  CU_TRY(upload(&t->lbs_idx_s, sidx));
  CU_TRY(upload(&t->lbs_w_s, sw));
  CU_TRY(upload(&t->csc_q, vq));
  CU_TRY(upload(&t->BmT, bt));
  CU_TRY(upload(&t->Bs_hi, bsh));
  CU_TRY(upload(&t->Bs_lo, bsl));
  CU_TRY(upload(&t->csc_ptr, ptr));
  CU_TRY(upload(&t->csc_vert, vert));
  CU_TRY(upload(&t->csc_w, w));
  t->Vs = Vs;
  t->ncols = ncols;
  t->Kp = Kp;
  t->vs = vs;  // publish last
  return 0;
}

const VsTables* get_vs_tables(const SmplB200Model* m, int vs) {
  if (vs < 1) vs = 1;
  // always under the mutex (uncontended: ~20 ns): an unlocked fast path would race with build_vs_tables' publication
  std::lock_guard<std::mutex> lk(g_vs_mutex);
  SmplB200Model* mm = const_cast<SmplB200Model*>(m);
  for (int i = 0; i < kMaxVsCache; ++i) {
    if (mm->vst[i].vs == vs) return &mm->vst[i];
    if (mm->vst[i].vs == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaSetDevice(m->device);
      int rc = build_vs_tables(mm, vs, &mm->vst[i]);
      if (rc != 0) { free_vs_tables(&mm->vst[i]); }            // a half-built slot is released and stays unused
      cudaSetDevice(dev);
      return rc == 0 ? &mm->vst[i] : nullptr;
    }
  }
  set_error("vertex_sampling cache full (max %d distinct values per model)", kMaxVsCache);
  return nullptr;
}

static bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

}  // namespace smplb200

using namespace smplb200;

extern "C" {

int smpl_b200_abi_version(void) { return SMPL_B200_ABI_VERSION; }
const char* smpl_b200_last_error(void) { return g_err; }
uint64_t smpl_b200_launch_count(void) { return g_launches.load(); }

int smpl_b200_profile_enable(int on) {
  g_prof_on.store(on ? 1 : 0);
  return SMPL_B200_OK;
}

int smpl_b200_profile_collect(SmplB200KernelStat* out, int max_stats, int* num_stats) {
  if (!num_stats || (max_stats > 0 && !out) || max_stats < 0) { set_error("profile_collect: bad argument"); return SMPL_B200_ERR_BAD_ARG; }
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  double ms[KID_COUNT] = {0};
  long long cnt[KID_COUNT] = {0};
  for (ProfRecord& r : g_prof_records) {
    float t = 0.f;
    cudaError_t e = cudaEventSynchronize(r.b);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&t, r.a, r.b);
    if (e != cudaSuccess) { cudaGetLastError(); set_error("profile_collect: %s", cudaGetErrorString(e)); g_prof_records.clear(); return SMPL_B200_ERR_CUDA; }
    ms[r.kid] += t; cnt[r.kid] += 1;
    g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b);
  }
  g_prof_records.clear();
  int n = 0;
  for (int k = 0; k < KID_COUNT; ++k) {
    if (!cnt[k]) continue;
    if (n < max_stats) { out[n].name = kKernelNames[k]; out[n].launches = cnt[k]; out[n].total_ms = ms[k]; }
    ++n;
  }
  *num_stats = n;
  return SMPL_B200_OK;
}

int smpl_b200_model_create(const SmplB200HostModel* h, int device, SmplB200Model** out) {
  if (!h || !out) { set_error("model_create: null argument"); return SMPL_B200_ERR_BAD_ARG; }
  *out = nullptr;
  if (!h->v_template || !h->shapedirs || !h->posedirs || !h->J_regressor || !h->lbs_weights || !h->parents) {
    set_error("model_create: null constant array"); return SMPL_B200_ERR_BAD_ARG;
  }
  if (h->num_joints != kJ || h->num_betas != kBetas || h->num_pose_basis != kPoseBasis || h->num_verts < 2 ||
      h->num_verts > 65535) {
    set_error("model_create: unsupported sizes V=%d J=%d betas=%d pose_basis=%d (need J=24, 10, 207, 2<=V<=65535)",
              h->num_verts, h->num_joints, h->num_betas, h->num_pose_basis);
    return SMPL_B200_ERR_UNSUPPORTED;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    set_error("model_create: no usable CUDA device %d (found %d); this library has no CPU path", device, ndev);
    return SMPL_B200_ERR_NO_DEVICE;
  }
  int prev = 0;
  cudaGetDevice(&prev);
  CU_TRY(cudaSetDevice(device));
  const int V = h->num_verts;
  SmplB200Model* m = new SmplB200Model();
  // every error return below releases the half-built model (device buffers included) and restores the caller's device
  struct Guard {
    SmplB200Model* m; int prev;
    ~Guard() { if (m) smpl_b200_model_destroy(m); cudaSetDevice(prev); }
  } guard{m, prev};
  m->device = device;
  m->V = V;
  m->LD = SMPL_B200_VPOSED_LD(V);
  m->R = h->joint_regressor ? h->num_reg_joints : 0;
  // kinematic tree
  for (int j = 0; j < kJ; ++j) {
    int p = (j == 0) ? -1 : h->parents[j];
    if (j > 0 && (p < 0 || p >= j)) {
      set_error("model_create: parents[%d]=%d is not a topologically ordered tree", j, p);
      return SMPL_B200_ERR_BAD_ARG;
    }
    m->tree.parent[j] = p;
    m->tree.depth[j] = (j == 0) ? 0 : m->tree.depth[p] + 1;
    for (int c = 0; c < 4; ++c) m->tree.child[j][c] = -1;
  }
  m->tree.max_depth = 0;
  for (int j = 1; j < kJ; ++j) {
    m->tree.max_depth = std::max(m->tree.max_depth, m->tree.depth[j]);
    int p = m->tree.parent[j], c = 0;
    while (c < 4 && m->tree.child[p][c] >= 0) ++c;
    if (c == 4) {
      set_error("model_create: joint %d has more than 4 children", p);
      return SMPL_B200_ERR_UNSUPPORTED;
    }
    m->tree.child[p][c] = j;
  }
  // blend matrix [kKPad][LD], template [LD]
  const size_t C = (size_t)3 * V;
  m->h_Bm = (float*)malloc(sizeof(float) * kK * C);
  m->h_W = (float*)malloc(sizeof(float) * (size_t)V * kJ);
  memcpy(m->h_Bm, h->shapedirs, sizeof(float) * kBetas * C);
  memcpy(m->h_Bm + kBetas * C, h->posedirs, sizeof(float) * kPoseBasis * C);
  memcpy(m->h_W, h->lbs_weights, sizeof(float) * (size_t)V * kJ);
  {
    std::vector<float> bm((size_t)kKPad * m->LD, 0.f), vt(m->LD, 0.f);
    for (int k = 0; k < kK; ++k) memcpy(&bm[(size_t)k * m->LD], m->h_Bm + (size_t)k * C, sizeof(float) * C);
    memcpy(vt.data(), h->v_template, sizeof(float) * C);
    CU_TRY(upload(&m->Bm, bm));
    CU_TRY(upload(&m->vt_pad, vt));
    // K-major copies for the tensor-core forward: BT[c][k] = Bm[k][c], split exactly into TF32 hi + remainder
    std::vector<float> bth((size_t)m->LD * kKPad, 0.f), btl((size_t)m->LD * kKPad, 0.f);
    for (int k = 0; k < kK; ++k)
      for (size_t c = 0; c < C; ++c) {
        const float x = m->h_Bm[(size_t)k * C + c];
        const float h = tf32_hi(x);
        bth[c * kKPad + k] = h;
        btl[c * kKPad + k] = x - h;
      }
    CU_TRY(upload(&m->BT_hi, bth));
    CU_TRY(upload(&m->BT_lo, btl));
    // fp16 split of the same matrix for the forward (twice the tensor rate of TF32): B * 2^e = hi + lo with max |B| 2^e in
    // [2^13, 2^14), so hi carries 11 bits and lo the next 11 (its own quantisation, 2^-24 / 2^e absolute, is far below
    // fp32's); products of two fp16 values are exact in the fp32 accumulator.  (A degenerate model -- all-zero or
    // non-finite blend shapes -- has no such scale and keeps the 3xTF32 tables for the forward too.)
    float bmax = 0.f;
    for (size_t i = 0; i < (size_t)kK * C; ++i) bmax = fmaxf(bmax, fabsf(m->h_Bm[i]));
    if (bmax > 0.f && std::isfinite(bmax)) {
      int ex = 0;
      frexpf(bmax, &ex);                                   // bmax = f * 2^ex, f in [0.5, 1)
      m->bt16_scale = ldexpf(1.0f, 14 - ex);
      std::vector<uint16_t> b16h((size_t)m->LD * kKPad, 0), b16l((size_t)m->LD * kKPad, 0);
      for (int k = 0; k < kK; ++k)
        for (size_t c = 0; c < C; ++c) {
          const float x = m->h_Bm[(size_t)k * C + c] * m->bt16_scale;
          const __half h = __float2half_rn(x);
          const __half l = __float2half_rn(x - __half2float(h));
          b16h[c * kKPad + k] = __half_as_ushort(h);
          b16l[c * kKPad + k] = __half_as_ushort(l);
        }
      CU_TRY(upload(&m->BT16_hi, b16h));
      CU_TRY(upload(&m->BT16_lo, b16l));
    }
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    m->num_sms = prop.multiProcessorCount;
    if (prop.major != 10) {
      set_error("model_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
      return SMPL_B200_ERR_NO_DEVICE;
    }
  }
  // folded joint regression: J = Jt + Jd * beta   (exact algebra of batch_smpl.py:106-115, evaluated in fp64)
  {
    std::vector<double> jt(kJ * 3, 0.0), jd((size_t)kJ * 3 * kBetas, 0.0);
    for (int v = 0; v < V; ++v)
      for (int j = 0; j < kJ; ++j) {
        const double r = h->J_regressor[(size_t)v * kJ + j];
        if (r == 0.0) continue;
        for (int c = 0; c < 3; ++c) {
          jt[j * 3 + c] += r * (double)h->v_template[v * 3 + c];
          for (int k = 0; k < kBetas; ++k)
            jd[((size_t)j * 3 + c) * kBetas + k] += r * (double)h->shapedirs[(size_t)k * C + v * 3 + c];
        }
      }
    std::vector<float> jtf(jt.begin(), jt.end()), jdf(jd.begin(), jd.end());
    CU_TRY(upload(&m->Jt, jtf));
    CU_TRY(upload(&m->Jd, jdf));
  }
  // LBS weights -> ELL (idx,w) of width KW = max non-zeros per row
  {
    int kw = 1;
    for (int v = 0; v < V; ++v) {
      int n = 0;
      for (int j = 0; j < kJ; ++j) n += h->lbs_weights[(size_t)v * kJ + j] != 0.f;
      kw = std::max(kw, n);
    }
    kw = (kw <= 4) ? 4 : (kw <= 8 ? 8 : kJ);
    m->KW = kw;
    std::vector<uint8_t> idx((size_t)V * kw, 0);
    std::vector<float> w((size_t)V * kw, 0.f);
    for (int v = 0; v < V; ++v) {
      int n = 0;
      for (int j = 0; j < kJ; ++j) {
        float x = h->lbs_weights[(size_t)v * kJ + j];
        if (x != 0.f) { idx[(size_t)v * kw + n] = (uint8_t)j; w[(size_t)v * kw + n] = x; ++n; }
      }
    }
    CU_TRY(upload(&m->lbs_idx, idx));
    CU_TRY(upload(&m->lbs_w, w));
  }
  // sparse keypoint regressor (CSR by keypoint)
  {
    std::vector<int> ptr(m->R + 1, 0), vert;
    std::vector<float> w;
    for (int r = 0; r < m->R; ++r) {
      for (int v = 0; v < V; ++v) {
        float x = h->joint_regressor[(size_t)v * h->num_reg_joints + r];
        if (x != 0.f) { vert.push_back(v); w.push_back(x); }
      }
      ptr[r + 1] = (int)vert.size();
    }
    CU_TRY(upload(&m->jr_ptr, ptr));
    CU_TRY(upload(&m->jr_vert, vert));
    CU_TRY(upload(&m->jr_w, w));
  }
  // the reference's three sampling modes are built eagerly so hot calls never allocate
  const int eager[3] = {1, 2, 5};
  for (int i = 0; i < 3; ++i) {
    int rc = build_vs_tables(m, eager[i], &m->vst[i]);
    if (rc != 0) return rc;
  }
  CU_TRY(cudaDeviceSynchronize());
  guard.m = nullptr;                 // success: the caller owns the model now (the guard still restores the device)
  *out = m;
  return SMPL_B200_OK;
}

void smpl_b200_model_destroy(SmplB200Model* m) {
  if (!m) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(m->device);
  cudaFree(m->vt_pad); cudaFree(m->Bm); cudaFree(m->BT_hi); cudaFree(m->BT_lo); cudaFree(m->BT16_hi); cudaFree(m->BT16_lo); cudaFree(m->Jt); cudaFree(m->Jd); cudaFree(m->lbs_idx); cudaFree(m->lbs_w);
  cudaFree(m->jr_ptr); cudaFree(m->jr_vert); cudaFree(m->jr_w);
  for (int i = 0; i < kMaxVsCache; ++i) free_vs_tables(&m->vst[i]);
  free(m->h_Bm); free(m->h_W);
  cudaSetDevice(prev);
  delete m;
}

int smpl_b200_model_num_verts(const SmplB200Model* m) { return m ? m->V : 0; }
int smpl_b200_model_lbs_width(const SmplB200Model* m) { return m ? m->KW : 0; }

int smpl_b200_parts_create(int device, int num_parts, const int32_t* part_ptr, const int32_t* part_idx,
                           int num_sampled_verts, SmplB200Parts** out) {
  if (!part_ptr || !out || num_parts < 1 || num_sampled_verts < 1) {
    set_error("parts_create: bad argument"); return SMPL_B200_ERR_BAD_ARG;
  }
  *out = nullptr;
  if (num_parts > 31) { set_error("parts_create: at most 31 parts supported (got %d)", num_parts); return SMPL_B200_ERR_UNSUPPORTED; }
  const int E = part_ptr[num_parts];
  if (part_ptr[0] != 0 || E < 0 || (E > 0 && !part_idx)) { set_error("parts_create: bad CSR"); return SMPL_B200_ERR_BAD_ARG; }
  std::vector<int> ptr(part_ptr, part_ptr + num_parts + 1), idx(part_idx, part_idx + E);
  std::vector<uint8_t> po(E);
  int max_part = 0;
  for (int k = 0; k < num_parts; ++k) {
    if (ptr[k + 1] < ptr[k]) { set_error("parts_create: part_ptr not monotone"); return SMPL_B200_ERR_BAD_ARG; }
    max_part = std::max(max_part, ptr[k + 1] - ptr[k]);
    for (int e = ptr[k]; e < ptr[k + 1]; ++e) {
      if (idx[e] < 0 || idx[e] >= num_sampled_verts) {
        set_error("parts_create: vertex index %d out of range [0,%d)", idx[e], num_sampled_verts);
        return SMPL_B200_ERR_BAD_ARG;
      }
      po[e] = (uint8_t)k;
    }
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    set_error("parts_create: no usable CUDA device %d", device); return SMPL_B200_ERR_NO_DEVICE;
  }
  int prev = 0;
  cudaGetDevice(&prev);
  CU_TRY(cudaSetDevice(device));
  SmplB200Parts* p = new SmplB200Parts();
  struct Guard {
    SmplB200Parts* p; int prev;
    ~Guard() { if (p) smpl_b200_parts_destroy(p); cudaSetDevice(prev); }
  } guard{p, prev};
  p->device = device; p->P = num_parts; p->E = E; p->Vs = num_sampled_verts; p->max_part = max_part;
  CU_TRY(upload(&p->ptr, ptr));
  CU_TRY(upload(&p->idx, idx));
  CU_TRY(upload(&p->part_of, po));
  std::vector<int> ob(num_parts + 1, 0);
  for (int k = 0; k < num_parts; ++k) ob[k + 1] = ob[k] + std::max(ptr[k + 1] - ptr[k] - 32, 0);
  p->ovf = ob[num_parts];
  p->keep_words = 32;     // a row of the seg forward's survivor table: 32 first words + the parts' further words
  for (int k = 0; k < num_parts; ++k) p->keep_words += std::max((ptr[k + 1] - ptr[k] + 31) / 32 - 1, 0);
  CU_TRY(upload(&p->obase, ob));
  CU_TRY(cudaDeviceSynchronize());
  guard.p = nullptr;
  *out = p;
  return SMPL_B200_OK;
}

void smpl_b200_parts_destroy(SmplB200Parts* p) {
  if (!p) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(p->device);
  cudaFree(p->ptr); cudaFree(p->idx); cudaFree(p->part_of); cudaFree(p->obase);
  cudaSetDevice(prev);
  delete p;
}

// ---- workspace layout ------------------------------------------------------------------------------------
static size_t ru(size_t x, size_t a) { return (x + a - 1) / a * a; }
struct DecodeWs {
  float *X, *Xlo, *A, *Jtr, *gA, *gX, *gcam, *gvp, *gvplo;
  size_t gvp_ld;
  int cam_chunks;
  size_t bytes;
};
static DecodeWs decode_ws(const SmplB200Model* m, int N, bool bwd, bool full_grad, int vs, void* base) {
  DecodeWs w{};
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t nfloat) { float* r = (float*)(p + off); off += ru(nfloat * sizeof(float), 256); return r; };
  w.X = take((size_t)N * kKPad);
  w.Xlo = take((size_t)N * kKPad);
  w.A = take((size_t)N * kJ * 12);
  w.Jtr = take((size_t)N * kJ * 3);
  if (bwd) {
    w.gA = take((size_t)N * kJ * 12);
    w.gX = take((size_t)N * kKPad * (N >= kDenseBatch ? kMaxBlendBwdSplit : 1));   // K-split planes of the tensor-core product
    const int Vs = (m->V + vs - 1) / vs;
    w.cam_chunks = lbs_bwd_cam_chunks(Vs);
    w.gcam = take((size_t)N * 4 * w.cam_chunks);
    w.gvp_ld = full_grad ? (size_t)m->LD : ru((size_t)Vs * 3, 32);   // >= VsTables::Kp, rows stay 16-byte aligned
    w.gvp = take((size_t)N * w.gvp_ld);
    w.gvplo = take((size_t)N * w.gvp_ld);
  }
  w.bytes = off;
  return w;
}

size_t smpl_b200_workspace_bytes(const SmplB200Model* m, int op, int N, int img_wh, int vertex_sampling) {
  if (!m || N < 0) return 0;
  const int vs = vertex_sampling < 1 ? 1 : vertex_sampling;
  (void)img_wh;
  switch (op) {
    case SMPL_B200_OP_DECODE_FWD: return decode_ws(m, N, false, false, 1, nullptr).bytes;
    case SMPL_B200_OP_DECODE_BWD: return decode_ws(m, N, true, vs == 1, vs, nullptr).bytes;
    case SMPL_B200_OP_SILHOUETTE_FWD:      // the arg-min map the forward hands to the backward: 2 bytes per pixel
    case SMPL_B200_OP_SILHOUETTE_BWD: return ru((size_t)N * img_wh * img_wh * 2, 256);
    case SMPL_B200_OP_FULL_FWD:      // decode scratch + the transient v_posed
      return decode_ws(m, N, false, false, 1, nullptr).bytes + ru((size_t)N * m->LD * sizeof(float), 256);
    case SMPL_B200_OP_FULL_BWD:      // decode scratch (sampled gradient) + g_projects
      return decode_ws(m, N, true, false, vs, nullptr).bytes + ru((size_t)N * ((m->V + vs - 1) / vs) * 3 * sizeof(float), 256);
    default: return 0;
  }
}

#define CHECK_LAUNCH(expr)                                                           \
  do {                                                                               \
    cudaError_t e__ = (expr);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      set_error("%s: %s", #expr, cudaGetErrorString(e__));                           \
      return SMPL_B200_ERR_CUDA;                                                     \
    }                                                                                \
  } while (0)

// A_keep (nullable, [N][24][12]): the bone transforms, kept for a backward that then skips their recomputation
static int decode_fwd_impl(const SmplB200Model* m, const float* params, int N, float* verts, float* joints24,
                           float* joints_reg, int num_reg_joints_used, float* v_posed_save, float* v_posed_sampled,
                           float* projects, int vertex_sampling, void* workspace, size_t workspace_bytes, void* stream,
                           float* A_keep) {
  if (N == 0) return SMPL_B200_OK;   // empty batch: nothing to launch
  if (!m || !params || N < 0 || (!verts && !projects)) { set_error("decode_fwd: null/invalid argument"); return SMPL_B200_ERR_BAD_ARG; }
  const int vs = vertex_sampling < 1 ? 1 : vertex_sampling;
  if (!aligned(params, 4) || (verts && !aligned(verts, 8)) || (v_posed_save && !aligned(v_posed_save, 16)) || !aligned(workspace, 16)) {
    set_error("decode_fwd: misaligned buffer (verts 8B, v_posed_save/workspace 16B)"); return SMPL_B200_ERR_BAD_ARG;
  }
  if (joints_reg && (num_reg_joints_used < 1 || num_reg_joints_used > m->R)) {
    set_error("decode_fwd: joints_reg requested with %d keypoints, model has %d", num_reg_joints_used, m->R);
    return SMPL_B200_ERR_BAD_ARG;
  }
  if (joints_reg && !verts) { set_error("decode_fwd: joints_reg needs verts"); return SMPL_B200_ERR_BAD_ARG; }
  if (v_posed_sampled && !projects) { set_error("decode_fwd: v_posed_sampled needs projects"); return SMPL_B200_ERR_BAD_ARG; }
  // v_posed must live somewhere: the caller's save buffer, else the tail of the workspace
  DecodeWs w = decode_ws(m, N, false, false, 1, workspace);
  size_t need = w.bytes + (v_posed_save ? 0 : ru((size_t)N * m->LD * sizeof(float), 256));
  if (!workspace || workspace_bytes < need) {
    set_error("decode_fwd: workspace too small (%zu < %zu bytes)", workspace_bytes, need); return SMPL_B200_ERR_WORKSPACE;
  }
  float* vp = v_posed_save ? v_posed_save : (float*)((char*)workspace + w.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const bool dense = N >= kDenseBatch;                                  // tensor cores only where the batch makes the GEMM dense
  float* Abuf = A_keep ? A_keep : w.A;
  CHECK_LAUNCH(launch_pose_fwd(m, params, N, w.X, dense ? w.Xlo : nullptr, Abuf, joints24 ? joints24 : w.Jtr, st));
  if (dense) CHECK_LAUNCH(launch_blend_fwd_tc(m, w.X, w.Xlo, N, vp, st));
  else CHECK_LAUNCH(launch_blend_fwd(m, w.X, N, vp, st));
  const int Vs = (m->V + vs - 1) / vs;
  CHECK_LAUNCH(launch_lbs_fwd(m, vp, Abuf, params, N, verts, projects, vs, v_posed_sampled, SMPL_B200_VPS_LD(Vs), st));
  if (joints_reg) CHECK_LAUNCH(launch_joints_reg_fwd(m, verts, N, num_reg_joints_used, joints_reg, st));
  return SMPL_B200_OK;
}

int smpl_b200_decode_fwd(const SmplB200Model* m, const float* params, int N, float* verts, float* joints24,
                         float* joints_reg, int num_reg_joints_used, float* v_posed_save, float* v_posed_sampled,
                         float* projects, int vertex_sampling, void* workspace, size_t workspace_bytes, void* stream) {
  return decode_fwd_impl(m, params, N, verts, joints24, joints_reg, num_reg_joints_used, v_posed_save, v_posed_sampled,
                         projects, vertex_sampling, workspace, workspace_bytes, stream, nullptr);
}

// A_saved (nullable): decode_fwd_impl's A_keep; without it the transforms are recomputed from params
static int decode_bwd_impl(const SmplB200Model* m, const float* params, int N, const float* v_posed_save,
                           const float* v_posed_sampled, const float* g_verts, const float* g_projects,
                           int vertex_sampling, const float* g_joints24, float* g_params, void* workspace,
                           size_t workspace_bytes, void* stream, const float* A_saved) {
  if (N == 0) return SMPL_B200_OK;   // empty batch: nothing to launch
  if (!m || !params || !g_params || N < 0) { set_error("decode_bwd: null/invalid argument"); return SMPL_B200_ERR_BAD_ARG; }
  const int vs_in = vertex_sampling < 1 ? 1 : vertex_sampling;
  const bool full = (g_verts != nullptr) || !g_projects;     // dense vertex gradient (or none at all): every vertex
  if (!v_posed_save && (full || !v_posed_sampled)) {
    set_error("decode_bwd: v_posed_save is required unless the gradient arrives through g_projects only and "
              "v_posed_sampled is given");
    return SMPL_B200_ERR_BAD_ARG;
  }
  const int vs_t = full ? 1 : vs_in;
  const VsTables* t = get_vs_tables(m, vs_t);
  if (!t) return SMPL_B200_ERR_CUDA;
  DecodeWs w = decode_ws(m, N, true, full, vs_t, workspace);
  if (!workspace || workspace_bytes < w.bytes || !aligned(workspace, 16)) {
    set_error("decode_bwd: workspace too small or misaligned (%zu < %zu bytes)", workspace_bytes, w.bytes);
    return SMPL_B200_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const bool dense = N >= kDenseBatch;
  if (!A_saved) CHECK_LAUNCH(launch_pose_fwd(m, params, N, w.X, nullptr, w.A, w.Jtr, st));   // recompute A instead of saving it
  const float* Abuf = A_saved ? A_saved : w.A;
  int cam_chunks = w.cam_chunks;
  CHECK_LAUNCH(launch_lbs_bwd(m, t, vs_in, v_posed_save, Abuf, params, g_verts, g_projects, N, w.gvp,
                              dense ? w.gvplo : nullptr, w.gvp_ld, w.gA, w.gcam, &cam_chunks,
                              full ? nullptr : v_posed_sampled, SMPL_B200_VPS_LD(t->Vs), st));
  int gx_planes = 1;
  if (dense) CHECK_LAUNCH(launch_blend_bwd_tc(m, t, w.gvp, w.gvplo, w.gvp_ld, N, w.gX, &gx_planes, st));
  else CHECK_LAUNCH(launch_blend_bwd(m, t, w.gvp, w.gvp_ld, N, w.gX, st));
  CHECK_LAUNCH(launch_pose_bwd(m, params, w.gA, w.gX, gx_planes, g_joints24, w.gcam, cam_chunks, N, g_params, st));
  return SMPL_B200_OK;
}

int smpl_b200_decode_bwd(const SmplB200Model* m, const float* params, int N, const float* v_posed_save,
                         const float* v_posed_sampled, const float* g_verts, const float* g_projects,
                         int vertex_sampling, const float* g_joints24, float* g_params, void* workspace,
                         size_t workspace_bytes, void* stream) {
  return decode_bwd_impl(m, params, N, v_posed_save, v_posed_sampled, g_verts, g_projects, vertex_sampling, g_joints24,
                         g_params, workspace, workspace_bytes, stream, nullptr);
}

int smpl_b200_project_fwd(const float* verts, const float* params, int N, int V, int vertex_sampling, float* projects,
                          void* stream) {
  if (N == 0) return SMPL_B200_OK;   // empty batch: nothing to launch
  if (!verts || !params || !projects || N < 0 || V < 1) { set_error("project_fwd: null/invalid argument"); return SMPL_B200_ERR_BAD_ARG; }
  CHECK_LAUNCH(launch_project_fwd(verts, params, N, V, vertex_sampling < 1 ? 1 : vertex_sampling, projects, (cudaStream_t)stream));
  return SMPL_B200_OK;
}

int smpl_b200_project_bwd(const float* verts, const float* params, const float* g_projects, int N, int V,
                          int vertex_sampling, float* g_verts, float* g_params, void* stream) {
  if (N == 0) return SMPL_B200_OK;   // empty batch: nothing to launch
  if (!verts || !params || !g_projects || !g_verts || !g_params || N < 0 || V < 1) {
    set_error("project_bwd: null/invalid argument"); return SMPL_B200_ERR_BAD_ARG;
  }
  CHECK_LAUNCH(launch_project_bwd(verts, params, g_projects, N, V, vertex_sampling < 1 ? 1 : vertex_sampling, g_verts, g_params,
                                  (cudaStream_t)stream));
  return SMPL_B200_OK;
}

int smpl_b200_mask_fwd(const float* projects, int N, int Vs, float* mask, void* stream) {
  if (N == 0) return SMPL_B200_OK;   // empty batch: nothing to launch
  if (!projects || !mask || N < 0 || Vs < 1) { set_error("mask_fwd: null/invalid argument"); return SMPL_B200_ERR_BAD_ARG; }
  CHECK_LAUNCH(launch_mask_fwd(projects, N, Vs, mask, (cudaStream_t)stream));
  return SMPL_B200_OK;
}

static int seg_check(const SmplB200Parts* p, const void* a, const void* b, int N, int Vs, int wh, const char* who) {
  if (!p || !a || !b || N < 0) { set_error("%s: null/invalid argument", who); return SMPL_B200_ERR_BAD_ARG; }
  if (Vs != p->Vs) { set_error("%s: Vs=%d does not match the part table (%d)", who, Vs, p->Vs); return SMPL_B200_ERR_BAD_ARG; }
  if (wh < 1 || wh > 128) { set_error("%s: img_wh=%d unsupported (1..128)", who, wh); return SMPL_B200_ERR_UNSUPPORTED; }
  if (p->E > 65534) { set_error("%s: part table too large (%d entries, max 65534)", who, p->E); return SMPL_B200_ERR_UNSUPPORTED; }
  return 0;
}

size_t smpl_b200_seg_saved_bytes(int N, int img_wh) {
  if (N < 0 || img_wh < 0) return 0;
  return seg_saved_bytes(N, img_wh);
}

int smpl_b200_seg_fwd(const SmplB200Parts* parts, const float* projects, const float* mask, int N, int Vs, int img_wh,
                      float* seg, void* saved, void* stream) {
  if (N == 0) return SMPL_B200_OK;   // empty batch: nothing to launch
  int rc = seg_check(parts, projects, mask, N, Vs, img_wh, "seg_fwd");
  if (rc) return rc;
  if (!seg && !saved) { set_error("seg_fwd: both outputs are null"); return SMPL_B200_ERR_BAD_ARG; }
  if ((seg && !aligned(seg, 16)) || (saved && !aligned(saved, 16))) {
    set_error("seg_fwd: seg/saved must be 16-byte aligned"); return SMPL_B200_ERR_BAD_ARG;
  }
  cudaError_t e = launch_seg_fwd(parts, projects, mask, N, Vs, img_wh, seg, (unsigned char*)saved, (cudaStream_t)stream);
  if (e == cudaErrorInvalidConfiguration) {
    set_error("seg_fwd: part table (%d entries) and img_wh=%d need more than 227 KB of shared memory", parts->E, img_wh);
    return SMPL_B200_ERR_UNSUPPORTED;
  }
  CHECK_LAUNCH(e);
  return SMPL_B200_OK;
}

int smpl_b200_seg_bwd(const SmplB200Parts* parts, const float* projects, const float* mask, const float* g_seg,
                      const void* saved, int N, int Vs, int img_wh, float* g_projects, void* stream) {
  if (N == 0) return SMPL_B200_OK;   // empty batch: nothing to launch
  int rc = seg_check(parts, projects, mask, N, Vs, img_wh, "seg_bwd");
  if (rc) return rc;
  if (!g_seg || !saved || !g_projects) {
    set_error("seg_bwd: null g_seg / saved / g_projects (run seg_fwd with a `saved` buffer first)");
    return SMPL_B200_ERR_BAD_ARG;
  }
  cudaError_t e = launch_seg_bwd(parts, projects, mask, g_seg, (const unsigned char*)saved, N, Vs, img_wh, g_projects,
                                 (cudaStream_t)stream);
  if (e == cudaErrorInvalidConfiguration) {
    set_error("seg_bwd: part table (%d entries), Vs=%d and img_wh=%d need more than 227 KB of shared memory", parts->E, Vs, img_wh);
    return SMPL_B200_ERR_UNSUPPORTED;
  }
  CHECK_LAUNCH(e);
  return SMPL_B200_OK;
}

// ---- projects_to_seg + Reshape -> softmax -> categorical focal loss, fused (integer labels) ----------------------------
size_t smpl_b200_seg_loss_state_bytes(int N, int img_wh) {
  if (N < 0 || img_wh < 0) return 0;
  return seg_saved_bytes(N, img_wh) + (size_t)N * img_wh * img_wh * 16;
}

int smpl_b200_seg_loss_fwd(const SmplB200Parts* parts, const float* projects, const float* mask, int N, int Vs,
                           int img_wh, const uint8_t* labels, float gamma, const float* class_weights, float* seg,
                           float* loss, void* state, void* stream) {
  if (N == 0) return SMPL_B200_OK;
  int rc = seg_check(parts, projects, mask, N, Vs, img_wh, "seg_loss_fwd");
  if (rc) return rc;
  if (!labels || !loss || !state) { set_error("seg_loss_fwd: null labels / loss / state"); return SMPL_B200_ERR_BAD_ARG; }
  if ((seg && !aligned(seg, 16)) || !aligned(state, 16) || !aligned(loss, 4)) {
    set_error("seg_loss_fwd: seg/state must be 16-byte aligned"); return SMPL_B200_ERR_BAD_ARG;
  }
  unsigned char* saved = (unsigned char*)state;
  void* aux = saved + seg_saved_bytes(N, img_wh);
  cudaError_t e = launch_seg_loss_fwd(parts, projects, mask, N, Vs, img_wh, labels, gamma, class_weights, seg, loss, saved,
                                      aux, (cudaStream_t)stream);
  if (e == cudaErrorInvalidConfiguration) {
    set_error("seg_loss_fwd: part table (%d entries) and img_wh=%d need more than 227 KB of shared memory", parts->E, img_wh);
    return SMPL_B200_ERR_UNSUPPORTED;
  }
  CHECK_LAUNCH(e);
  return SMPL_B200_OK;
}

int smpl_b200_seg_loss_bwd(const SmplB200Parts* parts, const float* projects, const float* mask, const float* g_loss,
                           const void* state, int N, int Vs, int img_wh, float* g_projects, void* stream) {
  if (N == 0) return SMPL_B200_OK;
  int rc = seg_check(parts, projects, mask, N, Vs, img_wh, "seg_loss_bwd");
  if (rc) return rc;
  if (!g_loss || !state || !g_projects) { set_error("seg_loss_bwd: null g_loss / state / g_projects"); return SMPL_B200_ERR_BAD_ARG; }
  if (!aligned(state, 16)) { set_error("seg_loss_bwd: state must be 16-byte aligned"); return SMPL_B200_ERR_BAD_ARG; }
  const unsigned char* saved = (const unsigned char*)state;
  const void* aux = saved + seg_saved_bytes(N, img_wh);
  cudaError_t e = launch_seg_loss_bwd(parts, projects, mask, g_loss, saved, aux, N, Vs, img_wh, g_projects,
                                      (cudaStream_t)stream);
  if (e == cudaErrorInvalidConfiguration) {
    set_error("seg_loss_bwd: part table (%d entries), Vs=%d and img_wh=%d need more than 227 KB of shared memory", parts->E, Vs, img_wh);
    return SMPL_B200_ERR_UNSUPPORTED;
  }
  CHECK_LAUNCH(e);
  return SMPL_B200_OK;
}

// ---- the regression module ahead of the decoder: Dense layers (model.py:63-105) -------------------------------------
static int current_sms() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

size_t smpl_b200_dense_workspace_bytes(int M, int in, int out) {
  if (M < 0 || in < 1 || out < 1) return 0;
  size_t a = dense_gemm_ws(M, out, in);        // forward:  [M][in]  x [out][in]^T
  size_t b = dense_gemm_ws(M, in, out);        // gX:       [M][out] x [in][out]^T
  size_t c = dense_gemm_ws(in, out, M);        // gW:       [in][M]  x [out][M]^T
  return std::max(a, std::max(b, c));
}

static int dense_check(const char* who, const void* X, const void* W, int M, int in, int out, const void* ws, size_t ws_bytes) {
  if (!X || !W || M < 0 || in < 1 || out < 1) { set_error("%s: null/invalid argument", who); return SMPL_B200_ERR_BAD_ARG; }
  if (!ws || ws_bytes < smpl_b200_dense_workspace_bytes(M, in, out) || !aligned(ws, 256)) {
    set_error("%s: workspace too small or not 256-byte aligned (%zu < %zu bytes)", who, ws_bytes,
              smpl_b200_dense_workspace_bytes(M, in, out));
    return SMPL_B200_ERR_WORKSPACE;
  }
  return 0;
}

int smpl_b200_dense_fwd(const float* X, int ldx, const float* W, const float* bias, int M, int in, int out, int relu, float* Y,
                        int ldy, void* workspace, size_t workspace_bytes, void* stream) {
  if (M == 0) return SMPL_B200_OK;
  int rc = dense_check("dense_fwd", X, W, M, in, out, workspace, workspace_bytes);
  if (rc) return rc;
  if (!Y || ldx < in || ldy < out) { set_error("dense_fwd: null output or row stride shorter than the row"); return SMPL_B200_ERR_BAD_ARG; }
  CHECK_LAUNCH(launch_dense_fwd(X, ldx, W, bias, M, in, out, relu != 0, Y, ldy, workspace, current_sms(), (cudaStream_t)stream));
  return SMPL_B200_OK;
}

int smpl_b200_dense_bwd(const float* X, int ldx, const float* W, const float* Y, int ldy, const float* gY, int ldg, int M,
                        int in, int out, int relu, float* gX, int ldgx, float* gW, float* gb, int accumulate, void* workspace,
                        size_t workspace_bytes, void* stream) {
  if (M == 0) return SMPL_B200_OK;
  int rc = dense_check("dense_bwd", X, W, M, in, out, workspace, workspace_bytes);
  if (rc) return rc;
  if (!gY || ldg < out || (relu && (!Y || ldy < out)) || (gX && ldgx < in)) {
    set_error("dense_bwd: null gY / Y (needed for the ReLU gate) or row stride shorter than the row"); return SMPL_B200_ERR_BAD_ARG;
  }
  CHECK_LAUNCH(launch_dense_bwd(X, ldx, W, Y, ldy, gY, ldg, M, in, out, relu != 0, gX, ldgx, gW, gb, accumulate != 0, workspace,
                                current_sms(), (cudaStream_t)stream));
  return SMPL_B200_OK;
}

int smpl_b200_axpy_cols(const float* a, int lda, const float* d, int ldd, float scale, int rows, int cols, float* out, int ldo,
                        void* stream) {
  if (rows == 0 || cols == 0) return SMPL_B200_OK;
  if (!out || rows < 0 || cols < 0 || (!a && !d)) { set_error("axpy_cols: null/invalid argument"); return SMPL_B200_ERR_BAD_ARG; }
  CHECK_LAUNCH(launch_axpy_cols(a, lda, d, ldd, scale, rows, cols, out, ldo, (cudaStream_t)stream));
  return SMPL_B200_OK;
}

// ---- the whole path in one call (model.py:108-118) ---------------------------------------------------------------
static size_t full_vps_bytes(const SmplB200Model* m, int N, int vs) {
  const int Vs = (m->V + vs - 1) / vs;
  return ru((size_t)N * SMPL_B200_VPS_LD(Vs) * sizeof(float), 256);
}

static size_t full_A_bytes(int N) { return ru((size_t)N * kJ * 12 * sizeof(float), 256); }

// state = [compact sampled v_posed][bone transforms][seg arg-min bytes]
size_t smpl_b200_full_state_bytes(const SmplB200Model* m, int N, int img_wh, int vertex_sampling) {
  if (!m || N < 0 || img_wh < 0) return 0;
  const int vs = vertex_sampling < 1 ? 1 : vertex_sampling;
  return full_vps_bytes(m, N, vs) + full_A_bytes(N) + seg_saved_bytes(N, img_wh);
}

int smpl_b200_full_fwd(const SmplB200Model* m, const SmplB200Parts* parts, const float* params, int N, int img_wh,
                       int vertex_sampling, float* verts, float* joints24, float* projects, float* mask, float* seg,
                       void* state, void* workspace, size_t workspace_bytes, void* stream) {
  if (N == 0) return SMPL_B200_OK;
  if (!m || !parts || !params || !projects || !mask || !seg || N < 0) { set_error("full_fwd: null/invalid argument"); return SMPL_B200_ERR_BAD_ARG; }
  const int vs = vertex_sampling < 1 ? 1 : vertex_sampling;
  const int Vs = (m->V + vs - 1) / vs;
  int rc = seg_check(parts, projects, mask, N, Vs, img_wh, "full_fwd");
  if (rc) return rc;
  if (!aligned(seg, 16) || (state && !aligned(state, 16)) || !aligned(workspace, 16) || (verts && !aligned(verts, 8))) {
    set_error("full_fwd: misaligned buffer (seg/state/workspace 16B, verts 8B)"); return SMPL_B200_ERR_BAD_ARG;
  }
  const size_t need = smpl_b200_workspace_bytes(m, SMPL_B200_OP_FULL_FWD, N, img_wh, vs);
  if (!workspace || workspace_bytes < need) {
    set_error("full_fwd: workspace too small (%zu < %zu bytes)", workspace_bytes, need); return SMPL_B200_ERR_WORKSPACE;
  }
  float* vps = state ? (float*)state : nullptr;
  float* A_keep = state ? (float*)((char*)state + full_vps_bytes(m, N, vs)) : nullptr;
  unsigned char* saved = state ? (unsigned char*)state + full_vps_bytes(m, N, vs) + full_A_bytes(N) : nullptr;
  rc = decode_fwd_impl(m, params, N, verts, joints24, nullptr, 0, nullptr, vps, projects, vs, workspace, workspace_bytes,
                       stream, A_keep);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  CHECK_LAUNCH(launch_mask_fwd(projects, N, Vs, mask, st));
  cudaError_t e = launch_seg_fwd(parts, projects, mask, N, Vs, img_wh, seg, saved, st);
  if (e == cudaErrorInvalidConfiguration) {
    set_error("full_fwd: part table (%d entries) and img_wh=%d need more than 227 KB of shared memory", parts->E, img_wh);
    return SMPL_B200_ERR_UNSUPPORTED;
  }
  CHECK_LAUNCH(e);
  return SMPL_B200_OK;
}

int smpl_b200_full_bwd(const SmplB200Model* m, const SmplB200Parts* parts, const float* params, int N, int img_wh,
                       int vertex_sampling, const float* projects, const float* mask, const float* g_seg,
                       const void* state, float* g_params, void* workspace, size_t workspace_bytes, void* stream) {
  if (N == 0) return SMPL_B200_OK;
  if (!m || !parts || !params || !projects || !mask || !g_seg || !state || !g_params || N < 0) {
    set_error("full_bwd: null/invalid argument (full_fwd must have been given a `state` buffer)"); return SMPL_B200_ERR_BAD_ARG;
  }
  const int vs = vertex_sampling < 1 ? 1 : vertex_sampling;
  const int Vs = (m->V + vs - 1) / vs;
  int rc = seg_check(parts, projects, mask, N, Vs, img_wh, "full_bwd");
  if (rc) return rc;
  const size_t need = smpl_b200_workspace_bytes(m, SMPL_B200_OP_FULL_BWD, N, img_wh, vs);
  if (!workspace || workspace_bytes < need || !aligned(workspace, 16)) {
    set_error("full_bwd: workspace too small or misaligned (%zu < %zu bytes)", workspace_bytes, need); return SMPL_B200_ERR_WORKSPACE;
  }
  const size_t dbytes = decode_ws(m, N, true, false, vs, nullptr).bytes;
  float* g_proj = (float*)((char*)workspace + dbytes);
  const float* vps = (const float*)state;
  const float* A_saved = (const float*)((const char*)state + full_vps_bytes(m, N, vs));
  const unsigned char* saved = (const unsigned char*)state + full_vps_bytes(m, N, vs) + full_A_bytes(N);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = launch_seg_bwd(parts, projects, mask, g_seg, saved, N, Vs, img_wh, g_proj, st);
  if (e == cudaErrorInvalidConfiguration) {
    set_error("full_bwd: part table (%d entries), Vs=%d and img_wh=%d need more than 227 KB of shared memory", parts->E, Vs, img_wh);
    return SMPL_B200_ERR_UNSUPPORTED;
  }
  CHECK_LAUNCH(e);
  return decode_bwd_impl(m, params, N, nullptr, vps, nullptr, g_proj, vs, nullptr, g_params, workspace, dbytes, stream, A_saved);
}

int smpl_b200_silhouette_fwd(const float* projects, int N, int Vs, int img_wh, float* sil, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (N == 0) return SMPL_B200_OK;   // empty batch: nothing to launch
  if (!projects || !sil || N < 0 || Vs < 1 || img_wh < 1) { set_error("silhouette_fwd: null/invalid argument"); return SMPL_B200_ERR_BAD_ARG; }
  if (img_wh > 4096) { set_error("silhouette_fwd: img_wh=%d unsupported", img_wh); return SMPL_B200_ERR_UNSUPPORTED; }
  if (workspace && (workspace_bytes < (size_t)N * img_wh * img_wh * 2 || !aligned(workspace, 2))) {
    set_error("silhouette_fwd: arg-min workspace too small (%zu < %zu bytes)", workspace_bytes, (size_t)N * img_wh * img_wh * 2);
    return SMPL_B200_ERR_WORKSPACE;
  }
  CHECK_LAUNCH(launch_sil_fwd(projects, N, Vs, img_wh, sil, (unsigned short*)workspace, (cudaStream_t)stream));
  return SMPL_B200_OK;
}

int smpl_b200_silhouette_bwd(const float* projects, const float* g_sil, int N, int Vs, int img_wh, float* g_projects,
                             void* workspace, size_t workspace_bytes, void* stream) {
  if (N == 0) return SMPL_B200_OK;   // empty batch: nothing to launch
  if (!projects || !g_sil || !g_projects || N < 0 || Vs < 1 || img_wh < 1) { set_error("silhouette_bwd: null/invalid argument"); return SMPL_B200_ERR_BAD_ARG; }
  if (img_wh > 4096) { set_error("silhouette_bwd: img_wh=%d unsupported", img_wh); return SMPL_B200_ERR_UNSUPPORTED; }
  if (workspace && (workspace_bytes < (size_t)N * img_wh * img_wh * 2 || !aligned(workspace, 2))) {
    set_error("silhouette_bwd: arg-min workspace too small (%zu < %zu bytes)", workspace_bytes, (size_t)N * img_wh * img_wh * 2);
    return SMPL_B200_ERR_WORKSPACE;
  }
  cudaError_t e = launch_sil_bwd(projects, g_sil, N, Vs, img_wh, g_projects, (const unsigned short*)workspace, (cudaStream_t)stream);
  if (e == cudaErrorInvalidConfiguration) { set_error("silhouette_bwd: Vs=%d needs more than 227 KB of shared memory", Vs); return SMPL_B200_ERR_UNSUPPORTED; }
  CHECK_LAUNCH(e);
  return SMPL_B200_OK;
}

int smpl_b200_focal_loss_fwd(const float* seg, const float* y_true, const uint8_t* labels, long long num_pixels,
                             int num_classes, float gamma, const float* class_weights, int from_logits, float* loss,
                             void* stream) {
  if (num_pixels == 0) return SMPL_B200_OK;
  if (!seg || !loss || num_pixels < 0 || (!y_true) == (!labels)) {
    set_error("focal_loss_fwd: null/invalid argument (exactly one of y_true / labels)"); return SMPL_B200_ERR_BAD_ARG;
  }
  if (num_classes < 1 || num_classes > 32) { set_error("focal_loss_fwd: num_classes=%d unsupported (1..32)", num_classes); return SMPL_B200_ERR_UNSUPPORTED; }
  CHECK_LAUNCH(launch_focal_loss_fwd(seg, y_true, labels, num_pixels, num_classes, gamma, class_weights, from_logits, loss,
                                     (cudaStream_t)stream));
  return SMPL_B200_OK;
}

int smpl_b200_focal_loss_bwd(const float* seg, const float* y_true, const uint8_t* labels, const float* g_loss,
                             long long num_pixels, int num_classes, float gamma, const float* class_weights,
                             int from_logits, float* g_seg, void* stream) {
  if (num_pixels == 0) return SMPL_B200_OK;
  if (!seg || !g_loss || !g_seg || num_pixels < 0 || (!y_true) == (!labels)) {
    set_error("focal_loss_bwd: null/invalid argument (exactly one of y_true / labels)"); return SMPL_B200_ERR_BAD_ARG;
  }
  if (num_classes < 1 || num_classes > 32) { set_error("focal_loss_bwd: num_classes=%d unsupported (1..32)", num_classes); return SMPL_B200_ERR_UNSUPPORTED; }
  CHECK_LAUNCH(launch_focal_loss_bwd(seg, y_true, labels, g_loss, num_pixels, num_classes, gamma, class_weights, from_logits,
                                     g_seg, (cudaStream_t)stream));
  return SMPL_B200_OK;
}

// ---- mesh visualiser (renderer.py:23-115,146-197) ------------------------------------------------
void smpl_b200_renderer_destroy(SmplB200Renderer* r) {
  if (!r) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(r->device);
  cudaFree(r->faces); cudaFree(r->adj_ptr); cudaFree(r->adj_face); cudaFree(r->q8);
  cudaSetDevice(prev);
  delete r;
}

int smpl_b200_renderer_create(int device, const int32_t* faces, int num_faces, int num_verts, SmplB200Renderer** out) {
  if (!faces || !out || num_faces < 1 || num_verts < 1) { set_error("renderer_create: bad argument"); return SMPL_B200_ERR_BAD_ARG; }
  *out = nullptr;
  std::vector<int> f(faces, faces + (size_t)num_faces * 3), ptr(num_verts + 1, 0), adj((size_t)num_faces * 3);
  for (size_t i = 0; i < f.size(); ++i) {
    if (f[i] < 0 || f[i] >= num_verts) { set_error("renderer_create: vertex index %d out of range [0,%d)", f[i], num_verts); return SMPL_B200_ERR_BAD_ARG; }
    ++ptr[f[i] + 1];
  }
  for (int v = 0; v < num_verts; ++v) ptr[v + 1] += ptr[v];
  {
    std::vector<int> fill(ptr.begin(), ptr.end() - 1);
    for (int t = 0; t < num_faces; ++t)                    // ascending face index within each vertex's segment
      for (int c = 0; c < 3; ++c) adj[fill[f[(size_t)t * 3 + c]]++] = t;
  }
  std::vector<unsigned char> q8(256);
  for (int k = 0; k < 256; ++k) q8[k] = (unsigned char)(((double)k / 255.0) * 255.0);   // renderer.py:85 on OpenDR's k / 255.
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    set_error("renderer_create: no usable CUDA device %d", device); return SMPL_B200_ERR_NO_DEVICE;
  }
  int prev = 0;
  cudaGetDevice(&prev);
  CU_TRY(cudaSetDevice(device));
  SmplB200Renderer* r = new SmplB200Renderer();
  struct Guard {
    SmplB200Renderer* r; int prev;
    ~Guard() { if (r) smpl_b200_renderer_destroy(r); cudaSetDevice(prev); }
  } guard{r, prev};
  r->device = device; r->V = num_verts; r->F = num_faces;
  CU_TRY(upload(&r->faces, f));
  CU_TRY(upload(&r->adj_ptr, ptr));
  CU_TRY(upload(&r->adj_face, adj));
  CU_TRY(upload(&r->q8, q8));
  CU_TRY(cudaDeviceSynchronize());
  guard.r = nullptr;
  *out = r;
  return SMPL_B200_OK;
}

size_t smpl_b200_render_workspace_bytes(const SmplB200Renderer* r, int N) {
  if (!r || N < 0) return 0;
  // [N][V] screen positions + [N][V] colours (float4 each) + [N][F] packed tile ranges of the faces
  return (size_t)N * r->V * 2 * sizeof(float4) + (((size_t)N * r->F * sizeof(uint32_t) + 15) & ~(size_t)15);
}

int smpl_b200_render(const SmplB200Renderer* r, const float* verts, const float* cam, const float* near_far, int N,
                     int height, int width, const float* albedo, int albedo_per_vertex, const float* lights,
                     int num_lights, const uint8_t* background, int background_per_image, int channels, uint8_t* image,
                     void* workspace, size_t workspace_bytes, void* stream) {
  if (N == 0) return SMPL_B200_OK;
  if (!r || !verts || !cam || !near_far || !albedo || !image || N < 0 || (num_lights > 0 && !lights)) {
    set_error("render: null/invalid argument"); return SMPL_B200_ERR_BAD_ARG;
  }
  if (height < 1 || width < 1 || height > 8192 || width > 8192) { set_error("render: image %dx%d unsupported (1..8192)", height, width); return SMPL_B200_ERR_UNSUPPORTED; }
  if (channels != 3 && channels != 4) { set_error("render: channels=%d (3 or 4)", channels); return SMPL_B200_ERR_BAD_ARG; }
  if (num_lights < 0 || num_lights > kMaxLights) { set_error("render: num_lights=%d unsupported (0..%d)", num_lights, kMaxLights); return SMPL_B200_ERR_UNSUPPORTED; }
  if (N > 65535) { set_error("render: at most 65535 images per call (got %d)", N); return SMPL_B200_ERR_UNSUPPORTED; }
  const size_t need = smpl_b200_render_workspace_bytes(r, N);
  if (!workspace || workspace_bytes < need || !aligned(workspace, 16)) {
    set_error("render: workspace too small or misaligned (%zu < %zu bytes)", workspace_bytes, need); return SMPL_B200_ERR_WORKSPACE;
  }
  RenderLights L;
  L.count = num_lights;
  for (int l = 0; l < num_lights; ++l)
    for (int c = 0; c < 3; ++c) { L.pos[l][c] = lights[l * 6 + c]; L.color[l][c] = lights[l * 6 + 3 + c]; }
  float4* vscreen = reinterpret_cast<float4*>(workspace);
  float4* vcolor = vscreen + (size_t)N * r->V;
  uint32_t* fbox = reinterpret_cast<uint32_t*>(vcolor + (size_t)N * r->V);
  CHECK_LAUNCH(launch_render(r, verts, cam, near_far, N, height, width, albedo, albedo_per_vertex, L, background,
                             background_per_image, channels, vscreen, vcolor, fbox, image, (cudaStream_t)stream));
  return SMPL_B200_OK;
}

}  // extern "C"
