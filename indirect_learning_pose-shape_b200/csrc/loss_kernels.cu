// The op right after the path in every training graph: Reshape -> softmax -> categorical focal loss.
//
// Reference arithmetic (file:line relative to the reference tree), per pixel with scores z (C channels):
//   p = softmax(z)                                       model.py:119-120 (Activation('softmax') on the last axis)
//   p = clip(p, eps, 1 - eps), eps = K.epsilon() = 1e-7  focal_loss.py:15-16
//   ce_c = -y_c log p_c  [ * w_c when weight_classes ]   focal_loss.py:17, 19-41
//   loss = sum_c (1 - p_c)^gamma ce_c                    focal_loss.py:44-45
// Gradient (TF autodiff): d loss / d p_c = w_c y_c [ gamma (1-p_c)^(gamma-1) log p_c - (1-p_c)^gamma / p_c ] where the
// clip passes (eps <= softmax_c <= 1-eps, closed), 0 elsewhere; through the softmax d/dz_j = p_j (a_j - sum_c a_c p_c).
//
// One thread owns one pixel and keeps its C scores in registers.  Rows move between global and shared memory as
// block-wide coalesced float4 copies (a thread reading its own 128-byte row straight from global memory touches 32
// lines per instruction: measured 3.4 / 10.1 ms for the two kernels at 16384 x 48 x 48); the shared tile is padded to a
// 36-float row stride so the per-thread float4 row reads are conflict-free.  Both kernels are pure streaming.
#include <math_constants.h>
#include "common.cuh"

namespace smplb200 {

namespace {

constexpr float kKerasEps = 1e-7f;          // K.epsilon()
constexpr int kMaxC = 32;

constexpr int kPix = 128;                   // pixels (= threads) per block
constexpr int kRow = 36;                    // shared row stride in floats: 16-byte aligned and conflict-free for float4 rows

// Block-wide coalesced copy of kPix rows of Crt floats between global memory and the padded shared tile.
template <bool TO_SMEM>
__device__ __forceinline__ void copy_tile(float* tile, const float* __restrict__ gsrc, float* __restrict__ gdst, int Crt,
                                          long long first, long long npix, bool vec) {
  const int rows = (int)min((long long)kPix, npix - first);
  if (vec) {                                // Crt == 32, 16-byte aligned: 8 float4 per row
    for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
      const int r = i >> 3, k = i & 7;
      if (TO_SMEM) *reinterpret_cast<float4*>(tile + r * kRow + k * 4) = reinterpret_cast<const float4*>(gsrc + first * 32)[i];
      else reinterpret_cast<float4*>(gdst + first * 32)[i] = *reinterpret_cast<const float4*>(tile + r * kRow + k * 4);
    }
  } else {
    for (int i = threadIdx.x; i < rows * Crt; i += blockDim.x) {
      const int r = i / Crt, c = i - r * Crt;
      if (TO_SMEM) tile[r * kRow + c] = gsrc[first * Crt + i];
      else gdst[first * Crt + i] = tile[r * kRow + c];
    }
  }
}

template <int C>
__device__ __forceinline__ void read_row(const float* tile_row, int Crt, float pad, float (&v)[C]) {
#pragma unroll
  for (int k = 0; k < C / 4; ++k) {
    const float4 t = *reinterpret_cast<const float4*>(tile_row + 4 * k);
    v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
  }
#pragma unroll
  for (int c = 0; c < C; ++c) v[c] = c < Crt ? v[c] : pad;
}

__device__ __forceinline__ float fast_exp2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float pow_gamma(float x, float gamma) {     // K.pow(1 - p, gamma); gamma = 2 is the training value
  return gamma == 2.0f ? x * x : powf(x, gamma);
}

// p (clipped), gate (1 where the clip passes) for one pixel
template <int C>
__device__ __forceinline__ void probabilities(const float (&z)[C], int Crt, bool from_logits, float (&p)[C], float (&q)[C]) {
  if (from_logits) {
    float m = -CUDART_INF_F;
#pragma unroll
    for (int c = 0; c < C; ++c) m = fmaxf(m, z[c]);
    float Z = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {                                      // ex2.approx: 2^-22 relative, far inside the 2e-6 parity bound
      q[c] = c < Crt ? fast_exp2((z[c] - m) * 1.4426950408889634f) : 0.f;
      Z += q[c];
    }
    const float inv = 1.0f / Z;
#pragma unroll
    for (int c = 0; c < C; ++c) q[c] *= inv;                           // q = softmax (unclipped)
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) q[c] = c < Crt ? z[c] : 0.f;
  }
#pragma unroll
  for (int c = 0; c < C; ++c) p[c] = fminf(fmaxf(q[c], kKerasEps), 1.0f - kKerasEps);
}

template <bool SOFT, bool BWD>
__global__ void __launch_bounds__(kPix)
focal_loss_kernel(const float* __restrict__ seg, const float* __restrict__ y_true, const uint8_t* __restrict__ labels,
                  const float* __restrict__ g_loss, long long npix, int Crt, float gamma,
                  const float* __restrict__ class_w, int from_logits, int vec, float* __restrict__ out) {
  constexpr int C = kMaxC;
  __shared__ float sw[kMaxC];
  __shared__ __align__(16) float ztile[kPix * kRow];
  __shared__ __align__(16) float ytile[SOFT ? kPix * kRow : 4];
  if (threadIdx.x < kMaxC) sw[threadIdx.x] = (class_w && (int)threadIdx.x < Crt) ? class_w[threadIdx.x] : 1.0f;
  const long long first = (long long)blockIdx.x * kPix;
  copy_tile<true>(ztile, seg, nullptr, Crt, first, npix, vec != 0);
  if (SOFT) copy_tile<true>(ytile, y_true, nullptr, Crt, first, npix, vec != 0);
  __syncthreads();
  const long long i = first + threadIdx.x;
  const bool live = i < npix;
  float z[C], p[C], q[C], y[C];
  read_row<C>(ztile + threadIdx.x * kRow, Crt, -CUDART_INF_F, z);
  probabilities<C>(z, Crt, from_logits != 0, p, q);
  if (SOFT) read_row<C>(ytile + threadIdx.x * kRow, Crt, 0.f, y);
  if (!SOFT) {
    // class ids: one class contributes; pick its probability with selects (no dynamically indexed registers) and do the
    // transcendental work once per pixel, not once per (pixel, class) under divergent predicates
    const int lab = live ? labels[i] : 0;
    float pl = 1.0f, ql = 1.0f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      pl = (c == lab) ? p[c] : pl;
      ql = (c == lab) ? q[c] : ql;
    }
    const float wl = sw[min(lab, kMaxC - 1)];
    const bool inlab = lab < Crt;
    const float om = 1.0f - pl, lg = logf(pl);
    if (!BWD) {
      if (live) out[i] = inlab ? pow_gamma(om, gamma) * ((-lg) * wl) : 0.f;       // focal_loss.py:17, 40, 44-45
    } else {
      const float g = live ? g_loss[i] : 0.f;
      float al = 0.f;
      if (inlab && ql >= kKerasEps && ql <= 1.0f - kKerasEps) {
        const float dpow = gamma == 2.0f ? 2.0f * om : gamma * powf(om, gamma - 1.0f);
        al = wl * (dpow * lg - pow_gamma(om, gamma) / pl);
      }
      const float dot = al * ql;
      float* row = ztile + threadIdx.x * kRow;                         // this thread's own row: no hazard before the sync
#pragma unroll
      for (int k = 0; k < C / 4; ++k) {
        float t[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = 4 * k + e;
          const float ac = (c == lab) ? al : 0.f;
          t[e] = from_logits ? g * q[c] * (ac - dot) : g * ac;
        }
        *reinterpret_cast<float4*>(row + 4 * k) = make_float4(t[0], t[1], t[2], t[3]);
      }
      __syncthreads();
      copy_tile<false>(ztile, nullptr, out, Crt, first, npix, vec != 0);
    }
    return;
  }
  if (!BWD) {
    float loss = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (c < Crt && y[c] != 0.f) {
        const float ce = (-y[c] * logf(p[c])) * sw[c];                 // focal_loss.py:17, 40
        loss += pow_gamma(1.0f - p[c], gamma) * ce;                    // :44-45
      }
    }
    if (live) out[i] = loss;
  } else {
    const float g = live ? g_loss[i] : 0.f;
    float a[C];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      a[c] = 0.f;
      if (c < Crt && y[c] != 0.f && q[c] >= kKerasEps && q[c] <= 1.0f - kKerasEps) {
        const float om = 1.0f - p[c];
        const float dpow = gamma == 2.0f ? 2.0f * om : gamma * powf(om, gamma - 1.0f);
        a[c] = sw[c] * y[c] * (dpow * logf(p[c]) - pow_gamma(om, gamma) / p[c]);
      }
      dot += a[c] * q[c];
    }
    float* row = ztile + threadIdx.x * kRow;                           // this thread's own row: no hazard before the sync
#pragma unroll
    for (int k = 0; k < C / 4; ++k) {
      float4 t;
      t.x = from_logits ? g * q[4 * k] * (a[4 * k] - dot) : g * a[4 * k];
      t.y = from_logits ? g * q[4 * k + 1] * (a[4 * k + 1] - dot) : g * a[4 * k + 1];
      t.z = from_logits ? g * q[4 * k + 2] * (a[4 * k + 2] - dot) : g * a[4 * k + 2];
      t.w = from_logits ? g * q[4 * k + 3] * (a[4 * k + 3] - dot) : g * a[4 * k + 3];
      *reinterpret_cast<float4*>(row + 4 * k) = t;
    }
    __syncthreads();
    copy_tile<false>(ztile, nullptr, out, Crt, first, npix, vec != 0);
  }
}

template <bool BWD>
cudaError_t launch_focal(const float* seg, const float* y_true, const uint8_t* labels, const float* g_loss, long long npix,
                         int C, float gamma, const float* class_w, int from_logits, float* out, cudaStream_t st) {
  if (C < 1 || C > kMaxC) return cudaErrorInvalidValue;
  const int vec = C == kMaxC && ((uintptr_t)seg % 16 == 0) && (!y_true || (uintptr_t)y_true % 16 == 0) &&
                  (!BWD || (uintptr_t)out % 16 == 0);
  const unsigned blocks = (unsigned)((npix + kPix - 1) / kPix);
  LaunchScope scope(BWD ? KID_FOCAL_BWD : KID_FOCAL_FWD, st);
  if (y_true) focal_loss_kernel<true, BWD><<<blocks, kPix, 0, st>>>(seg, y_true, labels, g_loss, npix, C, gamma, class_w, from_logits, vec, out);
  else focal_loss_kernel<false, BWD><<<blocks, kPix, 0, st>>>(seg, y_true, labels, g_loss, npix, C, gamma, class_w, from_logits, vec, out);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_focal_loss_fwd(const float* seg, const float* y_true, const uint8_t* labels, long long npix, int C,
                                  float gamma, const float* class_w, int from_logits, float* loss, cudaStream_t st) {
  return launch_focal<false>(seg, y_true, labels, nullptr, npix, C, gamma, class_w, from_logits, loss, st);
}
cudaError_t launch_focal_loss_bwd(const float* seg, const float* y_true, const uint8_t* labels, const float* g_loss,
                                  long long npix, int C, float gamma, const float* class_w, int from_logits, float* g_seg,
                                  cudaStream_t st) {
  return launch_focal<true>(seg, y_true, labels, g_loss, npix, C, gamma, class_w, from_logits, g_seg, st);
}

}  // namespace smplb200
