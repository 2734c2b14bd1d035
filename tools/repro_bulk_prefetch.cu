// Stand-alone check of cp.async.bulk.prefetch.L2 against the loads that follow it ("a withdrawn
// optimisation").  A kernel shaped like the seg backward's main loop: one block of six warps per "sample", a warp walks
// 48-pixel rows, lane = channel, every pixel is one 128-byte row of floats read with ld.global.nc plus one byte per lane of
// a second buffer.  Both buffers hold a known function of the index; every loaded value is compared with it.
//   mode 0: no prefetch    mode 1: one lane requests the warp's NEXT row with cp.async.bulk.prefetch.L2
//   mode 2: the same rows requested with prefetch.global.L2, one 128-byte line per lane
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/repro_bulk_prefetch.bin tools/repro_bulk_prefetch.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kWh = 48, kPx = kWh * kWh, kC = 32;

__host__ __device__ inline float pat_f(unsigned long long i) { return (float)((unsigned)(i * 2654435761ull) >> 8) * (1.0f / 16777216.0f); }
__host__ __device__ inline unsigned char pat_b(unsigned long long i) { return (unsigned char)((i * 40503ull + 17ull) >> 3); }

__global__ void fill(float* g, unsigned char* s, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    g[i] = pat_f(i);
    s[i] = pat_b(i);
  }
}

__global__ void __launch_bounds__(192) walk(const float* __restrict__ g, const unsigned char* __restrict__ s, int mode,
                                            unsigned long long* bad, float* sink) {
  const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const size_t base = (size_t)n * kPx * kC;
  float acc = 0.f;
  unsigned long long wrong = 0;
  for (int r = warp; r < kWh; r += nwarps) {
    const int rn = r + nwarps;
    if (rn < kWh) {
      const float* gp = g + base + (size_t)rn * kWh * kC;
      const unsigned char* sp = s + base + (size_t)rn * kWh * kC;
      if (mode == 1 && lane < 2) {
        if (lane == 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gp), "r"(kWh * kC * 4) : "memory");
        else asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(sp), "r"(kWh * kC) : "memory");
      } else if (mode == 2) {
        for (int o = lane * 128; o < kWh * kC * 4; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)gp + o));
        if (lane * 128 < kWh * kC) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + lane * 128));
      }
    }
    for (int c = 0; c < kWh; c += 4) {
      float v[4];
      unsigned b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const size_t i = base + ((size_t)r * kWh + c + j) * kC + lane;
        v[j] = __ldg(g + i);
        b[j] = __ldg(s + i);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const size_t i = base + ((size_t)r * kWh + c + j) * kC + lane;
        wrong += (v[j] != pat_f(i)) + (b[j] != pat_b(i));
        acc += v[j] * (float)b[j];
      }
    }
  }
  if (wrong) atomicAdd(bad, wrong);
  if (acc == 123.456f) sink[0] = acc;
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 16384, reps = argc > 2 ? atoi(argv[2]) : 10;
  const size_t n = (size_t)N * kPx * kC;
  float* g; unsigned char* s; unsigned long long* bad; float* sink;
  if (cudaMalloc(&g, n * 4) || cudaMalloc(&s, n) || cudaMalloc(&bad, 8) || cudaMalloc(&sink, 4)) { printf("alloc failed\n"); return 1; }
  fill<<<4096, 256>>>(g, s, n);
  cudaDeviceSynchronize();
  for (int mode = 0; mode < 3; ++mode) {
    unsigned long long total = 0;
    float ms_sum = 0.f;
    for (int r = 0; r < reps; ++r) {
      cudaMemset(bad, 0, 8);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      walk<<<N, 192>>>(g, s, mode, bad, sink);
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms_sum += ms;
      unsigned long long h; cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost); total += h;
    }
    printf("mode %d (%s): %llu wrong values in %d launches of %d samples (%.3f ms per launch); %s\n", mode,
           mode == 0 ? "no prefetch" : mode == 1 ? "cp.async.bulk.prefetch.L2" : "prefetch.global.L2", total, reps, N,
           ms_sum / reps, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
