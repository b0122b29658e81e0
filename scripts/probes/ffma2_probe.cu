// Probe: throughput of packed FP32 (fma.rn.f32x2 -> FFMA2) against scalar FFMA on sm_100a, alone and mixed with
// shared-memory loads and integer instructions.  Standalone: nvcc -arch=sm_100a -o ffma2_probe ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float lo(unsigned long long a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); return x + y; }

template <int MODE>
__global__ void __launch_bounds__(512) k(const int iters, const float a, const float b, float* __restrict__ sink) {
  __shared__ float sh[1024];
  sh[threadIdx.x] = threadIdx.x; sh[threadIdx.x + 512] = 1.0f;
  __syncthreads();
  float s = 0.0f;
  if (MODE == 0) {          // scalar FFMA, 16 chains
    float x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = threadIdx.x * 1e-3f + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = fmaf(x[j], a, b);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) s += x[j];
  } else if (MODE == 1) {   // FFMA2, 8 packed chains = 16 scalar chains
    unsigned long long x[8];
    const unsigned long long A = pk(a, a), B = pk(b, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = pk(threadIdx.x * 1e-3f + j, threadIdx.x * 1e-3f + j + 8);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = fma2(x[j], A, B);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s += lo(x[j]);
  } else if (MODE == 2) {   // FFMA2 16 packed chains (more ILP)
    unsigned long long x[16];
    const unsigned long long A = pk(a, a), B = pk(b, b);
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = pk(threadIdx.x * 1e-3f + j, threadIdx.x * 1e-3f + j + 8);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = fma2(x[j], A, B);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) s += lo(x[j]);
  } else if (MODE == 3) {   // scalar FFMA 16 chains + 1 LDS per 4 FFMA
    float x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = threadIdx.x * 1e-3f + j;
    int idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = fmaf(x[j], a, b);
#pragma unroll
        for (int q = 0; q < 4; ++q) { s += sh[(idx + 32 * q + 128 * u) & 1023]; }
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) s += x[j];
  } else if (MODE == 4) {   // FFMA2 8 chains + 1 LDS per 2 FFMA2 (same work as MODE 3)
    unsigned long long x[8];
    const unsigned long long A = pk(a, a), B = pk(b, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = pk(threadIdx.x * 1e-3f + j, threadIdx.x * 1e-3f + j + 8);
    int idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = fma2(x[j], A, B);
#pragma unroll
        for (int q = 0; q < 4; ++q) { s += sh[(idx + 32 * q + 128 * u) & 1023]; }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s += lo(x[j]);
  } else if (MODE == 5) {   // add.f32x2 8 chains
    unsigned long long x[8];
    const unsigned long long B = pk(b, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = pk(threadIdx.x * 1e-3f + j, threadIdx.x * 1e-3f + j + 8);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = add2(x[j], B);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s += lo(x[j]);
  }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int fma_per_iter_thread) {
  const int nb = 148 * 4, it = 4096;
  float* sink; cudaMalloc(&sink, nb * 512 * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(a);
    k<MODE><<<nb, 512>>>(it, 0.999f, 0.001f, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (r && ms < best) best = ms;
  }
  const double fma = double(nb) * 512 * it * fma_per_iter_thread;
  printf("%-40s %8.3f ms  %7.2f TFLOP/s (2 flop per scalar FMA/ADD lane-op counted as FMA)\n", name, best, fma * 2 / (best * 1e-3) / 1e12);
  cudaFree(sink);
}

int main() {
  run<0>("scalar FFMA x16 chains", 64);
  run<1>("FFMA2 x8 packed chains", 64);
  run<2>("FFMA2 x16 packed chains", 64);
  run<3>("scalar FFMA + LDS (4:1)", 64);
  run<4>("FFMA2 + LDS (2:1)", 64);
  run<5>("FADD2 x8 packed chains", 64);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
