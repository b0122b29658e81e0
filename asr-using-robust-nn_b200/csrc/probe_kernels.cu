// Measurement probes (not on the product path): the FP32 FMA peak of the device bench.py's FP32 roofline is quoted against
// ("measure an FMA-chain FP32 peak on the box in the same run").
#include "common.cuh"

namespace asr {

// 8 independent FMA chains per thread, `iters` rounds of 64 FMAs each; the result is stored so that nothing is optimised away.
__global__ void __launch_bounds__(512) fma_peak_kernel(const int iters, const float a, const float b, float* __restrict__ sink) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.0f, x2 = x0 + 2.0f, x3 = x0 + 3.0f, x4 = x0 + 4.0f, x5 = x0 + 5.0f, x6 = x0 + 6.0f,
        x7 = x0 + 7.0f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace asr

using namespace asr;

// Launches the probe: n_blocks x 512 threads x iters x 64 FMAs (x 2 flops).  sink_dev: n_blocks * 512 floats.
extern "C" int asr_fp32_peak_probe(int32_t n_blocks, int32_t iters, float* sink_dev, void* stream) {
  if (n_blocks < 1 || iters < 1 || !sink_dev) { set_error("asr_fp32_peak_probe: bad argument"); return ASR_ERR_INVALID; }
  fma_peak_kernel<<<n_blocks, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(iters, 0.999f, 0.001f, sink_dev);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}
