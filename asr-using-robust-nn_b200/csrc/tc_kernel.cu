// Tensor-core (tcgen05 + TMEM) kernels of the n_fft = 512 front end, and their self-test.
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace asr {

// ------------------------------------------------------------------------------------------------
// Self-test of the tcgen05 plumbing this library relies on (descriptor encoding, K-major unswizzled operand tiles written
// by threads, accumulation over K steps and over several products, commit -> mbarrier, TMEM loads by lane quarter):
//   D[128][32] = A1[128][32] * B1[32][32]^T + A2 * B2^T      (float16 operands, float32 accumulation)
// One CTA of 128 threads.  tests/test_tc_gpu.py compares with numpy.
__global__ void __launch_bounds__(128) tc_selftest_kernel(const __half* __restrict__ a1, const __half* __restrict__ b1,
                                                          const __half* __restrict__ a2, const __half* __restrict__ b2,
                                                          float* __restrict__ d) {
  constexpr int M = 128, N = 32, K = 32, CH = K / 8;       // CH 16-byte chunks per row
  __shared__ __align__(128) uint4 sA[2][CH][M];
  __shared__ __align__(128) uint4 sB[2][CH][N];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  // operand tiles, chunk-major: tile[c][row]
  for (int p = 0; p < 2; ++p) {
    const uint4* ga = reinterpret_cast<const uint4*>(p ? a2 : a1);     // row-major [M][K] halfs = [M][CH] uint4
    const uint4* gb = reinterpret_cast<const uint4*>(p ? b2 : b1);     // [N][K]
    for (int c = 0; c < CH; ++c) sA[p][c][tid] = ga[tid * CH + c];
    if (tid < N)
      for (int c = 0; c < CH; ++c) sB[p][c][tid] = gb[tid * CH + c];
  }
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_init_fence(); }
  if (warp == 0) tc::tmem_alloc(&tmem_base, 32);
  tc::fence_async_smem();                                  // generic-proxy stores -> visible to the tensor core (async proxy)
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = tc::idesc_f16_f32(M, N);
    uint32_t acc = 0;
    for (int p = 0; p < 2; ++p)
      for (int ks = 0; ks < K / 16; ++ks) {                // one MMA covers K = 16 = two chunks
        const uint64_t ad = tc::smem_desc(tc::smem_u32(&sA[p][2 * ks][0]), 16 * M, 128);
        const uint64_t bd = tc::smem_desc(tc::smem_u32(&sB[p][2 * ks][0]), 16 * N, 128);
        tc::mma_f16(tm, ad, bd, idesc, acc);
        acc = 1;
      }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  uint32_t r[32];
  tc::tmem_ld32(tm + (static_cast<uint32_t>(32 * warp) << 16), r);
  tc::tmem_ld_wait();
  for (int j = 0; j < N; ++j) d[tid * N + j] = __uint_as_float(r[j]);
  // strided two-column reads (what the FFT epilogue does): columns 2j, 2j+1 again, into the second half of the output
  for (int j = 0; j < N / 2; ++j) {
    float v[2];
    tc::tmem_ld2(tm + (static_cast<uint32_t>(32 * warp) << 16) + 2 * j, v);
    tc::tmem_ld_wait();
    d[M * N + tid * N + 2 * j] = v[0];
    d[M * N + tid * N + 2 * j + 1] = v[1];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, 32);
}

}  // namespace asr

using namespace asr;

extern "C" int asr_tc_selftest(const void* a1_dev, const void* b1_dev, const void* a2_dev, const void* b2_dev, float* d_dev,
                               void* stream) {
  if (!a1_dev || !b1_dev || !a2_dev || !b2_dev || !d_dev) { set_error("asr_tc_selftest: null pointer"); return ASR_ERR_INVALID; }
  tc_selftest_kernel<<<1, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(a1_dev), static_cast<const __half*>(b1_dev), static_cast<const __half*>(a2_dev),
      static_cast<const __half*>(b2_dev), d_dev);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}
