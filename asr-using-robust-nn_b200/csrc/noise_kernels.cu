// Additive-noise path for sm_100a (reference: VDR/attacks.py:73-86,145-183,222-245;
// SR/attacks.py:81-94,149-189,228-251).
//
//   clip_power_kernel  P = np.mean(sample**2) in float32, bit-exact with numpy's pairwise summation
//   snr_sigma_kernel   the float32 scalar chain of add_white_noise_with_snr
//   mix_*_kernel       float64(x) + s*z with two separately rounded float64 operations
//   randn_kernel       seeded Philox4x32-10 + Box-Muller standard-normal stream
#include "common.cuh"

namespace asr {

// ------------------------------------------------------------------------------------------------
// numpy pairwise_sum(a, n):
//   n < 8    : serial
//   n <= 128 : 8 strided accumulators, ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail serially
//   else     : n2 = n/2 - (n/2 % 8) ; pairwise(a, n2) + pairwise(a+n2, n-n2)
// The recursion tree only depends on n.  It is laid out as a binary heap in shared memory
// (node i -> children 2i, 2i+1), leaves are summed one thread each, parents level by level.
constexpr int kHeap = 512;        // 9 levels: enough for any sub-tree of <= kSubMax elements
constexpr int kHeapLevels = 8;    // levels that can still split
constexpr int kSubMax = 16384;    // elements staged in shared memory per sub-tree (64 KB)

struct Heap {
  int off[kHeap];
  int len[kHeap];
  float val[kHeap];
};

// Expand node 1 = (off0, n0): nodes longer than `limit` split the way numpy does.
__device__ void heap_expand(Heap& h, const int off0, const int n0, const int limit) {
  const int tid = threadIdx.x;
  for (int i = tid; i < kHeap; i += blockDim.x) h.len[i] = 0;
  __syncthreads();
  if (tid == 0) { h.off[1] = off0; h.len[1] = n0; }
  __syncthreads();
  for (int d = 0; d < kHeapLevels; ++d) {
    const int first = 1 << d;
    for (int t = tid; t < first; t += blockDim.x) {
      const int i = first + t;
      const int n = h.len[i];
      if (n > limit) {
        int n2 = n / 2;
        n2 -= n2 % 8;
        h.off[2 * i] = h.off[i];         h.len[2 * i] = n2;
        h.off[2 * i + 1] = h.off[i] + n2; h.len[2 * i + 1] = n - n2;
      }
    }
    __syncthreads();
  }
}

// Sum parents bottom-up; a node is a parent iff its length exceeds `limit`.
__device__ void heap_combine(Heap& h, const int limit) {
  const int tid = threadIdx.x;
  for (int d = kHeapLevels - 1; d >= 0; --d) {
    const int first = 1 << d;
    for (int t = tid; t < first; t += blockDim.x) {
      const int i = first + t;
      if (h.len[i] > limit) h.val[i] = __fadd_rn(h.val[2 * i], h.val[2 * i + 1]);
    }
    __syncthreads();
  }
}

// One leaf (n <= 128) summed by a group of 8 lanes, lane j owning numpy's accumulator r[j]; the xor-shuffle
// tree reproduces ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) (float addition is commutative), lane 0 adds the tail.
__device__ __forceinline__ float leaf_sum8(const float* __restrict__ a, const int n, const int j, const unsigned gmask) {
  if (n < 8) {
    float res = 0.0f;
    if (j == 0)
      for (int i = 0; i < n; ++i) res = __fadd_rn(res, a[i]);
    return res;
  }
  const int n8 = n - (n % 8);
  float r = a[j];
  for (int i = 8; i < n8; i += 8) r = __fadd_rn(r, a[i + j]);
  r = __fadd_rn(r, __shfl_xor_sync(gmask, r, 1));
  r = __fadd_rn(r, __shfl_xor_sync(gmask, r, 2));
  r = __fadd_rn(r, __shfl_xor_sync(gmask, r, 4));
  if (j == 0)
    for (int i = n8; i < n; ++i) r = __fadd_rn(r, a[i]);
  return r;
}

__global__ void __launch_bounds__(256) clip_power_kernel(const void* __restrict__ audio, const int dtype,
                                                         const long long* __restrict__ offsets,
                                                         const int* __restrict__ lengths, float* __restrict__ power,
                                                         const int vec_ok) {
  extern __shared__ __align__(16) float sq[];   // kSubMax squares
  __shared__ Heap top, sub;
  __shared__ int leaf_list[kHeap];
  __shared__ int n_leaves;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int L = lengths[b];
  const long long base = offsets[b];
  if (L <= 0 || L > 4000000) {                  // np.mean of an empty array is nan; > 4M samples exceeds the heap depth
    if (tid == 0) power[b] = __int_as_float(0x7fc00000);
    return;
  }
  heap_expand(top, 0, L, kSubMax);
  for (int node = 1; node < kHeap; ++node) {
    const int n = top.len[node];
    if (n <= 0 || n > kSubMax) continue;          // absent or internal (uniform over the block)
    const int off = top.off[node];
    // ---- stage sample**2 (exact float32 products) for this sub-tree ----
    const long long e0 = base + off;
    if (vec_ok && (e0 & 7) == 0) {
      const int n8 = n & ~7;
      if (dtype == ASR_I16) {
        const int4* p = reinterpret_cast<const int4*>(reinterpret_cast<const short*>(audio) + e0);
        for (int g = tid; g < n8 / 8; g += blockDim.x) {
          const int4 raw = __ldg(p + g);
          const int w[4] = {raw.x, raw.y, raw.z, raw.w};
          float v[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x0 = static_cast<float>(static_cast<short>(w[j])) * (1.0f / 32768.0f);
            const float x1 = static_cast<float>(w[j] >> 16) * (1.0f / 32768.0f);
            v[2 * j] = __fmul_rn(x0, x0);
            v[2 * j + 1] = __fmul_rn(x1, x1);
          }
          float4* d = reinterpret_cast<float4*>(sq + 8 * g);
          d[0] = make_float4(v[0], v[1], v[2], v[3]);
          d[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      } else {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(audio) + e0);
        for (int g = tid; g < n8 / 4; g += blockDim.x) {
          const float4 x = __ldg(p + g);
          reinterpret_cast<float4*>(sq)[g] = make_float4(__fmul_rn(x.x, x.x), __fmul_rn(x.y, x.y), __fmul_rn(x.z, x.z),
                                                        __fmul_rn(x.w, x.w));
        }
      }
      for (int i = n8 + tid; i < n; i += blockDim.x) {
        float x;
        if (dtype == ASR_I16) x = static_cast<float>(__ldg(reinterpret_cast<const short*>(audio) + e0 + i)) * (1.0f / 32768.0f);
        else x = __ldg(reinterpret_cast<const float*>(audio) + e0 + i);
        sq[i] = __fmul_rn(x, x);
      }
    } else {
      for (int i = tid; i < n; i += blockDim.x) {
        float x;
        if (dtype == ASR_I16) x = static_cast<float>(__ldg(reinterpret_cast<const short*>(audio) + e0 + i)) * (1.0f / 32768.0f);
        else x = __ldg(reinterpret_cast<const float*>(audio) + e0 + i);
        sq[i] = __fmul_rn(x, x);
      }
    }
    if (tid == 0) n_leaves = 0;
    heap_expand(sub, 0, n, 128);                   // barriers inside also publish sq[] and n_leaves
    for (int i = 1 + tid; i < kHeap; i += blockDim.x) {
      const int ln = sub.len[i];
      if (ln > 0 && ln <= 128) leaf_list[atomicAdd(&n_leaves, 1)] = i;
    }
    __syncthreads();
    {
      const int j = tid & 7;
      const unsigned gmask = 0xFFu << (threadIdx.x & 24);
      const int nl = n_leaves;
      for (int li = tid >> 3; li < nl; li += blockDim.x >> 3) {
        const int nd = leaf_list[li];
        const float r = leaf_sum8(sq + sub.off[nd], sub.len[nd], j, gmask);
        if (j == 0) sub.val[nd] = r;
      }
    }
    __syncthreads();
    heap_combine(sub, 128);
    if (tid == 0) top.val[node] = sub.val[1];
    __syncthreads();
  }
  heap_combine(top, kSubMax);
  // np.mean: float32 sum / count evaluated in float64, rounded to float32
  if (tid == 0) power[b] = static_cast<float>(static_cast<double>(top.val[1]) / static_cast<double>(L));
}

__global__ void snr_sigma_kernel(const float* __restrict__ power, const float snr_db, double* __restrict__ sigma,
                                 const int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float p = power[i];
  const float lg = static_cast<float>(log10(static_cast<double>(p)));      // np.log10(float32)
  const float sdb = __fmul_rn(10.0f, lg);                                    // 10 * ...
  const float ndb = __fsub_rn(sdb, snr_db);                                  // - target_snr_db
  const float t = __fdiv_rn(ndb, 10.0f);                                     // / 10
  const float w = static_cast<float>(pow(10.0, static_cast<double>(t)));    // 10 ** ...
  sigma[i] = static_cast<double>(__fsqrt_rn(w));                             // np.sqrt
}

__device__ __forceinline__ double audio_f64(const void* __restrict__ audio, const int dtype, const long long i) {
  if (dtype == ASR_I16) return static_cast<double>(static_cast<float>(__ldg(reinterpret_cast<const short*>(audio) + i)) * (1.0f / 32768.0f));
  if (dtype == ASR_F32) return static_cast<double>(__ldg(reinterpret_cast<const float*>(audio) + i));
  return __ldg(reinterpret_cast<const double*>(audio) + i);
}

__global__ void __launch_bounds__(256) mix_white_kernel(const void* __restrict__ audio, const int dtype,
                                                        const long long* __restrict__ offsets,
                                                        const int* __restrict__ lengths, const double* __restrict__ z,
                                                        const double* __restrict__ sigma, double* __restrict__ out) {
  const int b = blockIdx.x;
  const int L = lengths[b];
  const long long base = offsets[b];
  const double s = sigma[b];
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < L; i += gridDim.y * blockDim.x)
    out[base + i] = __dadd_rn(audio_f64(audio, dtype, base + i), __dmul_rn(s, __ldg(z + base + i)));
}

__global__ void __launch_bounds__(256) mix_mixture_kernel(const void* __restrict__ audio, const int dtype,
                                                          const long long* __restrict__ offsets,
                                                          const int* __restrict__ lengths, const double* __restrict__ q,
                                                          const double* __restrict__ g, const double p, const double s0,
                                                          const double s1, double* __restrict__ out) {
  const int b = blockIdx.x;
  const int L = lengths[b];
  const long long base = offsets[b];
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < L; i += gridDim.y * blockDim.x) {
    const double sel = (fabs(__ldg(q + base + i)) < p) ? s1 : s0;
    out[base + i] = __dadd_rn(audio_f64(audio, dtype, base + i), __dmul_rn(sel, __ldg(g + base + i)));
  }
}

__global__ void __launch_bounds__(256) mix_rows_kernel(const double* __restrict__ x, const long long n,
                                                       const double* __restrict__ q, const double* __restrict__ g,
                                                       const int mixture, const double p, const double s0,
                                                       const double s1, double* __restrict__ out) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    double sel = s0;
    if (mixture) sel = (fabs(__ldg(q + i)) < p) ? s1 : s0;
    out[i] = __dadd_rn(__ldg(x + i), __dmul_rn(sel, __ldg(g + i)));
  }
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11); counter = element index / 4, key = seed.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&o)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

__global__ void __launch_bounds__(256) randn_kernel(const unsigned long long seed, const unsigned long long first,
                                                    const long long n, double* __restrict__ out) {
  // one thread per group of 4 consecutive GLOBAL indices (aligned to 4), so a value only depends on (seed, index)
  const unsigned long long g0 = first / 4;
  const unsigned long long g1 = (first + n + 3) / 4;
  for (unsigned long long g = g0 + static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; g < g1;
       g += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
    uint32_t r[4];
    philox4x32_10(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), 0u, 0u, static_cast<uint32_t>(seed),
                  static_cast<uint32_t>(seed >> 32), r);
    float zf[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float u1 = (static_cast<float>(r[2 * h] >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0,1)
      const float u2 = (static_cast<float>(r[2 * h + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
      const float rad = sqrtf(-2.0f * logf(u1));
      float s, c;
      sincospif(2.0f * u2, &s, &c);
      zf[2 * h] = rad * c;
      zf[2 * h + 1] = rad * s;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned long long idx = g * 4 + j;
      if (idx >= first && idx < first + static_cast<unsigned long long>(n)) out[idx - first] = static_cast<double>(zf[j]);
    }
  }
}

}  // namespace asr

// ------------------------------------------------------------------------------------------------
using namespace asr;

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" int asr_clip_power(const void* audio_dev, int32_t dtype, const int64_t* offsets_dev,
                              const int32_t* lengths_dev, int32_t n_clips, float* power_dev, void* stream) {
  if (!audio_dev || !offsets_dev || !lengths_dev || !power_dev || n_clips < 0) {
    set_error("asr_clip_power: null pointer or negative count");
    return ASR_ERR_INVALID;
  }
  if (dtype != ASR_I16 && dtype != ASR_F32) {
    set_error("asr_clip_power: dtype must be ASR_I16 or ASR_F32 (the reference's audio is float32)");
    return ASR_ERR_INVALID;
  }
  if (n_clips == 0) return ASR_OK;
  static bool attr_done = false;
  if (!attr_done) {
    ASR_CUDA_TRY(cudaFuncSetAttribute(clip_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSubMax * 4));
    attr_done = true;
  }
  clip_power_kernel<<<n_clips, 256, kSubMax * 4, as_stream(stream)>>>(
      audio_dev, dtype, reinterpret_cast<const long long*>(offsets_dev), lengths_dev, power_dev,
      (reinterpret_cast<uintptr_t>(audio_dev) & 15) == 0 ? 1 : 0);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}

extern "C" int asr_snr_sigma(const float* power_dev, float target_snr_db, double* sigma_dev, int32_t n_clips,
                             void* stream) {
  if (!power_dev || !sigma_dev || n_clips < 0) {
    set_error("asr_snr_sigma: null pointer or negative count");
    return ASR_ERR_INVALID;
  }
  if (n_clips == 0) return ASR_OK;
  snr_sigma_kernel<<<(n_clips + 255) / 256, 256, 0, as_stream(stream)>>>(power_dev, target_snr_db, sigma_dev, n_clips);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}

extern "C" int asr_mix_white(const void* audio_dev, int32_t dtype, const int64_t* offsets_dev,
                             const int32_t* lengths_dev, int32_t n_clips, const double* z_dev,
                             const double* sigma_dev, double* out_dev, void* stream) {
  if (!audio_dev || !offsets_dev || !lengths_dev || !z_dev || !sigma_dev || !out_dev || n_clips < 0 ||
      dtype < ASR_I16 || dtype > ASR_F64) {
    set_error("asr_mix_white: invalid argument");
    return ASR_ERR_INVALID;
  }
  if (n_clips == 0) return ASR_OK;
  mix_white_kernel<<<dim3(n_clips, 8), 256, 0, as_stream(stream)>>>(
      audio_dev, dtype, reinterpret_cast<const long long*>(offsets_dev), lengths_dev, z_dev, sigma_dev, out_dev);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}

extern "C" int asr_mix_mixture(const void* audio_dev, int32_t dtype, const int64_t* offsets_dev,
                               const int32_t* lengths_dev, int32_t n_clips, const double* q_dev, const double* g_dev,
                               double p, double sigma0, double sigma1, double* out_dev, void* stream) {
  if (!audio_dev || !offsets_dev || !lengths_dev || !q_dev || !g_dev || !out_dev || n_clips < 0 ||
      dtype < ASR_I16 || dtype > ASR_F64) {
    set_error("asr_mix_mixture: invalid argument");
    return ASR_ERR_INVALID;
  }
  if (n_clips == 0) return ASR_OK;
  mix_mixture_kernel<<<dim3(n_clips, 8), 256, 0, as_stream(stream)>>>(
      audio_dev, dtype, reinterpret_cast<const long long*>(offsets_dev), lengths_dev, q_dev, g_dev, p, sigma0, sigma1,
      out_dev);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}

extern "C" int asr_mix_rows_white(const double* x_dev, int64_t n, const double* z_dev, double sigma, double* out_dev,
                                  void* stream) {
  if (!x_dev || !z_dev || !out_dev || n < 0) {
    set_error("asr_mix_rows_white: invalid argument");
    return ASR_ERR_INVALID;
  }
  if (n == 0) return ASR_OK;
  const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 148 * 16));
  mix_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x_dev, n, nullptr, z_dev, 0, 0.0, sigma, sigma, out_dev);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}

extern "C" int asr_mix_rows_mixture(const double* x_dev, int64_t n, const double* q_dev, const double* g_dev, double p,
                                    double sigma0, double sigma1, double* out_dev, void* stream) {
  if (!x_dev || !q_dev || !g_dev || !out_dev || n < 0) {
    set_error("asr_mix_rows_mixture: invalid argument");
    return ASR_ERR_INVALID;
  }
  if (n == 0) return ASR_OK;
  const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 148 * 16));
  mix_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x_dev, n, q_dev, g_dev, 1, p, sigma0, sigma1, out_dev);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}

extern "C" int asr_randn_f64(uint64_t seed, uint64_t first_index, int64_t n, double* out_dev, void* stream) {
  if (!out_dev || n < 0) {
    set_error("asr_randn_f64: invalid argument");
    return ASR_ERR_INVALID;
  }
  if (n == 0) return ASR_OK;
  const int64_t groups = (n + 3) / 4 + 1;
  const int blocks = static_cast<int>(std::min<int64_t>((groups + 255) / 256, 148 * 16));
  randn_kernel<<<blocks, 256, 0, as_stream(stream)>>>(seed, first_index, n, out_dev);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}
