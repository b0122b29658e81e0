// Dataset-level standardisation ("CMVN") for sm_100a.
// Reference: standardize_dataset, VDR/attacks.py:48-69 (sklearn StandardScaler.fit_transform on the
// concatenated train+dev+test rows) and its inline copies in the training scripts.
//
// sklearn's first-batch _incremental_mean_and_var is two passes over the rows:
//   mean = sum(x)/n ;  temp = x - mean ;  corr = sum(temp) ;  ssq = sum(temp^2)
//   var  = (ssq - corr^2/n)/n ;  scale = sqrt(var), 1 where the feature is constant
// Rows may be sharded over GPUs: the accumulators are plain float64 vectors the caller all-reduces.
// Column sums are deterministic: per-slab partials in a workspace, then a fixed-order reduction.
#include "common.cuh"

namespace asr {

constexpr int kColTile = 32;    // columns per block (one per lane: coalesced row reads)
constexpr int kRowLanes = 8;    // row-parallel warps per block
constexpr int kMaxSlabs = 512;

template <int DT>   // ASR_F32 or ASR_F64, a template parameter so that batched loads are not separated by a dtype branch
__device__ __forceinline__ double load_x(const void* __restrict__ x, const long long i) {
  if constexpr (DT == ASR_F32) return static_cast<double>(__ldg(reinterpret_cast<const float*>(x) + i));
  else return __ldg(reinterpret_cast<const double*>(x) + i);
}

// partial[(slab*2 + j)*n_cols + c] ; CENTERED: j=0 sum(x-mean), j=1 sum((x-mean)^2) ; else j=0 sum(x)
template <bool CENTERED, int DT>
__global__ void __launch_bounds__(kColTile * kRowLanes) colsum_partial_kernel(
    const void* __restrict__ x, const long long n_rows, const int n_cols, const long long ld,
    const double* __restrict__ mean, const long long rows_per_slab, double* __restrict__ partial) {
  __shared__ double s0[kRowLanes][kColTile], s1[kRowLanes][kColTile];
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * kColTile + lane;
  const long long r0 = blockIdx.y * rows_per_slab;
  const long long r1 = min(n_rows, r0 + rows_per_slab);
  double a0 = 0.0, a1 = 0.0;
  if (c < n_cols) {
    const double m = CENTERED ? mean[c] : 0.0;
    long long r = r0 + rl;
    for (; r + 7 * kRowLanes < r1; r += 8 * kRowLanes) {     // eight loads in flight, added in the same (ascending) order
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = load_x<DT>(x, (r + u * kRowLanes) * ld + c);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (CENTERED) {
          const double t = __dsub_rn(v[u], m);
          a0 = __dadd_rn(a0, t);
          a1 = __dadd_rn(a1, __dmul_rn(t, t));
        } else {
          a0 = __dadd_rn(a0, v[u]);
        }
      }
    }
    for (; r < r1; r += kRowLanes) {
      const double v = load_x<DT>(x, r * ld + c);
      if (CENTERED) {
        const double t = __dsub_rn(v, m);
        a0 = __dadd_rn(a0, t);
        a1 = __dadd_rn(a1, __dmul_rn(t, t));
      } else {
        a0 = __dadd_rn(a0, v);
      }
    }
  }
  s0[rl][lane] = a0;
  s1[rl][lane] = a1;
  __syncthreads();
  if (rl == 0 && c < n_cols) {
    double t0 = s0[0][lane], t1 = s1[0][lane];
#pragma unroll
    for (int j = 1; j < kRowLanes; ++j) { t0 = __dadd_rn(t0, s0[j][lane]); t1 = __dadd_rn(t1, s1[j][lane]); }
    const long long slab = blockIdx.y;
    partial[(slab * 2 + 0) * n_cols + c] = t0;
    if (CENTERED) partial[(slab * 2 + 1) * n_cols + c] = t1;
  }
}

__global__ void colsum_final_kernel(const double* __restrict__ partial, const int n_slabs, const int n_cols,
                                    const int n_acc, double* __restrict__ acc) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  for (int j = 0; j < n_acc; ++j) {
    double t = 0.0;
    for (int s = 0; s < n_slabs; ++s) t = __dadd_rn(t, partial[(static_cast<long long>(s) * 2 + j) * n_cols + c]);
    acc[static_cast<long long>(j) * n_cols + c] = __dadd_rn(acc[static_cast<long long>(j) * n_cols + c], t);
  }
}

__global__ void cmvn_mean_kernel(const double* __restrict__ acc1, const long long n_total, const int n_cols,
                                 double* __restrict__ mean) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_cols) mean[c] = __ddiv_rn(acc1[c], static_cast<double>(n_total));
}

__global__ void cmvn_finalize_kernel(const double* __restrict__ acc2, const double* __restrict__ mean,
                                     const long long n_total, const int n_cols, double* __restrict__ var,
                                     double* __restrict__ scale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  const double n = static_cast<double>(n_total);
  const double corr = acc2[c], ssq = acc2[n_cols + c];
  const double unnorm = __dsub_rn(ssq, __ddiv_rn(__dmul_rn(corr, corr), n));
  const double v = __ddiv_rn(unnorm, n);
  // sklearn.preprocessing._data._is_constant_feature
  const double eps = 2.220446049250313e-16;
  const double nme = __dmul_rn(__dmul_rn(n, mean[c]), eps);
  const double upper = __dadd_rn(__dmul_rn(__dmul_rn(n, eps), v), __dmul_rn(nme, nme));
  var[c] = v;
  scale[c] = (v <= upper) ? 1.0 : __dsqrt_rn(v);
}

// out[r][c] = (x[r][c] - mean[c]) / scale[c] in float64 (sklearn: X -= mean_; X /= scale_), thread <-> column.
template <int DT>
__global__ void __launch_bounds__(kColTile * kRowLanes) cmvn_apply_kernel(
    const void* __restrict__ x, const long long n_rows, const int n_cols, const long long ld,
    const double* __restrict__ mean, const double* __restrict__ scale, void* __restrict__ out, const int out_f64,
    const long long rows_per_slab) {
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * kColTile + lane;
  if (c >= n_cols) return;
  const long long r0 = blockIdx.y * rows_per_slab;
  const long long r1 = min(n_rows, r0 + rows_per_slab);
  const double m = mean[c], sc = scale[c];
  long long r = r0 + rl;
  for (; r + 3 * kRowLanes < r1; r += 4 * kRowLanes) {
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = load_x<DT>(x, (r + u * kRowLanes) * ld + c);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double y = __ddiv_rn(__dsub_rn(v[u], m), sc);
      const long long o = (r + u * kRowLanes) * n_cols + c;
      if (out_f64) reinterpret_cast<double*>(out)[o] = y;
      else reinterpret_cast<float*>(out)[o] = static_cast<float>(y);
    }
  }
  for (; r < r1; r += kRowLanes) {
    const double y = __ddiv_rn(__dsub_rn(load_x<DT>(x, r * ld + c), m), sc);
    if (out_f64) reinterpret_cast<double*>(out)[r * n_cols + c] = y;
    else reinterpret_cast<float*>(out)[r * n_cols + c] = static_cast<float>(y);
  }
}

template <bool CENTERED>
static int run_colsum(const void* x, int dtype, int64_t n_rows, int n_cols, int64_t ld, const double* mean,
                      double* acc, cudaStream_t st) {
  const int col_blocks = (n_cols + kColTile - 1) / kColTile;
  // enough slabs to fill 148 SMs a few times over, each at least 64 rows
  int64_t slabs = (148 * 8 + col_blocks - 1) / col_blocks;
  slabs = std::max<int64_t>(1, std::min<int64_t>(slabs, std::min<int64_t>(kMaxSlabs, (n_rows + 63) / 64)));
  const int64_t rows_per_slab = (n_rows + slabs - 1) / slabs;
  slabs = (n_rows + rows_per_slab - 1) / rows_per_slab;
  double* partial = nullptr;
  ASR_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&partial), sizeof(double) * 2 * slabs * n_cols, st));
  const dim3 grid(col_blocks, static_cast<unsigned>(slabs));
  if (dtype == ASR_F32)
    colsum_partial_kernel<CENTERED, ASR_F32><<<grid, kColTile * kRowLanes, 0, st>>>(x, n_rows, n_cols, ld, mean, rows_per_slab, partial);
  else
    colsum_partial_kernel<CENTERED, ASR_F64><<<grid, kColTile * kRowLanes, 0, st>>>(x, n_rows, n_cols, ld, mean, rows_per_slab, partial);
  ASR_CUDA_TRY(cudaGetLastError());
  colsum_final_kernel<<<(n_cols + 127) / 128, 128, 0, st>>>(partial, static_cast<int>(slabs), n_cols,
                                                           CENTERED ? 2 : 1, acc);
  ASR_CUDA_TRY(cudaGetLastError());
  ASR_CUDA_TRY(cudaFreeAsync(partial, st));
  return ASR_OK;
}

}  // namespace asr

using namespace asr;
static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static bool bad_matrix(const void* x, int32_t dtype, int64_t n_rows, int32_t n_cols, int64_t ld, const char* who) {
  if ((!x && n_rows != 0) || n_rows < 0 || n_cols <= 0 || ld < n_cols || (dtype != ASR_F32 && dtype != ASR_F64)) {
    set_error(std::string(who) + ": invalid matrix argument (dtype must be ASR_F32 or ASR_F64, ld >= n_cols)");
    return true;
  }
  return false;
}

extern "C" int asr_cmvn_colsum(const void* x_dev, int32_t dtype, int64_t n_rows, int32_t n_cols, int64_t ld,
                               double* acc1_dev, void* stream) {
  if (bad_matrix(x_dev, dtype, n_rows, n_cols, ld, "asr_cmvn_colsum") || !acc1_dev) {
    if (!acc1_dev) set_error("asr_cmvn_colsum: null accumulator");
    return ASR_ERR_INVALID;
  }
  if (n_rows == 0) return ASR_OK;
  return run_colsum<false>(x_dev, dtype, n_rows, n_cols, ld, nullptr, acc1_dev, as_stream(stream));
}

extern "C" int asr_cmvn_mean(const double* acc1_dev, int64_t n_total, int32_t n_cols, double* mean_dev, void* stream) {
  if (!acc1_dev || !mean_dev || n_total <= 0 || n_cols <= 0) {
    set_error("asr_cmvn_mean: invalid argument");
    return ASR_ERR_INVALID;
  }
  cmvn_mean_kernel<<<(n_cols + 127) / 128, 128, 0, as_stream(stream)>>>(acc1_dev, n_total, n_cols, mean_dev);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}

extern "C" int asr_cmvn_colsum_centered(const void* x_dev, int32_t dtype, int64_t n_rows, int32_t n_cols, int64_t ld,
                                        const double* mean_dev, double* acc2_dev, void* stream) {
  if (bad_matrix(x_dev, dtype, n_rows, n_cols, ld, "asr_cmvn_colsum_centered") || !mean_dev || !acc2_dev) {
    if (!mean_dev || !acc2_dev) set_error("asr_cmvn_colsum_centered: null pointer");
    return ASR_ERR_INVALID;
  }
  if (n_rows == 0) return ASR_OK;
  return run_colsum<true>(x_dev, dtype, n_rows, n_cols, ld, mean_dev, acc2_dev, as_stream(stream));
}

extern "C" int asr_cmvn_finalize(const double* acc2_dev, const double* mean_dev, int64_t n_total, int32_t n_cols,
                                 double* var_dev, double* scale_dev, void* stream) {
  if (!acc2_dev || !mean_dev || !var_dev || !scale_dev || n_total <= 0 || n_cols <= 0) {
    set_error("asr_cmvn_finalize: invalid argument");
    return ASR_ERR_INVALID;
  }
  cmvn_finalize_kernel<<<(n_cols + 127) / 128, 128, 0, as_stream(stream)>>>(acc2_dev, mean_dev, n_total, n_cols,
                                                                           var_dev, scale_dev);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}

extern "C" int asr_cmvn_apply(const void* x_dev, int32_t dtype, int64_t n_rows, int32_t n_cols, int64_t ld,
                              const double* mean_dev, const double* scale_dev, void* out_dev, int32_t out_dtype,
                              void* stream) {
  if (bad_matrix(x_dev, dtype, n_rows, n_cols, ld, "asr_cmvn_apply") || !mean_dev || !scale_dev || !out_dev ||
      (out_dtype != ASR_F32 && out_dtype != ASR_F64)) {
    if (!mean_dev || !scale_dev || !out_dev) set_error("asr_cmvn_apply: null pointer");
    else if (out_dtype != ASR_F32 && out_dtype != ASR_F64) set_error("asr_cmvn_apply: out_dtype must be ASR_F32/ASR_F64");
    return ASR_ERR_INVALID;
  }
  if (n_rows == 0) return ASR_OK;
  const int col_blocks = (n_cols + kColTile - 1) / kColTile;
  int64_t slabs = std::max<int64_t>(1, std::min<int64_t>((148 * 16 + col_blocks - 1) / col_blocks, (n_rows + 31) / 32));
  const int64_t rows_per_slab = (n_rows + slabs - 1) / slabs;
  slabs = (n_rows + rows_per_slab - 1) / rows_per_slab;
  const dim3 grid(col_blocks, static_cast<unsigned>(slabs));
  if (dtype == ASR_F32)
    cmvn_apply_kernel<ASR_F32><<<grid, kColTile * kRowLanes, 0, as_stream(stream)>>>(
        x_dev, n_rows, n_cols, ld, mean_dev, scale_dev, out_dev, out_dtype == ASR_F64, rows_per_slab);
  else
    cmvn_apply_kernel<ASR_F64><<<grid, kColTile * kRowLanes, 0, as_stream(stream)>>>(
        x_dev, n_rows, n_cols, ld, mean_dev, scale_dev, out_dev, out_dtype == ASR_F64, rows_per_slab);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}
