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


// ================================================================================================
// Fused form (round 2): the same two passes and the same summation order, in three launches on one GPU and with ONE
// exchange when the rows are sharded.
//   pass 1   slab partials of sum(x)                                   colsum2_kernel<false>
//   pass 2   slab partials of sum(x - m), sum((x - m)^2), m = LOCAL mean: every CTA first finishes pass 1 for its 32 columns
//            (fixed slab order, the same value in every CTA)           colsum2_kernel<true>
//   apply    single rank: every CTA finishes pass 2 for its columns, derives var / scale (sklearn's formulas) and
//            standardises its rows; CTAs of the first row slab also store mean / var / scale      apply2_kernel<.., true>
//   sharded  local_message_kernel -> [n, S, C, Q] per rank -> ONE all-gather -> merge_kernel (fixed rank order, Chan's
//            update of the centred sums to the global mean) -> apply2_kernel<.., false>.
// With one rank the merge is the identity, so both forms give the statistics of the classic entry points bit for bit.
// Row noise (VDR/attacks.py:186-219, add_white_noise_on_dataset / add_noise_mixture_on_dataset followed by
// standardize_dataset, :433-491): when a noise descriptor is given every kernel reads x + sel*g (float64, two roundings)
// instead of x - the noisy matrix is never written.
struct RowNoise {
  const double* q;      // mixture selector stream (or null)
  const double* g;      // carrier stream (white: the standard-normal stream), [n_rows][n_cols] contiguous; null = no noise
  double p, s0, s1;     // white: s0 = sigma
};

template <int DT, bool NOISE>   // NOISE is a template parameter so that the batched loads of the clean form are not separated by a branch
__device__ __forceinline__ double load_xn(const void* __restrict__ x, const long long r, const int c, const long long ld,
                                          const int n_cols, const RowNoise& nz) {
  const double v = load_x<DT>(x, r * ld + c);
  if constexpr (!NOISE) return v;
  const long long i = r * n_cols + c;
  const double sel = (nz.q != nullptr && fabs(__ldg(nz.q + i)) < nz.p) ? nz.s1 : nz.s0;
  return __dadd_rn(v, __dmul_rn(sel, __ldg(nz.g + i)));
}

// fixed-order sum over the slabs of a partial table [n_slabs][rows_per][n_cols] at row `j` (loads eight at a time in
// flight, added in ascending slab order)
__device__ __forceinline__ double slab_sum(const double* __restrict__ part, const int n_slabs, const int rows_per, const int j,
                                           const int n_cols, const int c) {
  double t = 0.0;
  const long long step = static_cast<long long>(rows_per) * n_cols;
  const double* p = part + static_cast<long long>(j) * n_cols + c;
  int s = 0;
  for (; s + 8 <= n_slabs; s += 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcg(p + (s + u) * step);
#pragma unroll
    for (int u = 0; u < 8; ++u) t = __dadd_rn(t, v[u]);
  }
  for (; s < n_slabs; ++s) t = __dadd_rn(t, __ldcg(p + s * step));
  return t;
}

template <bool CENTERED, int DT, bool NOISE>
__global__ void __launch_bounds__(kColTile * kRowLanes, NOISE ? 1 : 8) colsum2_kernel(   // clean: 8 CTAs per SM (32 registers), the grid of slab_shape() is ONE wave
    const void* __restrict__ x, const long long n_rows, const int n_cols, const long long ld, const RowNoise nz,
    const double* __restrict__ part1, const int n_slabs1, const double n_local, const long long rows_per_slab,
    double* __restrict__ out_part, const int slab_base) {
  __shared__ double s0[kRowLanes][kColTile], s1[kRowLanes][kColTile];
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * kColTile + lane;
  const long long r0 = blockIdx.y * rows_per_slab;
  const long long r1 = min(n_rows, r0 + rows_per_slab);
  double a0 = 0.0, a1 = 0.0;
  if (CENTERED) {
    if (rl == 0 && c < n_cols) s0[0][lane] = __ddiv_rn(slab_sum(part1, n_slabs1, 1, 0, n_cols, c), n_local);
    __syncthreads();
  }
  const double m = CENTERED ? s0[0][lane] : 0.0;
  if (CENTERED) __syncthreads();                          // s0 is reused for the partial sums below
  if (c < n_cols) {
    long long r = r0 + rl;
    for (; r + 7 * kRowLanes < r1; r += 8 * kRowLanes) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = load_xn<DT, NOISE>(x, r + u * kRowLanes, c, ld, n_cols, nz);
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
      const double v = load_xn<DT, NOISE>(x, r, c, ld, n_cols, nz);
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
    const long long slab = slab_base + blockIdx.y;
    if (CENTERED) {
      out_part[(slab * 2 + 0) * n_cols + c] = t0;
      out_part[(slab * 2 + 1) * n_cols + c] = t1;
    } else {
      out_part[slab * n_cols + c] = t0;
    }
  }
}

// var / scale from the centred sums (sklearn: _incremental_mean_and_var first batch + _is_constant_feature)
__device__ __forceinline__ void finish_stats(const double corr, const double ssq, const double n, const double mean, double& var,
                                             double& scale) {
  const double unnorm = __dsub_rn(ssq, __ddiv_rn(__dmul_rn(corr, corr), n));
  const double v = __ddiv_rn(unnorm, n);
  const double eps = 2.220446049250313e-16;
  const double nme = __dmul_rn(__dmul_rn(n, mean), eps);
  const double upper = __dadd_rn(__dmul_rn(__dmul_rn(n, eps), v), __dmul_rn(nme, nme));
  var = v;
  scale = (v <= upper) ? 1.0 : __dsqrt_rn(v);
}

// this rank's message [n, S (D), C (D), Q (D)]
__global__ void local_message_kernel(const double* __restrict__ part1, const int n_slabs1, const double* __restrict__ part2,
                                     const int n_slabs2, const double n_local, const int n_cols, double* __restrict__ msg) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0) msg[0] = n_local;
  if (c >= n_cols) return;
  msg[1 + c] = slab_sum(part1, n_slabs1, 1, 0, n_cols, c);
  msg[1 + n_cols + c] = n_slabs2 > 0 ? slab_sum(part2, n_slabs2, 2, 0, n_cols, c) : 0.0;
  msg[1 + 2 * n_cols + c] = n_slabs2 > 0 ? slab_sum(part2, n_slabs2, 2, 1, n_cols, c) : 0.0;
}

// messages of all ranks, in rank order -> mean / var / scale of the whole dataset (identical on every rank)
__global__ void merge_kernel(const double* __restrict__ msgs, const int world, const int n_cols, double* __restrict__ mean,
                             double* __restrict__ var, double* __restrict__ scale, double* __restrict__ n_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = 1 + 3LL * n_cols;
  double n = 0.0;
  for (int r = 0; r < world; ++r) n = __dadd_rn(n, msgs[r * stride]);
  if (c == 0 && n_out) n_out[0] = n;
  if (c >= n_cols || n <= 0.0) return;
  double S = 0.0;
  for (int r = 0; r < world; ++r) S = __dadd_rn(S, msgs[r * stride + 1 + c]);
  const double mu = __ddiv_rn(S, n);
  double corr = 0.0, ssq = 0.0;
  for (int r = 0; r < world; ++r) {
    const double nr = msgs[r * stride];
    if (nr <= 0.0) continue;
    const double mr = __ddiv_rn(msgs[r * stride + 1 + c], nr);
    const double C = msgs[r * stride + 1 + n_cols + c], Q = msgs[r * stride + 1 + 2 * n_cols + c];
    const double d = __dsub_rn(mr, mu);                       // sum(x - mu) = C + n_r d ; sum((x - mu)^2) = Q + 2 d C + n_r d^2
    corr = __dadd_rn(corr, __dadd_rn(C, __dmul_rn(nr, d)));
    ssq = __dadd_rn(ssq, __dadd_rn(__dadd_rn(Q, __dmul_rn(__dmul_rn(2.0, d), C)), __dmul_rn(nr, __dmul_rn(d, d))));
  }
  double v, sc;
  finish_stats(corr, ssq, n, mu, v, sc);
  mean[c] = mu; var[c] = v; scale[c] = sc;
}

// out = (x [+ noise] - mean) / scale.  FUSED (single rank): mean / scale are finished here from the slab partials.
template <int DT, bool FUSED, bool NOISE>
__global__ void __launch_bounds__(kColTile * kRowLanes) apply2_kernel(
    const void* __restrict__ x, const long long n_rows, const int n_cols, const long long ld, const RowNoise nz,
    const double* __restrict__ part1, const int n_slabs1, const double* __restrict__ part2, const int n_slabs2, const double n_total,
    double* __restrict__ mean, double* __restrict__ var, double* __restrict__ scale, void* __restrict__ out, const int out_f64,
    const long long rows_per_slab) {
  __shared__ double s_m[kColTile], s_s[kColTile];
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * kColTile + lane;
  if (FUSED) {
    __shared__ double s_t[3][kColTile];
    if (rl < 3 && c < n_cols)                             // the three slab reductions side by side
      s_t[rl][lane] = rl == 0 ? slab_sum(part1, n_slabs1, 1, 0, n_cols, c) : slab_sum(part2, n_slabs2, 2, rl - 1, n_cols, c);
    __syncthreads();
    if (rl == 0 && c < n_cols) {
      const double mu = __ddiv_rn(s_t[0][lane], n_total);
      double v, sc;
      finish_stats(s_t[1][lane], s_t[2][lane], n_total, mu, v, sc);
      s_m[lane] = mu; s_s[lane] = sc;
      if (blockIdx.y == 0) { mean[c] = mu; var[c] = v; scale[c] = sc; }
    }
    __syncthreads();
  }
  if (c >= n_cols) return;
  const long long r0 = blockIdx.y * rows_per_slab;
  const long long r1 = min(n_rows, r0 + rows_per_slab);
  const double m = FUSED ? s_m[lane] : mean[c], sc = FUSED ? s_s[lane] : scale[c];
  long long r = r0 + rl;
  for (; r + 3 * kRowLanes < r1; r += 4 * kRowLanes) {
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = load_xn<DT, NOISE>(x, r + u * kRowLanes, c, ld, n_cols, nz);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double y = __ddiv_rn(__dsub_rn(v[u], m), sc);
      const long long o = (r + u * kRowLanes) * n_cols + c;
      if (out_f64) reinterpret_cast<double*>(out)[o] = y;
      else reinterpret_cast<float*>(out)[o] = static_cast<float>(y);
    }
  }
  for (; r < r1; r += kRowLanes) {
    const double y = __ddiv_rn(__dsub_rn(load_xn<DT, NOISE>(x, r, c, ld, n_cols, nz), m), sc);
    if (out_f64) reinterpret_cast<double*>(out)[r * n_cols + c] = y;
    else reinterpret_cast<float*>(out)[r * n_cols + c] = static_cast<float>(y);
  }
}

// ------------------------------------------------------------------------------------------------
// The path's one exchange as a one-shot all-gather over PEER MEMORY (NVLink / NVSwitch), no NCCL call: every rank owns a
// symmetric region [2][world][M] float64 receive slots (double-buffered by step parity) + world uint32 flags, mapped into every
// peer.  CTA p of rank r: store r's message into peer p's slot [parity][r] (remote stores), fence, raise p's flag[r] to the step
// number; wait until OUR flag[p] has reached the step number; copy slot [parity][p] of our region into the local msgs[p].
// Step numbers are per-peer counters in local memory (state[p]), so the kernel takes no per-step argument and can sit inside
// a CUDA graph.  A slot of parity e & 1 is rewritten at step e + 2, which the writer reaches only after it has seen our flag of
// step e + 1, i.e. after our step-e kernel (which read the slot) has finished.
__global__ void __launch_bounds__(256) exchange_p2p_kernel(const double* __restrict__ msg, const int M, const int rank, const int world,
                                                           void* const* __restrict__ regions, unsigned* __restrict__ state,
                                                           double* __restrict__ msgs, const long long flags_off) {
  const int p = blockIdx.x, tid = threadIdx.x;
  const unsigned e = state[p] + 1u;
  const unsigned par = e & 1u;
  char* const peer = static_cast<char*>(regions[p]);
  char* const own = static_cast<char*>(regions[rank]);
  double* const dst = reinterpret_cast<double*>(peer) + (static_cast<size_t>(par) * world + rank) * M;
  for (int i = tid; i < M; i += blockDim.x) dst[i] = msg[i];
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();                               // the message before the flag, system-wide
    *reinterpret_cast<volatile unsigned*>(peer + flags_off + 4 * rank) = e;
    const volatile unsigned* mine = reinterpret_cast<const volatile unsigned*>(own + flags_off + 4 * p);
    unsigned long long spins = 0;
    while (static_cast<int>(*mine - e) < 0) {             // peer p's message of this step has not landed yet
      __nanosleep(64);
      if (++spins > (1ull << 28)) __trap();               // ~20 s: a peer that never arrives fails the launch instead of hanging
    }
    __threadfence_system();
  }
  __syncthreads();
  const double* src = reinterpret_cast<const double*>(own) + (static_cast<size_t>(par) * world + p) * M;
  for (int i = tid; i < M; i += blockDim.x) msgs[static_cast<size_t>(p) * M + i] = __ldcv(src + i);   // written by a peer: not through L1
  __syncthreads();
  if (tid == 0) state[p] = e;
}

static void slab_shape(int64_t n_rows, int n_cols, int64_t* slabs, int64_t* rows_per_slab) {
  const int col_blocks = (n_cols + kColTile - 1) / kColTile;
  int64_t s = (148 * 8 + col_blocks - 1) / col_blocks;
  s = std::max<int64_t>(1, std::min<int64_t>(s, std::min<int64_t>(kMaxSlabs / 4, (n_rows + 63) / 64)));
  *rows_per_slab = (n_rows + s - 1) / s;
  *slabs = (n_rows + *rows_per_slab - 1) / *rows_per_slab;
}
static RowNoise row_noise(const asr_noise* nz) {
  RowNoise r{nullptr, nullptr, 0.0, 0.0, 0.0};
  if (nz && nz->mode == ASR_NOISE_WHITE) { r.g = nz->z_dev; r.s0 = nz->sigma0; r.s1 = nz->sigma0; }
  else if (nz && nz->mode == ASR_NOISE_MIXTURE) { r.q = nz->z_dev; r.g = nz->z2_dev; r.p = nz->p; r.s0 = nz->sigma0; r.s1 = nz->sigma1; }
  return r;
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

// ---- fused form -----------------------------------------------------------------------------------------------------
// Workspace: [kMaxSlabs][D] pass-1 partials, then [kMaxSlabs][2][D] pass-2 partials.
extern "C" size_t asr_cmvn_workspace_bytes(int32_t n_cols) {
  return n_cols > 0 ? sizeof(double) * 3 * static_cast<size_t>(kMaxSlabs) * n_cols : 0;
}

static bool bad_ws(const void* ws, size_t bytes, int32_t n_cols, const char* who) {
  if (!ws || bytes < asr_cmvn_workspace_bytes(n_cols) || (reinterpret_cast<uintptr_t>(ws) & 7)) {
    set_error(std::string(who) + ": workspace missing, misaligned or smaller than asr_cmvn_workspace_bytes(n_cols)");
    return true;
  }
  return false;
}
static bool bad_row_noise(const asr_noise* nz, const char* who) {
  if (!nz || nz->mode == ASR_NOISE_NONE) return false;
  if ((nz->mode == ASR_NOISE_WHITE && nz->z_dev) || (nz->mode == ASR_NOISE_MIXTURE && nz->z_dev && nz->z2_dev)) return false;
  set_error(std::string(who) + ": row noise needs z_dev (white: the carrier, sigma in sigma0) or z_dev + z2_dev (mixture)");
  return true;
}

extern "C" int32_t asr_cmvn_partial_sums(const void* x_dev, int32_t dtype, int64_t n_rows, int32_t n_cols, int64_t ld,
                                         const asr_noise* row_noise_desc, int32_t pass, int32_t n_slabs_pass1,
                                         int64_t n_local_rows, int32_t slab_base, void* workspace_dev, size_t workspace_bytes,
                                         void* stream) {
  if (bad_matrix(x_dev, dtype, n_rows, n_cols, ld, "asr_cmvn_partial_sums") || bad_ws(workspace_dev, workspace_bytes, n_cols, "asr_cmvn_partial_sums") ||
      bad_row_noise(row_noise_desc, "asr_cmvn_partial_sums"))
    return ASR_ERR_INVALID;
  if ((pass != 1 && pass != 2) || slab_base < 0 || (pass == 2 && (n_slabs_pass1 < 1 || n_local_rows < 1))) {
    set_error("asr_cmvn_partial_sums: pass must be 1 or 2 (pass 2 needs the pass-1 slab count and the local row count)");
    return ASR_ERR_INVALID;
  }
  if (n_rows == 0) return 0;
  int64_t slabs, rps;
  slab_shape(n_rows, n_cols, &slabs, &rps);
  if (slab_base + slabs > kMaxSlabs) { set_error("asr_cmvn_partial_sums: too many row blocks for the workspace"); return ASR_ERR_TOO_LARGE; }
  double* part1 = static_cast<double*>(workspace_dev);
  double* part2 = part1 + static_cast<size_t>(kMaxSlabs) * n_cols;
  const RowNoise nz = row_noise(row_noise_desc);
  const dim3 grid((n_cols + kColTile - 1) / kColTile, static_cast<unsigned>(slabs));
  cudaStream_t st = as_stream(stream);
  const double nl = static_cast<double>(n_local_rows);
  const bool noisy = nz.g != nullptr;
#define ASR_COLSUM2(C, DT, N) colsum2_kernel<C, DT, N><<<grid, kColTile * kRowLanes, 0, st>>>(x_dev, n_rows, n_cols, ld, nz, part1, n_slabs_pass1, nl, rps, (C) ? part2 : part1, slab_base)
  if (pass == 1) {
    if (dtype == ASR_F32) { if (noisy) ASR_COLSUM2(false, ASR_F32, true); else ASR_COLSUM2(false, ASR_F32, false); }
    else { if (noisy) ASR_COLSUM2(false, ASR_F64, true); else ASR_COLSUM2(false, ASR_F64, false); }
  } else {
    if (dtype == ASR_F32) { if (noisy) ASR_COLSUM2(true, ASR_F32, true); else ASR_COLSUM2(true, ASR_F32, false); }
    else { if (noisy) ASR_COLSUM2(true, ASR_F64, true); else ASR_COLSUM2(true, ASR_F64, false); }
  }
#undef ASR_COLSUM2
  ASR_CUDA_TRY(cudaGetLastError());
  return static_cast<int32_t>(slabs);
}

extern "C" int asr_cmvn_local_message(const void* workspace_dev, size_t workspace_bytes, int32_t n_slabs_pass1, int32_t n_slabs_pass2,
                                      int64_t n_local_rows, int32_t n_cols, double* msg_dev, void* stream) {
  if (bad_ws(workspace_dev, workspace_bytes, n_cols, "asr_cmvn_local_message") || !msg_dev || n_slabs_pass1 < 0 || n_slabs_pass2 < 0 || n_local_rows < 0) {
    if (!msg_dev) set_error("asr_cmvn_local_message: null message");
    return ASR_ERR_INVALID;
  }
  const double* part1 = static_cast<const double*>(workspace_dev);
  const double* part2 = part1 + static_cast<size_t>(kMaxSlabs) * n_cols;
  local_message_kernel<<<(n_cols + 127) / 128, 128, 0, as_stream(stream)>>>(part1, n_slabs_pass1, part2, n_slabs_pass2,
                                                                            static_cast<double>(n_local_rows), n_cols, msg_dev);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}

extern "C" int asr_cmvn_merge(const double* msgs_dev, int32_t world, int32_t n_cols, double* mean_dev, double* var_dev,
                              double* scale_dev, double* n_total_dev, void* stream) {
  if (!msgs_dev || world < 1 || n_cols <= 0 || !mean_dev || !var_dev || !scale_dev) {
    set_error("asr_cmvn_merge: invalid argument");
    return ASR_ERR_INVALID;
  }
  merge_kernel<<<(n_cols + 127) / 128, 128, 0, as_stream(stream)>>>(msgs_dev, world, n_cols, mean_dev, var_dev, scale_dev, n_total_dev);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}

extern "C" int asr_cmvn_apply2(const void* x_dev, int32_t dtype, int64_t n_rows, int32_t n_cols, int64_t ld,
                               const asr_noise* row_noise_desc, const void* workspace_dev, size_t workspace_bytes,
                               int32_t n_slabs_pass1, int32_t n_slabs_pass2, int64_t n_total_rows, double* mean_dev,
                               double* var_dev, double* scale_dev, void* out_dev, int32_t out_dtype, void* stream) {
  if (bad_matrix(x_dev, dtype, n_rows, n_cols, ld, "asr_cmvn_apply2") || bad_row_noise(row_noise_desc, "asr_cmvn_apply2") || !mean_dev ||
      !var_dev || !scale_dev || !out_dev || (out_dtype != ASR_F32 && out_dtype != ASR_F64)) {
    if (!mean_dev || !var_dev || !scale_dev || !out_dev) set_error("asr_cmvn_apply2: null pointer");
    else if (out_dtype != ASR_F32 && out_dtype != ASR_F64) set_error("asr_cmvn_apply2: out_dtype must be ASR_F32/ASR_F64");
    return ASR_ERR_INVALID;
  }
  const bool fused = workspace_dev != nullptr;             // single rank: the statistics are finished inside this launch
  if (fused && (bad_ws(workspace_dev, workspace_bytes, n_cols, "asr_cmvn_apply2") || n_slabs_pass1 < 1 || n_slabs_pass2 < 1 || n_total_rows < 1)) {
    if (n_slabs_pass1 < 1 || n_slabs_pass2 < 1 || n_total_rows < 1) set_error("asr_cmvn_apply2: fused form needs both slab counts and the row count");
    return ASR_ERR_INVALID;
  }
  if (n_rows == 0) return ASR_OK;
  const int col_blocks = (n_cols + kColTile - 1) / kColTile;
  // apply2_kernel holds 4 CTAs per SM (62 registers for the float64 division): at most 148 x 4 CTAs = ONE wave
  int64_t slabs = std::max<int64_t>(1, std::min<int64_t>((148 * 4) / col_blocks, (n_rows + 31) / 32));
  const int64_t rps = (n_rows + slabs - 1) / slabs;
  slabs = (n_rows + rps - 1) / rps;
  const dim3 grid(col_blocks, static_cast<unsigned>(slabs));
  const double* part1 = static_cast<const double*>(workspace_dev);
  const double* part2 = fused ? part1 + static_cast<size_t>(kMaxSlabs) * n_cols : nullptr;
  const RowNoise nz = row_noise(row_noise_desc);
  cudaStream_t st = as_stream(stream);
  const double nt = static_cast<double>(n_total_rows);
  const int f64 = out_dtype == ASR_F64;
  const bool noisy = nz.g != nullptr;
#define ASR_APPLY2(DT, F, N) apply2_kernel<DT, F, N><<<grid, kColTile * kRowLanes, 0, st>>>(x_dev, n_rows, n_cols, ld, nz, part1, n_slabs_pass1, part2, n_slabs_pass2, nt, mean_dev, var_dev, scale_dev, out_dev, f64, rps)
#define ASR_APPLY2_N(DT, F) do { if (noisy) ASR_APPLY2(DT, F, true); else ASR_APPLY2(DT, F, false); } while (0)
  if (dtype == ASR_F32) { if (fused) ASR_APPLY2_N(ASR_F32, true); else ASR_APPLY2_N(ASR_F32, false); }
  else { if (fused) ASR_APPLY2_N(ASR_F64, true); else ASR_APPLY2_N(ASR_F64, false); }
#undef ASR_APPLY2_N
#undef ASR_APPLY2
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}

// Region a rank exports to its peers for asr_cmvn_exchange_p2p: [2][world][3*n_cols+1] float64 + world uint32 flags.
extern "C" size_t asr_cmvn_p2p_region_bytes(int32_t world, int32_t n_cols) {
  if (world < 1 || n_cols < 1) return 0;
  const size_t m = 3 * static_cast<size_t>(n_cols) + 1;
  return ((2 * static_cast<size_t>(world) * m * sizeof(double) + 4 * static_cast<size_t>(world)) + 255) & ~static_cast<size_t>(255);
}

extern "C" int asr_cmvn_exchange_p2p(const double* msg_dev, int32_t n_cols, int32_t rank, int32_t world, void* const* peer_regions_dev,
                                     uint32_t* state_dev, double* msgs_dev, void* stream) {
  if (!msg_dev || !peer_regions_dev || !state_dev || !msgs_dev || n_cols < 1 || world < 1 || rank < 0 || rank >= world) {
    set_error("asr_cmvn_exchange_p2p: invalid argument");
    return ASR_ERR_INVALID;
  }
  const int M = 3 * n_cols + 1;
  const long long flags_off = 2LL * world * M * static_cast<long long>(sizeof(double));
  exchange_p2p_kernel<<<world, 256, 0, as_stream(stream)>>>(msg_dev, M, rank, world, peer_regions_dev, state_dev, msgs_dev, flags_off);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}
