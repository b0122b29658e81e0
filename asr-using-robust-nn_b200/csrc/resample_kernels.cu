// Audio ingest for sm_100a: polyphase resampling of a packed batch of clips.
//
// The reference decodes every file with librosa.load(path, mono=True) (VDR/extract_features_construct_dataset.py:27,
// SR/...:210): libsndfile -> float32 -> resample to 22 050 Hz with resampy's `kaiser_best` filter, whose table is not
// in the image.  The stated stand-in is scipy.signal.resample_poly(x, up, down): Kaiser(5.0)-windowed
// sinc of 20*max(up,down)+1 taps designed in float64, cast to the dtype of x, scaled by `up`, applied by upfirdn in
// float32 with the taps of a phase taken in ascending input order.  asr_resample_design restates the design,
// resample_kernel the filtering (one thread per output sample, same summation order, separate multiply and add).
#include <cmath>
#include <numeric>
#include "common.cuh"

namespace asr {

// modified Bessel function I0 by its power series (double; terms fall below 1e-17 relative within ~30 steps for beta <= 20)
static double bessel_i0(double x) {
  const double q = 0.25 * x * x;
  double term = 1.0, sum = 1.0;
  for (int k = 1; k < 200; ++k) {
    term *= q / (static_cast<double>(k) * k);
    sum += term;
    if (term < 1e-18 * sum) break;
  }
  return sum;
}

template <int DT>
__global__ void __launch_bounds__(256) resample_kernel(const void* __restrict__ in, const long long* __restrict__ in_off,
                                                       const int* __restrict__ in_len, const int up, const int down,
                                                       const float* __restrict__ h, const int n_taps, const int n_pre_remove,
                                                       float* __restrict__ out, const long long* __restrict__ out_off) {
  const int b = blockIdx.y;
  const int n_in = in_len[b];
  if (n_in <= 0) return;
  const long long prod = static_cast<long long>(n_in) * up;
  const long long n_out = prod / down + (prod % down != 0);
  const long long ib = in_off[b], ob = out_off[b];
  for (long long m = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; m < n_out;
       m += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long t = (m + n_pre_remove) * down;       // position in the up-sampled signal
    const long long i_hi = t / up;                       // newest input sample under the filter
    const int k0 = static_cast<int>(t - i_hi * up);      // its tap
    // taps k = k0 + j*up < n_taps meet inputs i = i_hi - j; ascending i = descending j (scipy's upfirdn order)
    int j_hi = (n_taps - 1 - k0) / up;
    if (k0 >= n_taps) j_hi = -1;
    long long j_lo = i_hi - (n_in - 1);                  // i <= n_in - 1
    if (j_lo < 0) j_lo = 0;
    if (j_hi > i_hi) j_hi = static_cast<int>(i_hi);      // i >= 0
    float acc = 0.0f;
    for (long long j = j_hi; j >= j_lo; --j) {
      const long long i = i_hi - j;
      float x;
      if (DT == ASR_I16) x = static_cast<float>(__ldg(reinterpret_cast<const short*>(in) + ib + i)) * (1.0f / 32768.0f);
      else x = __ldg(reinterpret_cast<const float*>(in) + ib + i);
      acc = __fadd_rn(acc, __fmul_rn(x, __ldg(h + k0 + j * up)));
    }
    out[ob + m] = acc;
  }
}

}  // namespace asr

using namespace asr;

extern "C" int64_t asr_resample_out_len(int64_t n_in, int32_t up, int32_t down) {
  if (n_in <= 0 || up < 1 || down < 1) return 0;
  const int g = std::gcd(up, down);
  const int64_t prod = n_in * (up / g);
  return prod / (down / g) + (prod % (down / g) != 0);
}

extern "C" int asr_resample_design(int32_t up, int32_t down, double kaiser_beta, float* taps_host, int32_t capacity,
                                   int32_t* up_out, int32_t* down_out, int32_t* n_taps_out, int32_t* n_pre_remove_out) {
  if (up < 1 || down < 1 || !up_out || !down_out || !n_taps_out || !n_pre_remove_out) {
    set_error("asr_resample_design: invalid argument");
    return ASR_ERR_INVALID;
  }
  const int g = std::gcd(up, down);
  up /= g; down /= g;
  *up_out = up; *down_out = down;
  if (up == 1 && down == 1) { *n_taps_out = 0; *n_pre_remove_out = 0; return ASR_OK; }
  const int max_rate = std::max(up, down);
  const double f_c = 1.0 / max_rate;
  const int half_len = 10 * max_rate;
  const int numtaps = 2 * half_len + 1;
  const int n_pre_pad = down - half_len % down;
  *n_taps_out = n_pre_pad + numtaps;
  *n_pre_remove_out = (half_len + n_pre_pad) / down;
  if (!taps_host) return ASR_OK;                           // size query
  if (capacity < *n_taps_out) { set_error("asr_resample_design: taps buffer too small"); return ASR_ERR_INVALID; }
  // scipy.signal.firwin(numtaps, f_c, window=('kaiser', beta)): cutoff*sinc(cutoff*m) * kaiser, unit gain at DC
  const double kPi = 3.141592653589793238462643383279502884;
  const double alpha = 0.5 * (numtaps - 1);
  std::vector<double> h(numtaps);
  const double i0b = bessel_i0(kaiser_beta);
  double s = 0.0;
  for (int n = 0; n < numtaps; ++n) {
    const double m = n - alpha;
    const double xs = f_c * m;
    const double sinc = xs == 0.0 ? 1.0 : std::sin(kPi * xs) / (kPi * xs);
    const double r = m / alpha;
    const double w = bessel_i0(kaiser_beta * std::sqrt(std::max(0.0, 1.0 - r * r))) / i0b;
    h[n] = f_c * sinc * w;
    s += h[n];
  }
  for (int k = 0; k < n_pre_pad; ++k) taps_host[k] = 0.0f;
  for (int n = 0; n < numtaps; ++n) {
    const float hf = static_cast<float>(h[n] / s);          // h.astype(x.dtype) ...
    taps_host[n_pre_pad + n] = hf * static_cast<float>(up); // ... h *= up (in float32)
  }
  return ASR_OK;
}

extern "C" int asr_resample_batch(const void* in_dev, int32_t dtype, const int64_t* in_offsets_dev,
                                  const int32_t* in_lengths_dev, int32_t n_clips, int32_t max_in_length, int32_t up,
                                  int32_t down, const float* taps_dev, int32_t n_taps, int32_t n_pre_remove,
                                  float* out_dev, const int64_t* out_offsets_dev, void* stream) {
  if (n_clips < 0 || up < 1 || down < 1 || n_taps < 1 || (dtype != ASR_I16 && dtype != ASR_F32)) {
    set_error("asr_resample_batch: invalid argument (dtype must be ASR_I16 or ASR_F32)");
    return ASR_ERR_INVALID;
  }
  if (n_clips == 0 || max_in_length <= 0) return ASR_OK;
  if (!in_dev || !in_offsets_dev || !in_lengths_dev || !taps_dev || !out_dev || !out_offsets_dev) {
    set_error("asr_resample_batch: null pointer");
    return ASR_ERR_INVALID;
  }
  const int64_t max_out = asr_resample_out_len(max_in_length, up, down);
  const unsigned bx = static_cast<unsigned>(std::min<int64_t>((max_out + 255) / 256, 1024));
  const dim3 grid(bx, static_cast<unsigned>(n_clips));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long* io = reinterpret_cast<const long long*>(in_offsets_dev);
  const long long* oo = reinterpret_cast<const long long*>(out_offsets_dev);
  if (dtype == ASR_I16)
    resample_kernel<ASR_I16><<<grid, 256, 0, st>>>(in_dev, io, in_lengths_dev, up, down, taps_dev, n_taps, n_pre_remove, out_dev, oo);
  else
    resample_kernel<ASR_F32><<<grid, 256, 0, st>>>(in_dev, io, in_lengths_dev, up, down, taps_dev, n_taps, n_pre_remove, out_dev, oo);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}
