// Sample access shared by the block-pipelined kernels (frames_kernel.cu, tile_kernel.cu): dtype decode,
// additive noise in float64 with two roundings (VDR/attacks.py:84-85,241-244), pre-emphasis, reflect / zero padding.
#pragma once
#include "common.cuh"

namespace asr {

// ------------------------------------------------------------------------------------------------
// sample decode (+ additive noise, float64 with two roundings) at GLOBAL element index i
template <int DT>
__device__ __forceinline__ float clean_at(const FParams& fp, const long long i) {
  if (DT == ASR_I16) return static_cast<float>(__ldg(reinterpret_cast<const short*>(fp.audio) + i)) * (1.0f / 32768.0f);
  if (DT == ASR_F32) return __ldg(reinterpret_cast<const float*>(fp.audio) + i);
  return static_cast<float>(__ldg(reinterpret_cast<const double*>(fp.audio) + i));
}

template <int DT>
__device__ __forceinline__ float value_at(const FParams& fp, const long long i, const double sig) {
  if (fp.noise_mode == ASR_NOISE_NONE) return clean_at<DT>(fp, i);
  double xd;
  if (DT == ASR_F64) xd = __ldg(reinterpret_cast<const double*>(fp.audio) + i);
  else xd = static_cast<double>(clean_at<DT>(fp, i));
  double nz;
  if (fp.noise_mode == ASR_NOISE_WHITE) {
    nz = __dmul_rn(sig, __ldg(fp.z + i));
  } else {
    const double sel = (fabs(__ldg(fp.z + i)) < fp.mix_p) ? fp.mix_s1 : fp.mix_s0;
    nz = __dmul_rn(sel, __ldg(fp.z2 + i));
  }
  return static_cast<float>(__dadd_rn(xd, nz));
}

// signal at ORIGINAL index o of the clip after [noise] and [pre-emphasis]
template <int DT>
__device__ __forceinline__ float signal_at(const FParams& fp, const long long base, const int o, const double sig) {
  const float x = value_at<DT>(fp, base + o, sig);
  if (fp.preemph == 0.0f) return x;
  if (o > 0) return __fadd_rn(x, __fmul_rn(-fp.preemph, value_at<DT>(fp, base + o - 1, sig)));
  // librosa.effects.preemphasis: lfilter state zi = 2*y[0]-y[1]  ->  out[0] = y[0] + zi
  const float y1 = value_at<DT>(fp, base + 1, sig);
  return __fadd_rn(x, __fadd_rn(2.0f * x, -y1));
}

// signal at padded position p (reflect / zero padding)
template <int DT>
__device__ __forceinline__ float padded_at(const FParams& fp, const long long base, const int L, const int p,
                                           const double sig) {
  int o = p - fp.pad;
  if (o < 0) { if (fp.pad_mode != ASR_PAD_REFLECT) return 0.0f; o = -o; }
  else if (o >= L) { if (fp.pad_mode != ASR_PAD_REFLECT) return 0.0f; o = 2 * (L - 1) - o; }
  return signal_at<DT>(fp, base, o, sig);
}

}  // namespace asr
