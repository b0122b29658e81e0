// Fused MFCC kernel for sm_100a: one CTA per clip, frames processed in batches of `fb`.
//
//   stage audio chunk (dtype decode, [noise mix], [pre-emphasis], reflect/zero padding) -> smem
//   per frame: window -> real FFT (register radix-2 DIT passes, one smem exchange) -> |X|^2  (smem)
//   per batch: sparse mel bank (lanes <-> frames, float4 smem reads) -> 10*log10 -> log-mel (smem)
//   per clip : clip-wide max -> top_db clamp -> DCT-II*lifter -> [delta, delta-delta] -> global
//
// Arithmetic restated from librosa.feature.mfcc (oracle/librosa_ref.py holds the CPU restatement);
// call sites replaced: VDR/extract_features_construct_dataset.py:30, SR/...:227-228,
// VDR/attacks.py:114,267, SR/attacks.py:140-141,289-290.
#include <cooperative_groups.h>
#include <cstring>
#include "common.cuh"

namespace asr {

// ------------------------------------------------------------------------------------------------
// compile-time twiddles exp(-2*pi*i*j/32), j = 0..15
__host__ __device__ constexpr float cos32(int j) {
  switch (j) {
    case 0: return 1.0f;
    case 1: return 0.98078528040323043f;
    case 2: return 0.92387953251128674f;
    case 3: return 0.83146961230254524f;
    case 4: return 0.70710678118654757f;
    case 5: return 0.55557023301960229f;
    case 6: return 0.38268343236508984f;
    case 7: return 0.19509032201612833f;
    case 8: return 0.0f;
    case 9: return -0.19509032201612819f;
    case 10: return -0.38268343236508973f;
    case 11: return -0.55557023301960196f;
    case 12: return -0.70710678118654746f;
    case 13: return -0.83146961230254535f;
    case 14: return -0.92387953251128674f;
    default: return -0.98078528040323043f;
  }
}
__host__ __device__ constexpr float sin32(int j) {
  switch (j) {
    case 0: return 0.0f;
    case 1: return 0.19509032201612825f;
    case 2: return 0.38268343236508978f;
    case 3: return 0.55557023301960218f;
    case 4: return 0.70710678118654746f;
    case 5: return 0.83146961230254524f;
    case 6: return 0.92387953251128674f;
    case 7: return 0.98078528040323043f;
    case 8: return 1.0f;
    case 9: return 0.98078528040323043f;
    case 10: return 0.92387953251128674f;
    case 11: return 0.83146961230254546f;
    case 12: return 0.70710678118654757f;
    case 13: return 0.55557023301960218f;
    case 14: return 0.38268343236508989f;
    default: return 0.19509032201612861f;
  }
}

template <int P>
__host__ __device__ constexpr int brev(int i) {
  int r = 0;
  for (int b = 1; b < P; b <<= 1) {
    r = (r << 1) | (i & 1);
    i >>= 1;
  }
  return r;
}

// In-register radix-2 decimation-in-time DFT of P complex points (P <= 32).
// Input in bit-reversed order, output in natural order.  Non-trivial butterflies use the
// 6-FMA form  a' = a + w*b ; b' = 2a - a'.
template <int P>
__device__ __forceinline__ void dft_dit(float (&re)[P], float (&im)[P]) {
#pragma unroll
  for (int m = 2; m <= P; m *= 2) {
    const int h = m / 2;
#pragma unroll
    for (int g = 0; g < P; g += m) {
#pragma unroll
      for (int j = 0; j < h; ++j) {
        const int a = g + j, b = a + h;
        const int tw = j * (32 / m);
        const float ar = re[a], ai = im[a], br = re[b], bi = im[b];
        if (tw == 0) {
          re[a] = ar + br; im[a] = ai + bi;
          re[b] = ar - br; im[b] = ai - bi;
        } else if (tw == 8) {   // w = -i : w*b = (bi, -br)
          re[a] = ar + bi; im[a] = ai - br;
          re[b] = ar - bi; im[b] = ai + br;
        } else {
          const float wr = cos32(tw), wi = -sin32(tw);
          float nr = fmaf(wr, br, ar);
          nr = fmaf(-wi, bi, nr);
          float ni = fmaf(wr, bi, ai);
          ni = fmaf(wi, br, ni);
          re[a] = nr; im[a] = ni;
          re[b] = fmaf(2.0f, ar, -nr);
          im[b] = fmaf(2.0f, ai, -ni);
        }
      }
    }
  }
}

template <int NFFT> struct FftCfg;
template <> struct FftCfg<512>  { static constexpr int M = 256,  G = 16, P = 16; };
template <> struct FftCfg<1024> { static constexpr int M = 512,  G = 16, P = 32; };
template <> struct FftCfg<2048> { static constexpr int M = 1024, G = 32, P = 32; };

// Power spectrum of one real frame of NFFT samples via a complex FFT of M = NFFT/2 points held by
// a group of G lanes (P = M/G points per lane):   n = n1 + G*n2 ,  k = k2 + P*k1
//   pass 1 (lane n1): DFT_P over n2, times W_M^(n1*k2)            -> smem exchange
//   pass 2 (lane l ): DFT_G over n1 for k2 = l + G*q              -> Z[k]
//   unpack          : X[k], X[M-k] from Z[k], conj Z[M-k]         -> |X|^2 into fbuf[0..M]
// fbuf is reused as exchange buffer, Z buffer and finally the power spectrum.
template <int NFFT>
__device__ __forceinline__ void frame_power_fft(const float* __restrict__ xs, const bool aligned2,
                                                const float2* __restrict__ win2, const float* __restrict__ twp,
                                                const float2* __restrict__ twu, float* __restrict__ fbuf,
                                                const int l) {
  constexpr int M = FftCfg<NFFT>::M, G = FftCfg<NFFT>::G, P = FftCfg<NFFT>::P;
  constexpr int Q = P / G;
  float2* xb = reinterpret_cast<float2*>(fbuf);
  {
    float re[P], im[P];
    if (aligned2) {
      const float2* xs2 = reinterpret_cast<const float2*>(xs);
#pragma unroll
      for (int n2 = 0; n2 < P; ++n2) {
        const int n = l + G * n2;
        const float2 x = xs2[n];
        const float2 w = win2[n];
        re[brev<P>(n2)] = x.x * w.x;
        im[brev<P>(n2)] = x.y * w.y;
      }
    } else {
#pragma unroll
      for (int n2 = 0; n2 < P; ++n2) {
        const int n = l + G * n2;
        const float2 w = win2[n];
        re[brev<P>(n2)] = xs[2 * n] * w.x;
        im[brev<P>(n2)] = xs[2 * n + 1] * w.y;
      }
    }
    dft_dit<P>(re, im);
    const float4* tw4 = reinterpret_cast<const float4*>(twp + l * (2 * P + 4));
#pragma unroll
    for (int k2 = 0; k2 < P; k2 += 2) {
      const float4 t = tw4[k2 / 2];
      if (k2 != 0) {
        const float r = re[k2], i = im[k2];
        re[k2] = fmaf(r, t.x, -i * t.y);
        im[k2] = fmaf(r, t.y, i * t.x);
      }
      const float r = re[k2 + 1], i = im[k2 + 1];
      re[k2 + 1] = fmaf(r, t.z, -i * t.w);
      im[k2 + 1] = fmaf(r, t.w, i * t.z);
    }
#pragma unroll
    for (int k2 = 0; k2 < P; ++k2) xb[l * (P + 1) + k2] = make_float2(re[k2], im[k2]);
  }
  __syncwarp();
  float ur[Q][G], ui[Q][G];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
#pragma unroll
    for (int n1 = 0; n1 < G; ++n1) {
      const float2 a = xb[n1 * (P + 1) + l + G * q];
      ur[q][brev<G>(n1)] = a.x;
      ui[q][brev<G>(n1)] = a.y;
    }
  }
#pragma unroll
  for (int q = 0; q < Q; ++q) dft_dit<G>(ur[q], ui[q]);
  __syncwarp();
#pragma unroll
  for (int q = 0; q < Q; ++q)
#pragma unroll
    for (int k1 = 0; k1 < G; ++k1) xb[l + G * q + P * k1] = make_float2(ur[q][k1], ui[q][k1]);
  __syncwarp();
  float2 zp[Q][G / 2];
#pragma unroll
  for (int q = 0; q < Q; ++q)
#pragma unroll
    for (int k1 = 0; k1 < G / 2; ++k1) {
      const int k = l + G * q + P * k1;
      zp[q][k1] = xb[(M - k) & (M - 1)];
    }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < Q; ++q)
#pragma unroll
    for (int k1 = 0; k1 < G / 2; ++k1) {
      const int k = l + G * q + P * k1;
      const float2 w = twu[k];                 // (-sin(2 pi k/N)/2, -cos(2 pi k/N)/2)
      const float ar = ur[q][k1], ai = ui[q][k1];
      const float br = zp[q][k1].x, bi = zp[q][k1].y;
      const float sr = ar + br, si = ai - bi;  // A + conj(B)
      const float dr = ar - br, di = ai + bi;  // A - conj(B)
      const float tr = fmaf(w.x, dr, -w.y * di);
      const float ti = fmaf(w.x, di, w.y * dr);
      const float xr = fmaf(0.5f, sr, tr), xi = fmaf(0.5f, si, ti);
      const float yr = fmaf(0.5f, sr, -tr), yi = fmaf(0.5f, si, -ti);
      fbuf[k] = fmaf(xr, xr, xi * xi);
      fbuf[M - k] = fmaf(yr, yr, yi * yi);
    }
  if (l == 0) fbuf[M / 2] = fmaf(ur[0][G / 2], ur[0][G / 2], ui[0][G / 2] * ui[0][G / 2]);
  if (l < 16) fbuf[M + 1 + l] = 0.0f;         // tail read by the (fixed 4-quad) mel tasks
}

// Direct DFT for any n_fft (the reference's speaker preset uses n_fft = 441 = 3^2 * 7^2).
// One warp per frame; lane handles bins k = lane + 32*j.  fbuf[0..n_fft) holds the windowed frame,
// the power spectrum goes to fbuf[s_off..].
__device__ __forceinline__ void frame_power_dft(const float* __restrict__ xs, const float* __restrict__ win,
                                                const float2* __restrict__ cs, const int n_fft, const int n_bins,
                                                float* __restrict__ fbuf, const int s_off, const int lane) {
  for (int n = lane; n < n_fft; n += 32) fbuf[n] = xs[n] * win[n];
  __syncwarp();
  constexpr int KJ = 8;
  for (int kb0 = 0; kb0 < n_bins; kb0 += 32 * KJ) {
    int k[KJ], idx[KJ];
    float ar[KJ], ai[KJ];
#pragma unroll
    for (int j = 0; j < KJ; ++j) {
      k[j] = kb0 + lane + 32 * j;
      if (k[j] >= n_bins) k[j] = 0;
      idx[j] = 0; ar[j] = 0.0f; ai[j] = 0.0f;
    }
    for (int n = 0; n < n_fft; ++n) {
      const float x = fbuf[n];
#pragma unroll
      for (int j = 0; j < KJ; ++j) {
        const float2 c = cs[idx[j]];
        ar[j] = fmaf(x, c.x, ar[j]);
        ai[j] = fmaf(x, c.y, ai[j]);
        idx[j] += k[j];
        if (idx[j] >= n_fft) idx[j] -= n_fft;
      }
    }
#pragma unroll
    for (int j = 0; j < KJ; ++j) {
      const int kk = kb0 + lane + 32 * j;
      if (kk < n_bins) fbuf[s_off + kk] = fmaf(ar[j], ar[j], ai[j] * ai[j]);
    }
  }
  if (lane < 16) fbuf[s_off + n_bins + lane] = 0.0f;   // tail read by the (fixed 4-quad) mel tasks
}

// ------------------------------------------------------------------------------------------------
// audio decode (+ fused additive noise).  The mix is float64(x) + s*z with two roundings.
__device__ __forceinline__ float clean_f32(const void* __restrict__ audio, const int dtype, const long long i) {
  if (dtype == ASR_I16) return static_cast<float>(__ldg(reinterpret_cast<const short*>(audio) + i)) * (1.0f / 32768.0f);
  if (dtype == ASR_F32) return __ldg(reinterpret_cast<const float*>(audio) + i);
  return static_cast<float>(__ldg(reinterpret_cast<const double*>(audio) + i));
}

__device__ __forceinline__ float sample_at(const KParams& kp, const long long i, const double sig) {
  if (kp.noise_mode == ASR_NOISE_NONE) return clean_f32(kp.audio, kp.dtype, i);
  double xd;
  if (kp.dtype == ASR_F64) xd = __ldg(reinterpret_cast<const double*>(kp.audio) + i);
  else xd = static_cast<double>(clean_f32(kp.audio, kp.dtype, i));
  double nz;
  if (kp.noise_mode == ASR_NOISE_WHITE) {
    nz = __dmul_rn(sig, __ldg(kp.z + i));
  } else {
    const double q = __ldg(kp.z + i);
    const double sel = (fabs(q) < kp.mix_p) ? kp.mix_s1 : kp.mix_s0;
    nz = __dmul_rn(sel, __ldg(kp.z2 + i));
  }
  return static_cast<float>(__dadd_rn(xd, nz));
}

// signal value at ORIGINAL index o (0 <= o < L) after [noise] and [pre-emphasis]
__device__ __forceinline__ float signal_at(const KParams& kp, const long long base, const int o, const double sig) {
  const float x = sample_at(kp, base + o, sig);
  if (kp.preemph == 0.0f) return x;
  if (o > 0) return __fadd_rn(x, __fmul_rn(-kp.preemph, sample_at(kp, base + o - 1, sig)));
  // librosa.effects.preemphasis: lfilter state zi = 2*y[0]-y[1]  ->  out[0] = y[0] + zi
  const float y1 = sample_at(kp, base + 1, sig);
  return __fadd_rn(x, __fadd_rn(2.0f * x, -y1));
}

// ------------------------------------------------------------------------------------------------
// Staging of one batch's padded signal span into shared memory.
//
// The span is cut into groups of 8 samples aligned in the GLOBAL sample index, so every group that
// lies inside the clip is fetched with 16-byte loads (1 for int16, 2 for float32, 4 for float64) and
// written with two 16-byte shared stores; s_audio[k] holds global sample g_al + k, the first padded
// position of the batch sits at s_audio[shift].  Groups that touch a clip edge (reflect / zero
// padding) or need pre-emphasis take the element-wise path.  For int16 / float32 audio the loads of
// the NEXT batch are issued before the current batch's FFTs (register prefetch), the noise streams of
// the next batch are pulled into L2 with prefetch hints.
struct Stage {
  long long g_al;   // global element index of s_audio[0] (multiple of 8)
  int shift;        // s_audio[shift] = padded position t0*hop
  int n;            // padded samples the batch needs
  int n_groups;
};

__device__ __forceinline__ Stage stage_setup(const KParams& kp, const long long base, const int t0, const int nb) {
  Stage s;
  const long long g0 = base + static_cast<long long>(t0) * kp.hop - kp.pad;
  s.g_al = (g0 >= 0 ? g0 : g0 - 7) / 8 * 8;
  s.shift = static_cast<int>(g0 - s.g_al);
  s.n = (nb - 1) * kp.hop + kp.n_fft;
  s.n_groups = (s.shift + s.n + 7) / 8;
  return s;
}

__device__ __forceinline__ bool group_inside(const Stage& s, const int gi, const long long base, const int L) {
  const long long e0 = s.g_al + 8LL * gi;
  return e0 >= base && e0 + 8 <= base + L;
}

__device__ __forceinline__ void stage_prefetch(const KParams& kp, const Stage& s, const long long base, const int L,
                                               const int tid, int4 (&pre)[2][2]) {
  if (!kp.vec_ok || kp.dtype == ASR_F64 || kp.preemph != 0.0f) return;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int gi = tid + r * kThreads;
    if (gi < s.n_groups && group_inside(s, gi, base, L)) {
      const long long e0 = s.g_al + 8LL * gi;
      if (kp.dtype == ASR_I16) {
        pre[r][0] = __ldg(reinterpret_cast<const int4*>(reinterpret_cast<const short*>(kp.audio) + e0));
      } else {
        const int4* p = reinterpret_cast<const int4*>(reinterpret_cast<const float*>(kp.audio) + e0);
        pre[r][0] = __ldg(p);
        pre[r][1] = __ldg(p + 1);
      }
    }
  }
  if (kp.noise_mode != ASR_NOISE_NONE) {
    // one 128-byte line of each noise stream per thread -> L2
    const long long lines = (static_cast<long long>(s.n_groups) * 64 + 127) / 128;
    for (long long ln = tid; ln < lines; ln += kThreads) {
      const long long e = s.g_al + ln * 16;
      if (e >= base && e < base + L) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(kp.z + e));
        if (kp.noise_mode == ASR_NOISE_MIXTURE) asm volatile("prefetch.global.L2 [%0];" ::"l"(kp.z2 + e));
      }
    }
  }
}

__device__ __forceinline__ void unpack_i16(const int4 raw, float (&v)[8]) {
  const int w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = static_cast<float>(static_cast<short>(w[j] & 0xffff)) * (1.0f / 32768.0f);
    v[2 * j + 1] = static_cast<float>(w[j] >> 16) * (1.0f / 32768.0f);
  }
}

// one group of 8 samples -> s_audio[8*gi .. 8*gi+8)
__device__ __forceinline__ void stage_group(const KParams& kp, const Stage& s, const long long base, const int L,
                                            const double sig, const int gi, const bool have_pre, const int4 pre0,
                                            const int4 pre1, float* __restrict__ s_audio) {
  const long long e0 = s.g_al + 8LL * gi;
  float v[8];
  if (kp.vec_ok && kp.preemph == 0.0f && group_inside(s, gi, base, L)) {
    double xd[8];
    if (kp.dtype == ASR_I16) {
      const int4 raw = have_pre ? pre0 : __ldg(reinterpret_cast<const int4*>(reinterpret_cast<const short*>(kp.audio) + e0));
      unpack_i16(raw, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) xd[j] = static_cast<double>(v[j]);
    } else if (kp.dtype == ASR_F32) {
      const int4* p = reinterpret_cast<const int4*>(reinterpret_cast<const float*>(kp.audio) + e0);
      const int4 a = have_pre ? pre0 : __ldg(p);
      const int4 b = have_pre ? pre1 : __ldg(p + 1);
      v[0] = __int_as_float(a.x); v[1] = __int_as_float(a.y); v[2] = __int_as_float(a.z); v[3] = __int_as_float(a.w);
      v[4] = __int_as_float(b.x); v[5] = __int_as_float(b.y); v[6] = __int_as_float(b.z); v[7] = __int_as_float(b.w);
#pragma unroll
      for (int j = 0; j < 8; ++j) xd[j] = static_cast<double>(v[j]);
    } else {
      const double2* p = reinterpret_cast<const double2*>(reinterpret_cast<const double*>(kp.audio) + e0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double2 d = __ldg(p + j);
        xd[2 * j] = d.x; xd[2 * j + 1] = d.y;
        v[2 * j] = static_cast<float>(d.x); v[2 * j + 1] = static_cast<float>(d.y);
      }
    }
    if (kp.noise_mode != ASR_NOISE_NONE) {
      double zz[8];
      const double2* zp = reinterpret_cast<const double2*>(kp.z + e0);
#pragma unroll
      for (int j = 0; j < 4; ++j) { const double2 d = __ldg(zp + j); zz[2 * j] = d.x; zz[2 * j + 1] = d.y; }
      if (kp.noise_mode == ASR_NOISE_WHITE) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = static_cast<float>(__dadd_rn(xd[j], __dmul_rn(sig, zz[j])));
      } else {
        const double2* gp = reinterpret_cast<const double2*>(kp.z2 + e0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double2 g = __ldg(gp + j);
          const double s0 = (fabs(zz[2 * j]) < kp.mix_p) ? kp.mix_s1 : kp.mix_s0;
          const double s1 = (fabs(zz[2 * j + 1]) < kp.mix_p) ? kp.mix_s1 : kp.mix_s0;
          v[2 * j] = static_cast<float>(__dadd_rn(xd[2 * j], __dmul_rn(s0, g.x)));
          v[2 * j + 1] = static_cast<float>(__dadd_rn(xd[2 * j + 1], __dmul_rn(s1, g.y)));
        }
      }
    }
  } else {
    // clip edge (reflect / zero padding), unaligned buffers or pre-emphasis: element-wise
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      const int i = 8 * gi + j - s.shift;            // padded position relative to the batch start
      long long o = e0 + j - base;                   // original sample index
      float x = 0.0f;
      if (i >= 0 && i < s.n) {
        bool ok = true;
        if (o < 0) { if (kp.pad_mode == ASR_PAD_REFLECT) o = -o; else ok = false; }
        else if (o >= L) { if (kp.pad_mode == ASR_PAD_REFLECT) o = 2LL * (L - 1) - o; else ok = false; }
        if (ok) x = signal_at(kp, base, static_cast<int>(o), sig);
      }
      s_audio[8 * gi + j] = x;
    }
    return;
  }
  float4* dst = reinterpret_cast<float4*>(s_audio + 8 * gi);
  dst[0] = make_float4(v[0], v[1], v[2], v[3]);
  dst[1] = make_float4(v[4], v[5], v[6], v[7]);
}

__device__ __forceinline__ void stage_commit(const KParams& kp, const Stage& s, const long long base, const int L,
                                             const double sig, const int tid, const int4 (&pre)[2][2],
                                             float* __restrict__ s_audio) {
  const bool have_pre = kp.vec_ok && kp.dtype != ASR_F64 && kp.preemph == 0.0f;
  if (tid < s.n_groups) stage_group(kp, s, base, L, sig, tid, have_pre, pre[0][0], pre[0][1], s_audio);
  if (tid + kThreads < s.n_groups)
    stage_group(kp, s, base, L, sig, tid + kThreads, have_pre, pre[1][0], pre[1][1], s_audio);
  for (int gi = tid + 2 * kThreads; gi < s.n_groups; gi += kThreads)
    stage_group(kp, s, base, L, sig, gi, false, make_int4(0, 0, 0, 0), make_int4(0, 0, 0, 0), s_audio);
}

__device__ __forceinline__ int num_frames_dev(const KParams& kp, const int L) {
  const int padded = L + 2 * kp.pad;
  if (padded < kp.n_fft) return 0;
  return 1 + (padded - kp.n_fft) / kp.hop;
}

__device__ __forceinline__ void store_out(const KParams& kp, const long long idx, const float v) {
  if (kp.out_f64) reinterpret_cast<double*>(kp.out)[idx] = static_cast<double>(v);
  else reinterpret_cast<float*>(kp.out)[idx] = v;
}

// ------------------------------------------------------------------------------------------------
// One CLUSTER of `cs` CTAs per clip (cs = 1 for clips whose log-mel matrix fits one CTA's shared
// memory).  CTA `rank` owns frames [rank*FC, rank*FC+FC), FC = ceil(T/cs); the clip-wide maximum and
// the delta halo rows are exchanged through distributed shared memory.
template <int NFFT>
__global__ void __launch_bounds__(kThreads, (NFFT == 512 ? 2 : 1)) mfcc_kernel(const __grid_constant__ KParams kp) {
  extern __shared__ __align__(16) float smem[];
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cs = kp.cluster_size;
  const int rank = cs > 1 ? static_cast<int>(cluster.block_rank()) : 0;
  const int b = blockIdx.x / cs;
  const int L = kp.lengths[b];
  const long long base = kp.offsets[b];
  const int rows = kp.logmel_only ? kp.n_mels : kp.out_rows;
  const long long out_base = static_cast<long long>(b) * rows * kp.out_frames;

  // ---- per-clip validity (the reference raises; we flag and emit zeros).  Uniform over the cluster. ----
  const int T = num_frames_dev(kp, L);
  int st = ASR_CLIP_OK;
  if (T <= 0 || (kp.pad_mode == ASR_PAD_REFLECT && kp.pad > 0 && L <= kp.pad) || (kp.preemph != 0.0f && L < 2))
    st = ASR_CLIP_TOO_SHORT;
  else if (kp.delta_orders > 0 && !kp.logmel_only && T < kp.delta_width)
    st = ASR_CLIP_TOO_FEW_FRAMES;
  if (tid == 0 && rank == 0 && kp.status) kp.status[b] = st;
  if (st != ASR_CLIP_OK) {
    if (rank == 0)
      for (int i = tid; i < rows * kp.out_frames; i += kThreads) store_out(kp, out_base + i, 0.0f);
    return;
  }
  const int FC = (T + cs - 1) / cs;                 // frames per CTA of this clip
  const int f0 = min(T, rank * FC), f1 = min(T, f0 + FC);
  const int nF = f1 - f0;

  // ---- tables: global blob -> shared ----
  {
    float4* dst = reinterpret_cast<float4*>(smem);
    for (int i = tid; i < kp.blob_f4; i += kThreads) dst[i] = __ldg(kp.blob + i);
  }
  const float* s_window = smem + kp.off_window;
  const float* s_twp = smem + kp.off_twp;
  const float2* s_twu = reinterpret_cast<const float2*>(smem + kp.off_twu);
  const int4* s_tasks = reinterpret_cast<const int4*>(smem + kp.off_tasks);
  const float4* s_melw = reinterpret_cast<const float4*>(smem + kp.off_melw);
  const int* s_sbeg = reinterpret_cast<const int*>(smem + kp.off_sbeg);
  const int* s_stasks = reinterpret_cast<const int*>(smem + kp.off_stasks);
  const int2* s_ftasks = reinterpret_cast<const int2*>(smem + kp.off_ftasks);
  const float* s_dct = smem + kp.off_dct;
  const float* s_taps = smem + kp.off_taps;
  float* s_audio = smem + kp.sm_audio;
  float* s_frames = smem + kp.sm_frames;
  float* s_part = smem + kp.sm_part;
  float* s_lm = smem + kp.sm_lm;                    // row r = frame f0 + r
  float* s_cbuf = smem + kp.sm_cbuf;
  float* s_red = smem + kp.sm_red;

  for (int i = tid; i < nF * kp.lm_pitch; i += kThreads) s_lm[i] = 0.0f;

  const double sig = (kp.noise_mode == ASR_NOISE_WHITE) ? kp.sigma[b] : 0.0;
  const int FB = kp.fb;                             // 8 or 16
  const int fb_sh = (FB == 16) ? 4 : 3;
  const int s_off = (NFFT == 0) ? ((kp.n_fft + 3) & ~3) : 0;
  float run_max = -3.0e38f;

  int4 pre[2][2];
  pre[0][0] = pre[0][1] = pre[1][0] = pre[1][1] = make_int4(0, 0, 0, 0);
  Stage sg = stage_setup(kp, base, f0, min(FB, max(nF, 1)));
  if (nF > 0) stage_prefetch(kp, sg, base, L, tid, pre);

  for (int t0 = f0; t0 < f1; t0 += FB) {
    const int nb = min(FB, f1 - t0);
    // ---- stage the padded signal span of this batch (s_audio[shift] = padded position t0*hop) ----
    stage_commit(kp, sg, base, L, sig, tid, pre, s_audio);
    const int shift = sg.shift;
    __syncthreads();
    if (t0 + FB < f1) {                              // next batch's loads fly during this batch's FFTs
      sg = stage_setup(kp, base, t0 + FB, min(FB, f1 - t0 - FB));
      stage_prefetch(kp, sg, base, L, tid, pre);
    }
    // ---- frames -> power spectra ----
    if constexpr (NFFT != 0) {
      constexpr int G = FftCfg<NFFT>::G;
      constexpr int GPW = 32 / G;                  // frames per warp
      const int fs = warp * GPW + lane / G;
      if (warp * GPW < nb) {
        const int fr = min(fs, nb - 1);            // idle groups recompute the last frame into their own buffer
        const float* xs = s_audio + shift + fr * kp.hop;
        const bool aligned2 = ((shift + fr * kp.hop) & 1) == 0;
        frame_power_fft<NFFT>(xs, aligned2, reinterpret_cast<const float2*>(s_window), s_twp, s_twu,
                              s_frames + fs * kp.frame_stride, lane % G);
      }
    } else {
      if (warp < nb)
        frame_power_dft(s_audio + shift + warp * kp.hop, s_window, s_twu, kp.n_fft, kp.n_bins,
                        s_frames + warp * kp.frame_stride, s_off, lane);
    }
    __syncthreads();
    // ---- sparse mel bank: lanes <-> frames of the batch, task streams <-> sub-warps ----
    {
      const int f = lane & (FB - 1);
      const int stream = (warp << (5 - fb_sh)) + (lane >> fb_sh);
      const float* S = s_frames + f * kp.frame_stride + s_off;
      const int tb = s_sbeg[stream], te = s_sbeg[stream + 1];
      for (int ti = tb; ti < te; ++ti) {
        const int task = s_stasks[ti];
        const int4 tk = s_tasks[task];             // (filter, k_start, n_quads, w_off); weights padded to 4 quads
        const float4* s4 = reinterpret_cast<const float4*>(S + tk.y);
        const float4* w4 = s_melw + tk.w;
        float acc = 0.0f;
#pragma unroll
        for (int q = 0; q < kMelChunkQuads; ++q) {
          const float4 s = s4[q];
          const float4 w = w4[q];
          acc = fmaf(s.x, w.x, acc);
          acc = fmaf(s.y, w.y, acc);
          acc = fmaf(s.z, w.z, acc);
          acc = fmaf(s.w, w.w, acc);
        }
        s_part[task * FB + f] = acc;
      }
    }
    __syncthreads();
    // ---- combine partials, 10*log10, running clip max ----
    {
      const int f = tid & (FB - 1);
      if (f < nb) {
        for (int i = tid >> fb_sh; i < kp.n_mels; i += kThreads >> fb_sh) {
          const int2 ft = s_ftasks[i];
          float m = 0.0f;
          for (int j = 0; j < ft.y; ++j) m += s_part[(ft.x + j) * FB + f];
          const float db = 3.01029995663981195f * __log2f(fmaxf(kp.amin, m));
          s_lm[(t0 - f0 + f) * kp.lm_pitch + i] = db;
          run_max = fmaxf(run_max, db);
        }
      }
    }
    // next batch's staging / FFT do not touch s_part or s_lm; its barriers order the reuse of s_part
  }

  // ---- clip-wide max (power_to_db top_db spans every frame of the call) ----
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) run_max = fmaxf(run_max, __shfl_xor_sync(0xffffffffu, run_max, o));
  if (lane == 0) s_red[warp] = run_max;
  __syncthreads();
  if (tid == 0) {
    float mx = s_red[0];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) mx = fmaxf(mx, s_red[w]);
    s_red[8] = mx;
  }
  if (cs > 1) {
    cluster.sync();                                  // every CTA's s_red[8] is published
    if (tid == 0) {
      float mx = s_red[8];
      for (int r = 0; r < cs; ++r) mx = fmaxf(mx, *cluster.map_shared_rank(s_red + 8, r));
      s_red[9] = mx;
    }
  } else if (tid == 0) {
    s_red[9] = s_red[8];
  }
  __syncthreads();
  const float thr = (kp.top_db >= 0.0f) ? s_red[9] - kp.top_db : -3.0e38f;
  for (int i = tid; i < nF * kp.lm_pitch; i += kThreads) {
    const int c = i % kp.lm_pitch;
    if (c < kp.n_mels) s_lm[i] = fmaxf(s_lm[i], thr);
  }
  __syncthreads();

  const int t_out = min(T, kp.out_frames);
  const int o1 = min(f1, t_out);                     // this CTA writes output frames [f0, o1)
  // ---- zero padding in the feature domain (VDR/extract...py:36-37), by the first CTA of the clip ----
  if (rank == 0 && t_out < kp.out_frames) {
    const int wpad = kp.out_frames - t_out;
    for (int e = tid; e < rows * wpad; e += kThreads) {
      const int r = e / wpad, t = t_out + e % wpad;
      store_out(kp, out_base + static_cast<long long>(r) * kp.out_frames + t, 0.0f);
    }
  }
  if (kp.logmel_only) {
    const int w = max(0, o1 - f0);
    for (int e = tid; e < kp.n_mels * w; e += kThreads) {
      const int t = f0 + e % w, i = e / w;
      store_out(kp, out_base + static_cast<long long>(i) * kp.out_frames + t, s_lm[(t - f0) * kp.lm_pitch + i]);
    }
    if (cs > 1) cluster.sync();                      // keep s_red alive until every peer has read it
    return;
  }

  // ---- DCT-II (ortho) with the lifter folded into the matrix: lanes <-> frames ----
  const bool has_delta = kp.delta_orders > 0;
  const int nq = kp.dct_pitch / 4;
  if (!has_delta) {
    const int w = max(0, o1 - f0);
    const int n_tblk = (w + 31) / 32;
    for (int item = warp; item < n_tblk * kp.n_mfcc; item += kWarps) {
      const int c = item % kp.n_mfcc, r = (item / kp.n_mfcc) * 32 + lane;
      if (r < w) {
        const float4* l4 = reinterpret_cast<const float4*>(s_lm + r * kp.lm_pitch);
        const float4* d4 = reinterpret_cast<const float4*>(s_dct + c * kp.dct_pitch);
        float acc = 0.0f;
        for (int q = 0; q < nq; ++q) {
          const float4 lv = l4[q];
          const float4 dv = d4[q];
          acc = fmaf(lv.x, dv.x, acc);
          acc = fmaf(lv.y, dv.y, acc);
          acc = fmaf(lv.z, dv.z, acc);
          acc = fmaf(lv.w, dv.w, acc);
        }
        store_out(kp, out_base + static_cast<long long>(c) * kp.out_frames + f0 + r, acc);
      }
    }
    if (cs > 1) cluster.sync();
    return;
  }

  // ---- with deltas: cepstra of own frames plus the halo the Savitzky-Golay window reaches ----
  // librosa.feature.delta = savgol_filter(width, polyorder=order, deriv=order, mode='interp'):
  // interior taps everywhere, with the window centre clamped to [half, T-1-half] at the edges.
  if (cs > 1) cluster.sync();                        // peers' clamped log-mel rows are final
  const int half = kp.delta_width / 2;
  if (o1 > f0) {
    const int c_first = min(max(f0, half), T - 1 - half), c_last = min(max(o1 - 1, half), T - 1 - half);
    const int need_lo = min(f0, c_first - half), need_hi = max(o1 - 1, c_last + half);   // inclusive
    const int wn = need_hi - need_lo + 1;
    const int n_tblk = (wn + 31) / 32;
    for (int item = warp; item < n_tblk * kp.n_mfcc; item += kWarps) {
      const int c = item % kp.n_mfcc, r = (item / kp.n_mfcc) * 32 + lane;
      if (r < wn) {
        const int t = need_lo + r;
        const int owner = t / FC;
        const float* row = (owner == rank ? s_lm : cluster.map_shared_rank(s_lm, owner)) + (t - owner * FC) * kp.lm_pitch;
        const float4* l4 = reinterpret_cast<const float4*>(row);
        const float4* d4 = reinterpret_cast<const float4*>(s_dct + c * kp.dct_pitch);
        float acc = 0.0f;
        for (int q = 0; q < nq; ++q) {
          const float4 lv = l4[q];
          const float4 dv = d4[q];
          acc = fmaf(lv.x, dv.x, acc);
          acc = fmaf(lv.y, dv.y, acc);
          acc = fmaf(lv.z, dv.z, acc);
          acc = fmaf(lv.w, dv.w, acc);
        }
        s_cbuf[c * kp.cbuf_pitch + r] = acc;
      }
    }
    __syncthreads();
    const int w = o1 - f0;
    const int n_oblk = (w + 31) / 32;
    const int n_items = (1 + kp.delta_orders) * kp.n_mfcc * n_oblk;
    for (int item = warp; item < n_items; item += kWarps) {
      const int c = item % kp.n_mfcc;
      const int o = (item / kp.n_mfcc) % (1 + kp.delta_orders);
      const int t = f0 + (item / (kp.n_mfcc * (1 + kp.delta_orders))) * 32 + lane;
      if (t < o1) {
        float v;
        if (o == 0) {
          v = s_cbuf[c * kp.cbuf_pitch + t - need_lo];
        } else {
          const int tc = min(max(t, half), T - 1 - half);
          const float* taps = s_taps + (o - 1) * kp.delta_width;
          const float* cr = s_cbuf + c * kp.cbuf_pitch + (tc - half - need_lo);
          v = 0.0f;
          for (int j = 0; j < kp.delta_width; ++j) v = fmaf(taps[j], cr[j], v);
        }
        store_out(kp, out_base + static_cast<long long>(o * kp.n_mfcc + c) * kp.out_frames + t, v);
      }
    }
  }
  if (cs > 1) cluster.sync();                        // peers may still be reading this CTA's rows
}

// ------------------------------------------------------------------------------------------------
template <int NFFT>
static cudaError_t init_one() {
  cudaError_t e = cudaFuncSetAttribute(mfcc_kernel<NFFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(mfcc_kernel<NFFT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
}

cudaError_t mfcc_kernel_init() {
  cudaError_t e = init_one<512>();
  if (e == cudaSuccess) e = init_one<1024>();
  if (e == cudaSuccess) e = init_one<2048>();
  if (e == cudaSuccess) e = init_one<0>();
  return e;
}

template <int NFFT>
static cudaError_t launch_one(const KParams& kp, int smem_bytes, cudaStream_t stream) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(kp.n_clips) * kp.cluster_size);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kp.cluster_size;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = kp.cluster_size > 1 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, mfcc_kernel<NFFT>, kp);
}

cudaError_t launch_mfcc(const KParams& kp, int smem_bytes, cudaStream_t stream) {
  switch (kp.fft_path ? kp.n_fft : 0) {
    case 512: return launch_one<512>(kp, smem_bytes, stream);
    case 1024: return launch_one<1024>(kp, smem_bytes, stream);
    case 2048: return launch_one<2048>(kp, smem_bytes, stream);
    default: return launch_one<0>(kp, smem_bytes, stream);
  }
}

}  // namespace asr
