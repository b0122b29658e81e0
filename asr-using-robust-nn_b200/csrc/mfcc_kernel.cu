// Fused MFCC kernel for sm_100a: one CTA per clip, frames processed in batches of `fb`.
//
//   per frame: samples straight from global/L1 (dtype decode, [noise mix], [pre-emphasis], reflect/zero
//              padding) -> window -> real FFT (register radix-2 DIT passes, one smem exchange,
//              shuffle unpack) -> |X|^2  (smem)
//   per batch: sparse mel bank (lanes <-> frames, float4 smem reads) -> 10*log10 -> log-mel (smem)
//   per clip : clip-wide max -> top_db clamp -> DCT-II*lifter -> [delta, delta-delta] -> global
//
// Arithmetic restated from librosa.feature.mfcc (oracle/librosa_ref.py holds the CPU restatement);
// call sites replaced: VDR/extract_features_construct_dataset.py:30, SR/...:227-228,
// VDR/attacks.py:114,267, SR/attacks.py:140-141,289-290.
#include <cooperative_groups.h>
#include <cstring>
#include "common.cuh"
#include "fft_core.cuh"

namespace asr {

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

// signal value at padded position `p` of a clip of L samples (reflect / zero padding), 0 outside
__device__ __forceinline__ float padded_at(const KParams& kp, const long long base, const int L, const int p,
                                           const double sig) {
  int o = p - kp.pad;
  if (o < 0) { if (kp.pad_mode != ASR_PAD_REFLECT) return 0.0f; o = -o; }
  else if (o >= L) { if (kp.pad_mode != ASR_PAD_REFLECT) return 0.0f; o = 2 * (L - 1) - o; }
  return signal_at(kp, base, o, sig);
}

// Pass-1 operands of lane `l`: complex points z[n] = (x[2n], x[2n+1]) * window, n = l + G*n2, stored in
// bit-reversed order for the DIT DFT.
//
// load_frame_global: the frame lies inside the clip and needs no noise / pre-emphasis - one 4/8/16-byte
// load per sample pair straight from global memory (through L1: consecutive frames overlap by n_fft-hop).
template <int NFFT, int DT>
__device__ __forceinline__ void load_frame_global(const KParams& kp, const long long e_start,
                                                  const float2* __restrict__ win2,
                                                  const float2* __restrict__ win2_i16, const int l,
                                                  float (&re)[FftCfg<NFFT>::P], float (&im)[FftCfg<NFFT>::P]) {
  constexpr int G = FftCfg<NFFT>::G, P = FftCfg<NFFT>::P;
  if (DT == ASR_I16) {
    const int* p = reinterpret_cast<const int*>(reinterpret_cast<const short*>(kp.audio) + e_start);
#pragma unroll
    for (int n2 = 0; n2 < P; ++n2) {
      const int n = l + G * n2;
      const unsigned v = static_cast<unsigned>(__ldg(p + n)) ^ 0x80008000u;
      const float2 w = win2_i16[n];                        // window / 32768 (exact power-of-two scaling)
      // exact int16 -> float on the ALU/FMA pipes: bits(2^23 + (s + 32768)) - (2^23 + 32768) = s
      re[brev<P>(n2)] = (__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7410)) - 8421376.0f) * w.x;
      im[brev<P>(n2)] = (__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7432)) - 8421376.0f) * w.y;
    }
  } else if (DT == ASR_F32) {
    const float2* p = reinterpret_cast<const float2*>(reinterpret_cast<const float*>(kp.audio) + e_start);
#pragma unroll
    for (int n2 = 0; n2 < P; ++n2) {
      const int n = l + G * n2;
      const float2 x = __ldg(p + n);
      const float2 w = win2[n];
      re[brev<P>(n2)] = x.x * w.x;
      im[brev<P>(n2)] = x.y * w.y;
    }
  } else {
    const double2* p = reinterpret_cast<const double2*>(reinterpret_cast<const double*>(kp.audio) + e_start);
#pragma unroll
    for (int n2 = 0; n2 < P; ++n2) {
      const int n = l + G * n2;
      const double2 x = __ldg(p + n);
      const float2 w = win2[n];
      re[brev<P>(n2)] = static_cast<float>(x.x) * w.x;
      im[brev<P>(n2)] = static_cast<float>(x.y) * w.y;
    }
  }
}

// load_frame_smem: the frame's samples were staged (float32, after noise / pre-emphasis / padding) at xs.
template <int NFFT>
__device__ __forceinline__ void load_frame_smem(const float* __restrict__ xs, const bool aligned2,
                                                const float2* __restrict__ win2, const int l,
                                                float (&re)[FftCfg<NFFT>::P], float (&im)[FftCfg<NFFT>::P]) {
  constexpr int G = FftCfg<NFFT>::G, P = FftCfg<NFFT>::P;
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
}

// Warp-cooperative staging of `count` samples starting at padded position p_begin of the clip into dst
// (float32): dtype decode, [float64 noise mix], [pre-emphasis], reflect / zero padding.  Spans inside the
// clip with an even global start are processed two samples per lane with vector loads.
template <int DT>
__device__ __forceinline__ void fill_span(const KParams& kp, const long long base, const int L, const int p_begin,
                                          const int count, const double sig, float* __restrict__ dst, const int lane) {
  const int o_begin = p_begin - kp.pad;
  const long long e_begin = base + o_begin;
  const bool vec = kp.vec_ok && kp.preemph == 0.0f && o_begin >= 0 && o_begin + count <= L && (e_begin & 1) == 0 &&
                   (count & 1) == 0;
  if (!vec) {
    for (int i = lane; i < count; i += 32) dst[i] = padded_at(kp, base, L, p_begin + i, sig);
    return;
  }
  float2* dst2 = reinterpret_cast<float2*>(dst);
  const bool white = kp.noise_mode == ASR_NOISE_WHITE;
  for (int n = lane; 2 * n < count; n += 32) {
    double x0, x1;
    if (DT == ASR_I16) {
      const int v = __ldg(reinterpret_cast<const int*>(reinterpret_cast<const short*>(kp.audio) + e_begin) + n);
      x0 = static_cast<double>(static_cast<float>(static_cast<short>(v)) * (1.0f / 32768.0f));
      x1 = static_cast<double>(static_cast<float>(v >> 16) * (1.0f / 32768.0f));
    } else if (DT == ASR_F32) {
      const float2 x = __ldg(reinterpret_cast<const float2*>(reinterpret_cast<const float*>(kp.audio) + e_begin) + n);
      x0 = static_cast<double>(x.x); x1 = static_cast<double>(x.y);
    } else {
      const double2 x = __ldg(reinterpret_cast<const double2*>(reinterpret_cast<const double*>(kp.audio) + e_begin) + n);
      x0 = x.x; x1 = x.y;
    }
    if (kp.noise_mode != ASR_NOISE_NONE) {
      const double2 z = __ldg(reinterpret_cast<const double2*>(kp.z + e_begin) + n);
      double s0 = sig, s1 = sig, g0 = z.x, g1 = z.y;
      if (!white) {
        const double2 g = __ldg(reinterpret_cast<const double2*>(kp.z2 + e_begin) + n);
        s0 = (fabs(z.x) < kp.mix_p) ? kp.mix_s1 : kp.mix_s0;
        s1 = (fabs(z.y) < kp.mix_p) ? kp.mix_s1 : kp.mix_s0;
        g0 = g.x; g1 = g.y;
      }
      x0 = __dadd_rn(x0, __dmul_rn(s0, g0));
      x1 = __dadd_rn(x1, __dmul_rn(s1, g1));
    }
    dst2[n] = make_float2(static_cast<float>(x0), static_cast<float>(x1));
  }
}

// Direct DFT for any n_fft (the reference's speaker preset uses n_fft = 441 = 3^2 * 7^2).
// One warp per frame; lane handles bins k = lane + 32*j.  fbuf[0..n_fft) holds the windowed frame,
// the power spectrum goes to fbuf[s_off..].
__device__ __forceinline__ void frame_power_dft(const KParams& kp, const float* __restrict__ win,
                                                const float2* __restrict__ cs, float* __restrict__ fbuf,
                                                const int s_off, const int lane) {
  const int n_fft = kp.n_fft, n_bins = kp.n_bins;
  for (int n = lane; n < n_fft; n += 32) fbuf[n] *= win[n];     // fbuf[0..n_fft) was staged by fill_span
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
template <int NFFT, int DT>
__global__ void __launch_bounds__(kThreads, (NFFT == 512 ? 3 : 1)) mfcc_kernel(const __grid_constant__ KParams kp) {
  extern __shared__ __align__(16) float smem[];
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cs = kp.chunk_mode ? 1 : kp.cluster_size;
  const int rank = kp.chunk_mode ? static_cast<int>(blockIdx.x % kp.n_chunks) : (cs > 1 ? static_cast<int>(cluster.block_rank()) : 0);
  const int b = kp.chunk_mode ? static_cast<int>(blockIdx.x / kp.n_chunks) : static_cast<int>(blockIdx.x / cs);
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
    if (kp.chunk_mode) return;                      // the cepstra kernel writes the zero rows of a clip without frames
    if (rank == 0)
      for (int i = tid; i < rows * kp.out_frames; i += kThreads) store_out(kp, out_base + i, 0.0f);
    return;
  }
  const int FC = kp.chunk_mode ? kp.t_cap : (T + cs - 1) / cs;   // frames per CTA of this clip
  const int f0 = min(T, rank * FC), f1 = min(T, f0 + FC);
  const int nF = f1 - f0;
  if (kp.chunk_mode && nF <= 0) return;              // (uniform over the CTA) a clip shorter than this chunk's first frame

  // ---- tables: global blob -> shared ----
  {
    float4* dst = reinterpret_cast<float4*>(smem);
    for (int i = tid; i < kp.blob_f4; i += kThreads) dst[i] = __ldg(kp.blob + i);
  }
  const float* s_window = smem + kp.off_window;
  const float* s_window_i16 = smem + kp.off_window_i16;
  const float* s_twp = smem + kp.off_twp;
  const float2* s_twu = reinterpret_cast<const float2*>(smem + kp.off_twu);
  const int2* s_tasks = reinterpret_cast<const int2*>(smem + kp.off_tasks);
  const float4* s_melw = reinterpret_cast<const float4*>(smem + kp.off_melw);
  const int2* s_ftasks = reinterpret_cast<const int2*>(smem + kp.off_ftasks);
  const int4* s_fslots = reinterpret_cast<const int4*>(smem + kp.off_fslots);
  const float* s_dct = smem + kp.off_dct;
  const float* s_taps = smem + kp.off_taps;
  float* s_frames = smem + kp.sm_frames;
  float* s_part = smem + kp.sm_part;
  float* s_lm = smem + kp.sm_lm;                    // row r = frame f0 + r
  float* s_cbuf = smem + kp.sm_cbuf;
  float* s_red = smem + kp.sm_red;

  for (int i = tid; i < nF * kp.lm_pitch; i += kThreads) s_lm[i] = 0.0f;

  const double sig = (kp.noise_mode == ASR_NOISE_WHITE) ? kp.sigma[b] : 0.0;
  const int s_off = (NFFT == 0) ? ((kp.n_fft + 3) & ~3) : 0;
  float run_max = -3.0e38f;

  for (int i = tid; i < kp.fb * kp.part_pitch; i += kThreads) s_part[i] = 0.0f;   // incl. the always-zero slots
  __syncthreads();                                  // tables and the zeroed log-mel rows are visible

  // ---- main loop: every warp runs frames -> power spectra -> mel -> log on its own (no CTA barrier) ----
  constexpr int FPW = (NFFT == 0) ? 1 : 32 / FftCfg<(NFFT == 0 ? 512 : NFFT)>::G;   // frames per warp iteration
  constexpr int SPF = 32 / FPW;                                                      // mel task streams per frame
  float* wbuf = s_frames + warp * FPW * kp.frame_stride;          // this warp's FPW frame buffers (contiguous)
  float* wpart = s_part + warp * FPW * kp.part_pitch;
  const int ntp = kp.n_tasks_padded;
  for (int tw = f0 + warp * FPW; tw < f1; tw += kWarps * FPW) {
    const int nbw = min(FPW, f1 - tw);                             // real frames of this iteration
    const int jf = lane / SPF;                                     // frame slot of this lane
    const int fr = min(jf, nbw - 1);                               // idle slots recompute the last frame
    if constexpr (NFFT != 0) {
      constexpr int G = FftCfg<NFFT>::G, P = FftCfg<NFFT>::P;
      const int l = lane % G;
      float re[P], im[P];
      const int o_first = tw * kp.hop - kp.pad, o_last = (tw + nbw - 1) * kp.hop - kp.pad;
      const bool direct = kp.noise_mode == ASR_NOISE_NONE && kp.vec_ok && kp.preemph == 0.0f && o_first >= 0 &&
                          o_last + NFFT <= L && ((base + o_first) & 1) == 0 && (nbw == 1 || (kp.hop & 1) == 0);
      if (direct) {
        load_frame_global<NFFT, DT>(kp, base + (tw + fr) * kp.hop - kp.pad, reinterpret_cast<const float2*>(s_window),
                                    reinterpret_cast<const float2*>(s_window_i16), l, re, im);
      } else {
        // stage the span of the iteration's frames in the warp's own buffers (shared by both frames when it fits)
        const int span = (nbw - 1) * kp.hop + NFFT;
        if (span <= FPW * kp.frame_stride) {
          fill_span<DT>(kp, base, L, tw * kp.hop, span, sig, wbuf, lane);
          __syncwarp();
          load_frame_smem<NFFT>(wbuf + fr * kp.hop, ((fr * kp.hop) & 1) == 0, reinterpret_cast<const float2*>(s_window), l,
                                re, im);
        } else {
          for (int j = 0; j < nbw; ++j) fill_span<DT>(kp, base, L, (tw + j) * kp.hop, NFFT, sig, wbuf + j * kp.frame_stride, lane);
          __syncwarp();
          load_frame_smem<NFFT>(wbuf + fr * kp.frame_stride, true, reinterpret_cast<const float2*>(s_window), l, re, im);
        }
      }
      frame_power_fft<NFFT, true>(re, im, s_twp, s_twu, wbuf + jf * kp.frame_stride, wbuf + jf * kp.frame_stride, l);
    } else {
      switch (kp.dtype) {
        case ASR_I16: fill_span<ASR_I16>(kp, base, L, tw * kp.hop, kp.n_fft, sig, wbuf, lane); break;
        case ASR_F32: fill_span<ASR_F32>(kp, base, L, tw * kp.hop, kp.n_fft, sig, wbuf, lane); break;
        default: fill_span<ASR_F64>(kp, base, L, tw * kp.hop, kp.n_fft, sig, wbuf, lane); break;
      }
      __syncwarp();
      frame_power_dft(kp, s_window, s_twu, wbuf, s_off, lane);
    }
    __syncwarp();
    // ---- sparse mel bank: lane <-> (frame slot, task stream); every task is 4 float4 groups of one filter ----
    {
      const float* S = wbuf + jf * kp.frame_stride + s_off;
      float* prt = wpart + jf * kp.part_pitch;
      for (int tp = lane % SPF; tp < ntp; tp += SPF) {
        const int2 tk = s_tasks[tp];                 // (first bin, partial slot)
        const float4* s4 = reinterpret_cast<const float4*>(S + tk.x);
        float acc = 0.0f;
#pragma unroll
        for (int q = 0; q < kMelChunkQuads; ++q) {
          const float4 sv = s4[q];
          const float4 w = s_melw[q * ntp + tp];
          acc = fmaf(sv.x, w.x, acc);
          acc = fmaf(sv.y, w.y, acc);
          acc = fmaf(sv.z, w.z, acc);
          acc = fmaf(sv.w, w.w, acc);
        }
        prt[tk.y] = acc;
      }
    }
    __syncwarp();
    // ---- combine partials in filter order, 10*log10, running clip max ----
    if (kp.fixed_slots) {
      // every filter has <= 8 tasks: slot lists padded with the always-zero slot, both frames in one pass
      for (int i = lane; i < kp.n_mels; i += 32) {
        const int4 sa = s_fslots[2 * i], sb = s_fslots[2 * i + 1];
#pragma unroll
        for (int j = 0; j < FPW; ++j) {
          if (j < nbw) {
            const float* prt = wpart + j * kp.part_pitch;
            float m = prt[sa.x];
            m += prt[sa.y]; m += prt[sa.z]; m += prt[sa.w];
            m += prt[sb.x]; m += prt[sb.y]; m += prt[sb.z]; m += prt[sb.w];
            const float db = 3.01029995663981195f * __log2f(fmaxf(kp.amin, m));
            s_lm[(tw + j - f0) * kp.lm_pitch + i] = db;
            run_max = fmaxf(run_max, db);
          }
        }
      }
    } else {
      for (int j = 0; j < nbw; ++j) {
        const float* prt = wpart + j * kp.part_pitch;
        float* lrow = s_lm + (tw + j - f0) * kp.lm_pitch;
        for (int i = lane; i < kp.n_mels; i += 32) {
          const int2 ft = s_ftasks[i];
          float m = 0.0f;
          for (int u = 0; u < ft.y; ++u) m += prt[ft.x + u];
          const float db = 3.01029995663981195f * __log2f(fmaxf(kp.amin, m));
          lrow[i] = db;
          run_max = fmaxf(run_max, db);
        }
      }
    }
    __syncwarp();                                     // partials / frame buffers are reused next iteration
  }

  if (kp.chunk_mode) {
    // ---- chunk mode: the chunk's log-mel rows -> transposed workspace (threads <-> frames: coalesced along time) ----
    __syncthreads();
    const long long g0 = static_cast<long long>(kp.fstart[b]) + f0;
    for (int i = warp; i < kp.n_mels; i += kWarps) {
      float* dst = kp.lm_global + static_cast<long long>(i) * kp.lm_stride + g0;
      for (int t = lane; t < nF; t += 32) dst[t] = s_lm[t * kp.lm_pitch + i];
    }
    return;
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
  const int ncg = (kp.n_mfcc + 3) / 4;                // coefficient groups of 4 (table rows padded with zeros)
  if (!has_delta) {
    const int w = max(0, o1 - f0);
    const int n_tblk = (w + 31) / 32;
    for (int item = warp; item < n_tblk * ncg; item += kWarps) {
      const int cg = item % ncg, r = (item / ncg) * 32 + lane;
      if (r < w) {
        const float4* l4 = reinterpret_cast<const float4*>(s_lm + r * kp.lm_pitch);
        const float4* d4 = reinterpret_cast<const float4*>(s_dct + 4 * cg * kp.dct_pitch);
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int q = 0; q < nq; ++q) {
          const float4 lv = l4[q];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 dv = d4[u * nq + q];
            acc[u] = fmaf(lv.x, dv.x, acc[u]);
            acc[u] = fmaf(lv.y, dv.y, acc[u]);
            acc[u] = fmaf(lv.z, dv.z, acc[u]);
            acc[u] = fmaf(lv.w, dv.w, acc[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (4 * cg + u < kp.n_mfcc)
            store_out(kp, out_base + static_cast<long long>(4 * cg + u) * kp.out_frames + f0 + r, acc[u]);
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
    for (int item = warp; item < n_tblk * ncg; item += kWarps) {
      const int cg = item % ncg, r = (item / ncg) * 32 + lane;
      if (r < wn) {
        const int t = need_lo + r;
        const int owner = t / FC;
        const float* row = (owner == rank ? s_lm : cluster.map_shared_rank(s_lm, owner)) + (t - owner * FC) * kp.lm_pitch;
        const float4* l4 = reinterpret_cast<const float4*>(row);
        const float4* d4 = reinterpret_cast<const float4*>(s_dct + 4 * cg * kp.dct_pitch);
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int q = 0; q < nq; ++q) {
          const float4 lv = l4[q];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 dv = d4[u * nq + q];
            acc[u] = fmaf(lv.x, dv.x, acc[u]);
            acc[u] = fmaf(lv.y, dv.y, acc[u]);
            acc[u] = fmaf(lv.z, dv.z, acc[u]);
            acc[u] = fmaf(lv.w, dv.w, acc[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (4 * cg + u < kp.n_mfcc) s_cbuf[(4 * cg + u) * kp.cbuf_pitch + r] = acc[u];
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
template <int NFFT, int DT>
static cudaError_t init_one() {
  cudaError_t e = cudaFuncSetAttribute(mfcc_kernel<NFFT, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(mfcc_kernel<NFFT, DT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
}

template <int NFFT>
static cudaError_t init_fft() {
  cudaError_t e = init_one<NFFT, ASR_I16>();
  if (e == cudaSuccess) e = init_one<NFFT, ASR_F32>();
  if (e == cudaSuccess) e = init_one<NFFT, ASR_F64>();
  return e;
}

cudaError_t mfcc_kernel_init() {
  cudaError_t e = init_fft<512>();
  if (e == cudaSuccess) e = init_fft<1024>();
  if (e == cudaSuccess) e = init_fft<2048>();
  if (e == cudaSuccess) e = init_one<0, ASR_F32>();   // the direct-DFT kernel decodes the dtype at run time
  return e;
}

template <int NFFT, int DT>
static cudaError_t launch_one(const KParams& kp, int smem_bytes, cudaStream_t stream) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(kp.n_clips) * (kp.chunk_mode ? kp.n_chunks : kp.cluster_size));
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
  if (kp.cluster_size > 1) {
    // a cluster of this size and shared-memory footprint must be co-schedulable at all (non-portable size 16 with
    // nearly full shared memory per CTA may not be): checked once per (cluster size, shared memory) and device
    static int checked_cs[kMaxDevices] = {0}, checked_smem[kMaxDevices] = {0}, checked_ok[kMaxDevices] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < kMaxDevices && (checked_cs[dev] != kp.cluster_size || checked_smem[dev] != smem_bytes)) {
      int n = 0;
      e = cudaOccupancyMaxActiveClusters(&n, mfcc_kernel<NFFT, DT>, &cfg);
      if (e != cudaSuccess) return e;
      checked_cs[dev] = kp.cluster_size; checked_smem[dev] = smem_bytes; checked_ok[dev] = n > 0;
    }
    if (dev >= 0 && dev < kMaxDevices && !checked_ok[dev]) return cudaErrorLaunchOutOfResources;
  }
  return cudaLaunchKernelEx(&cfg, mfcc_kernel<NFFT, DT>, kp);
}

template <int NFFT>
static cudaError_t launch_fft(const KParams& kp, int smem_bytes, cudaStream_t stream) {
  switch (kp.dtype) {
    case ASR_I16: return launch_one<NFFT, ASR_I16>(kp, smem_bytes, stream);
    case ASR_F32: return launch_one<NFFT, ASR_F32>(kp, smem_bytes, stream);
    default: return launch_one<NFFT, ASR_F64>(kp, smem_bytes, stream);
  }
}

cudaError_t launch_mfcc(const KParams& kp, int smem_bytes, cudaStream_t stream) {
  switch (kp.fft_path ? kp.n_fft : 0) {
    case 512: return launch_fft<512>(kp, smem_bytes, stream);
    case 1024: return launch_fft<1024>(kp, smem_bytes, stream);
    case 2048: return launch_fft<2048>(kp, smem_bytes, stream);
    default: return launch_one<0, ASR_F32>(kp, smem_bytes, stream);
  }
}

}  // namespace asr
