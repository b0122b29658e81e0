// Block-pipelined MFCC front end for n_fft = 512 on sm_100a: three launches per batch.
//
//   frame_prefix_kernel   per clip: frame count T_b, status, prefix sum of T (the flattened frame index),
//                         clip maximum initialised to -inf
//   frames512_kernel      persistent, one 512-thread CTA per SM.  The frames of ALL clips form one flat
//                         list; a CTA owns a contiguous range of it and walks it in blocks of 32 frames.
//                         Per block, software-pipelined over one __syncthreads per iteration:
//                           stage   (block i+1) samples of the block's frames -> shared memory, ONCE per sample:
//                                   dtype decode, [float64 noise mix], [pre-emphasis], reflect / zero padding
//                           fft     (block i)   2 frames per warp: window, 256-point complex FFT in registers
//                                   (radix-2 passes, one shared-memory exchange), shuffle unpack, |X|^2
//                                   -> S[frame][bin] in shared memory
//                           mel     (block i-1) lanes <-> frames, warps <-> bin ranges: every bin feeds the falling
//                                   slope of one filter and the rising slope of the next (Slaney triangles), weights
//                                   are warp-uniform shared-memory broadcasts -> partial sums per (range, segment)
//                           combine (block i-2) lanes <-> filters: partials -> 10*log10 -> log-mel row in HBM/L2,
//                                   clip-wide maximum by atomic max
//   cepstra_kernel        per clip tile: top_db clamp against the clip maximum, DCT-II (ortho) x lifter,
//                         [delta, delta-delta], truncate / zero-pad to out_frames, float32 or float64 rows.
//
// The clip-wide maximum of power_to_db(top_db=80) is a dependency across all frames of a clip, hence the
// split after the log; the log-mel rows (n_mels floats per frame) stay in the 126 MB L2 between the launches.
// Arithmetic restated from librosa.feature.mfcc (oracle/librosa_ref.py); call sites replaced:
// VDR/extract_features_construct_dataset.py:30, VDR/attacks.py:114,267 (and the SR twins for even n_fft).
#include <cstring>
#include "common.cuh"
#include "fft_core.cuh"
#include "sample_access.cuh"

namespace asr {

// ------------------------------------------------------------------------------------------------
// frame of 512 staged samples (float32, 8-byte aligned) x window -> pass-1 operands of lane l (bit-reversed)
__device__ __forceinline__ void fft512_load(const float* __restrict__ xs, const float2* __restrict__ win2, const int l,
                                            float (&re)[16], float (&im)[16]) {
  const float2* xs2 = reinterpret_cast<const float2*>(xs);
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) {
    const int n = l + 16 * n2;
    const float2 x = xs2[n];
    const float2 w = win2[n];
    re[brev<16>(n2)] = x.x * w.x;
    im[brev<16>(n2)] = x.y * w.y;
  }
}

__device__ __forceinline__ void atomic_max_float(float* addr, const float v) {
  if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// ------------------------------------------------------------------------------------------------
// T_b, status, exclusive prefix sum of T_b over the clips.  Single CTA; clips are taken in groups of 8 x 1024 with
// coalesced, independent loads (thread t <-> clips base + 1024*c + t), one block-wide scan per 1024 clips.
__global__ void __launch_bounds__(1024) frame_prefix_kernel(const FParams fp) {
  __shared__ int s_w[8][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int carry = 0;
  for (int base = 0; base < fp.n_clips; base += 8 * 1024) {
    int T[8], inc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int b = base + 1024 * c + tid;
      T[c] = 0;
      if (b < fp.n_clips) {
        const int L = fp.lengths[b];
        const long long padded = static_cast<long long>(L) + 2 * fp.pad;
        int t = padded < fp.n_fft ? 0 : static_cast<int>(1 + (padded - fp.n_fft) / fp.hop);
        int st = ASR_CLIP_OK;
        if (t <= 0 || (fp.pad_mode == ASR_PAD_REFLECT && fp.pad > 0 && L <= fp.pad) || (fp.preemph != 0.0f && L < 2))
          st = ASR_CLIP_TOO_SHORT;
        else if (fp.delta_orders > 0 && !fp.logmel_only && t < fp.delta_width)
          st = ASR_CLIP_TOO_FEW_FRAMES;
        if (st != ASR_CLIP_OK) t = 0;
        if (fp.status) fp.status[b] = st;
        fp.clipmax[b] = __int_as_float(0xff800000);
        fp.nframes[b] = t;
        T[c] = t;
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {                // inclusive scans inside the warps
      int v = T[c];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      inc[c] = v;
      if (lane == 31) s_w[c][warp] = v;
    }
    __syncthreads();
    if (warp < 8) {                              // warp c scans the 32 warp totals of chunk c
      int v = s_w[warp][lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      s_w[warp][lane] = v;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int b = base + 1024 * c + tid;
      const int before = carry + (warp > 0 ? s_w[c][warp - 1] : 0) + inc[c] - T[c];
      if (b < fp.n_clips) fp.fstart[b] = before;
      carry += s_w[c][31];
    }
    __syncthreads();                             // s_w is rewritten by the next group
  }
  if (tid == 0) fp.fstart[fp.n_clips] = carry;
}

// ------------------------------------------------------------------------------------------------
// Staging units: 4 consecutive samples starting at a global index that is a multiple of 4, inside the clip,
// no pre-emphasis.  The loads are issued early (unit_load*) and consumed after the mel phase (unit_convert),
// so their latency is covered by the warp's own work.
template <int DT> struct UnitRaw;
template <> struct UnitRaw<ASR_I16> { int2 a; };
template <> struct UnitRaw<ASR_F32> { float4 a; };
template <> struct UnitRaw<ASR_F64> { double2 a[2]; };
struct UnitZ { double2 z[2]; };

template <int DT>
__device__ __forceinline__ void unit_load(const FParams& fp, const long long e, UnitRaw<DT>& r) {
  if constexpr (DT == ASR_I16) {
    r.a = __ldg(reinterpret_cast<const int2*>(reinterpret_cast<const short*>(fp.audio) + e));
  } else if constexpr (DT == ASR_F32) {
    r.a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(fp.audio) + e));
  } else {
    const double2* p = reinterpret_cast<const double2*>(reinterpret_cast<const double*>(fp.audio) + e);
    r.a[0] = __ldg(p); r.a[1] = __ldg(p + 1);
  }
}
__device__ __forceinline__ void unit_load_z(const double* __restrict__ z, const long long e, UnitZ& r) {
  const double2* p = reinterpret_cast<const double2*>(z + e);
  r.z[0] = __ldg(p); r.z[1] = __ldg(p + 1);
}

// Clean int16 stays unscaled (the window table carries the exact 2^-15); the int16 -> float conversion is
// exact integer-in-mantissa arithmetic on the FMA/ALU pipes: bits(2^23 + (s + 32768)) - (2^23 + 32768) = s.
// Noise: float64(x) + s*z with two roundings, then one rounding to float32.  For the mixture, zz holds the
// selector stream q and gg the carrier g.
template <int DT>
__device__ __forceinline__ float4 unit_convert(const FParams& fp, const UnitRaw<DT>& r, const UnitZ& zz, const UnitZ& gg,
                                               const double sig) {
  float v[4];
  double xd[4];
  if constexpr (DT == ASR_I16) {
    const unsigned w[2] = {static_cast<unsigned>(r.a.x), static_cast<unsigned>(r.a.y)};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const unsigned x = w[j] ^ 0x80008000u;
      v[2 * j] = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7410)) - 8421376.0f;
      v[2 * j + 1] = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7432)) - 8421376.0f;
    }
    if (fp.noise_mode == ASR_NOISE_NONE) return make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
    for (int j = 0; j < 4; ++j) xd[j] = static_cast<double>(v[j] * (1.0f / 32768.0f));
  } else if constexpr (DT == ASR_F32) {
    if (fp.noise_mode == ASR_NOISE_NONE) return r.a;
    xd[0] = static_cast<double>(r.a.x); xd[1] = static_cast<double>(r.a.y);
    xd[2] = static_cast<double>(r.a.z); xd[3] = static_cast<double>(r.a.w);
  } else {
    xd[0] = r.a[0].x; xd[1] = r.a[0].y; xd[2] = r.a[1].x; xd[3] = r.a[1].y;
    if (fp.noise_mode == ASR_NOISE_NONE)
      return make_float4(static_cast<float>(xd[0]), static_cast<float>(xd[1]), static_cast<float>(xd[2]), static_cast<float>(xd[3]));
  }
  if (fp.noise_mode == ASR_NOISE_WHITE) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      v[2 * j] = static_cast<float>(__dadd_rn(xd[2 * j], __dmul_rn(sig, zz.z[j].x)));
      v[2 * j + 1] = static_cast<float>(__dadd_rn(xd[2 * j + 1], __dmul_rn(sig, zz.z[j].y)));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const double s0 = (fabs(zz.z[j].x) < fp.mix_p) ? fp.mix_s1 : fp.mix_s0;
      const double s1 = (fabs(zz.z[j].y) < fp.mix_p) ? fp.mix_s1 : fp.mix_s0;
      v[2 * j] = static_cast<float>(__dadd_rn(xd[2 * j], __dmul_rn(s0, gg.z[j].x)));
      v[2 * j + 1] = static_cast<float>(__dadd_rn(xd[2 * j + 1], __dmul_rn(s1, gg.z[j].y)));
    }
  }
  return make_float4(v[0], v[1], v[2], v[3]);
}

// ------------------------------------------------------------------------------------------------
// Block descriptors (shared memory ring).  A block = up to 32 frame slots made of at most max_runs runs of
// consecutive frames of one clip; a run's samples are staged contiguously.
struct FRun {
  long long base;   // element offset of the clip
  int L;            // clip length
  int clip;
  int p0;           // padded position of the first staged sample (t0 * hop)
  int count;        // staged samples: (n - 1) * hop + 512
  int aud0;         // staged position of padded sample p0
  int unit0;        // first linear unit index of the run (runs staged by 4-sample units; the others hold 0 units)
  int nu;           // units of the run: count / 4, or 0 when the run is staged sample by sample
  // asynchronous staging (ASYNC kernel): original samples [ra, rb) of the clip are copied raw to shared memory
  int ra, rb;
  int raw_a;        // byte offset of the raw audio inside the raw buffer (16-byte aligned)
  int raw_z;        // byte offset of the raw float64 noise
};
struct FBlock {
  int n_runs, n_slots, fin, n_units;
  int refill, pad0, pad1, pad2;   // >= 0: reload the clip-metadata cache starting at this clip before the next iteration
  FRun run[kFrMaxRuns];
  int slot_aud[kFrBlock];    // staged position of the slot's first sample (0 for empty slots)
  int slot_g[kFrBlock];      // flattened frame index
  int slot_clip[kFrBlock];
};
constexpr int kFrRing = 8;
constexpr int kClipCache = 64;    // clips whose (offset, length, frame count) are cached in shared memory
struct ClipMeta { long long off; int L; int T; };
constexpr int kStageBatch = 3;                    // units per producer thread whose loads are in flight together

// Producer warps: stage the samples of every run of the block, once per sample.  Loads of a batch of units are
// issued back to back, then converted and stored; nobody else waits on these warps until the iteration barrier.
template <int DT>
__device__ __forceinline__ void stage_block(const FParams& fp, const FBlock& blk, float* __restrict__ aud, const int ptid,
                                            const int n_threads) {
  // clean int16 is staged UNSCALED (the window table carries the exact 2^-15); scaling by 2^15 is exact
  const float scale = (DT == ASR_I16 && fp.noise_mode == ASR_NOISE_NONE) ? 32768.0f : 1.0f;
  for (int r = 0; r < blk.n_runs; ++r) {
    const FRun& run = blk.run[r];
    const double sig = (fp.noise_mode == ASR_NOISE_WHITE) ? __ldg(fp.sigma + run.clip) : 0.0;
    float* dst = aud + run.aud0;
    if (run.nu == 0) {        // odd alignment or pre-emphasis: sample by sample
      for (int i = ptid; i < run.count; i += n_threads) dst[i] = padded_at<DT>(fp, run.base, run.L, run.p0 + i, sig) * scale;
      continue;
    }
    const int nu = run.count >> 2, orig0 = run.p0 - fp.pad;
    for (int u0 = ptid; u0 < nu; u0 += n_threads * kStageBatch) {
      UnitRaw<DT> raw[kStageBatch];
      UnitZ zz[kStageBatch];
      unsigned inside = 0;
#pragma unroll
      for (int k = 0; k < kStageBatch; ++k) {
        const int u = u0 + k * n_threads, orig = orig0 + 4 * u;
        if (u < nu && orig >= 0 && orig + 4 <= run.L) {
          const long long e = run.base + orig;
          unit_load<DT>(fp, e, raw[k]);
          if (fp.noise_mode != ASR_NOISE_NONE) unit_load_z(fp.z, e, zz[k]);
          inside |= 1u << k;
        }
      }
#pragma unroll
      for (int k = 0; k < kStageBatch; ++k) {
        const int u = u0 + k * n_threads;
        if (u < nu) {
          float4 v;
          if ((inside >> k) & 1u) {
            UnitZ gg = zz[k];
            if (fp.noise_mode == ASR_NOISE_MIXTURE) unit_load_z(fp.z2, run.base + orig0 + 4 * u, gg);   // carrier stream (rare path: not batched)
            v = unit_convert<DT>(fp, raw[k], zz[k], gg, sig);
          } else {
            const int p = run.p0 + 4 * u;
            v.x = padded_at<DT>(fp, run.base, run.L, p, sig) * scale;
            v.y = padded_at<DT>(fp, run.base, run.L, p + 1, sig) * scale;
            v.z = padded_at<DT>(fp, run.base, run.L, p + 2, sig) * scale;
            v.w = padded_at<DT>(fp, run.base, run.L, p + 3, sig) * scale;
          }
          *reinterpret_cast<float4*>(dst + 4 * u) = v;
        }
      }
    }
  }
}

// Split form of stage_block for the steady state: the loads of a thread's first kStageBatch units are issued
// before the combine and mel phases and consumed after them, so their latency is covered by the warp's own work.
template <int DT>
struct StagePf {
  UnitRaw<DT> raw[kStageBatch];
  UnitZ zz[kStageBatch];
  int dst[kStageBatch];       // staged position of the unit, -1: no such unit
  int orig[kStageBatch];      // original sample index of the unit's first sample
  int run[kStageBatch];
  unsigned inside;
};

template <int DT>
__device__ __forceinline__ void stage_issue(const FParams& fp, const FBlock& blk, const int tid, StagePf<DT>& pf) {
  pf.inside = 0;
#pragma unroll
  for (int k = 0; k < kStageBatch; ++k) {
    const int u = tid + k * kFrThreads;
    pf.dst[k] = -1;
    if (u < blk.n_units) {
      int r = 0;
      while (r + 1 < blk.n_runs && u >= blk.run[r + 1].unit0) ++r;
      const FRun& run = blk.run[r];
      const int local = u - run.unit0;
      if (local < run.nu) {
        const int orig = run.p0 - fp.pad + 4 * local;
        pf.dst[k] = run.aud0 + 4 * local;
        pf.orig[k] = orig;
        pf.run[k] = r;
        if (orig >= 0 && orig + 4 <= run.L) {
          const long long e = run.base + orig;
          unit_load<DT>(fp, e, pf.raw[k]);
          if (fp.noise_mode != ASR_NOISE_NONE) unit_load_z(fp.z, e, pf.zz[k]);
          pf.inside |= 1u << k;
        }
      }
    }
  }
}

template <int DT>
__device__ __forceinline__ void stage_unit_direct(const FParams& fp, const FRun& run, float* __restrict__ aud, const int local) {
  const float scale = (DT == ASR_I16 && fp.noise_mode == ASR_NOISE_NONE) ? 32768.0f : 1.0f;
  const double sig = (fp.noise_mode == ASR_NOISE_WHITE) ? __ldg(fp.sigma + run.clip) : 0.0;
  const int orig = run.p0 - fp.pad + 4 * local;
  float4 v;
  if (orig >= 0 && orig + 4 <= run.L) {
    const long long e = run.base + orig;
    UnitRaw<DT> raw;
    UnitZ zz, gg;
    unit_load<DT>(fp, e, raw);
    if (fp.noise_mode != ASR_NOISE_NONE) unit_load_z(fp.z, e, zz);
    if (fp.noise_mode == ASR_NOISE_MIXTURE) unit_load_z(fp.z2, e, gg);
    v = unit_convert<DT>(fp, raw, zz, gg, sig);
  } else {
    const int p = orig + fp.pad;
    v.x = padded_at<DT>(fp, run.base, run.L, p, sig) * scale;
    v.y = padded_at<DT>(fp, run.base, run.L, p + 1, sig) * scale;
    v.z = padded_at<DT>(fp, run.base, run.L, p + 2, sig) * scale;
    v.w = padded_at<DT>(fp, run.base, run.L, p + 3, sig) * scale;
  }
  *reinterpret_cast<float4*>(aud + run.aud0 + 4 * local) = v;
}

template <int DT>
__device__ __forceinline__ void stage_finish(const FParams& fp, const FBlock& blk, float* __restrict__ aud, const int tid,
                                             const StagePf<DT>& pf) {
  // clean int16 is staged UNSCALED (the window table carries the exact 2^-15); scaling by 2^15 is exact
  const float scale = (DT == ASR_I16 && fp.noise_mode == ASR_NOISE_NONE) ? 32768.0f : 1.0f;
#pragma unroll
  for (int k = 0; k < kStageBatch; ++k) {
    if (pf.dst[k] >= 0) {
      const FRun& run = blk.run[pf.run[k]];
      const double sig = (fp.noise_mode == ASR_NOISE_WHITE) ? __ldg(fp.sigma + run.clip) : 0.0;
      float4 v;
      if ((pf.inside >> k) & 1u) {
        UnitZ gg = pf.zz[k];
        if (fp.noise_mode == ASR_NOISE_MIXTURE) unit_load_z(fp.z2, run.base + pf.orig[k], gg);   // carrier stream (rare path)
        v = unit_convert<DT>(fp, pf.raw[k], pf.zz[k], gg, sig);
      } else {
        const int p = pf.orig[k] + fp.pad;
        v.x = padded_at<DT>(fp, run.base, run.L, p, sig) * scale;
        v.y = padded_at<DT>(fp, run.base, run.L, p + 1, sig) * scale;
        v.z = padded_at<DT>(fp, run.base, run.L, p + 2, sig) * scale;
        v.w = padded_at<DT>(fp, run.base, run.L, p + 3, sig) * scale;
      }
      *reinterpret_cast<float4*>(aud + pf.dst[k]) = v;
    }
  }
  // units past the prefetched ones (large hops), then runs staged sample by sample
  for (int u = tid + kStageBatch * kFrThreads; u < blk.n_units; u += kFrThreads) {
    int r = 0;
    while (r + 1 < blk.n_runs && u >= blk.run[r + 1].unit0) ++r;
    if (u - blk.run[r].unit0 < blk.run[r].nu) stage_unit_direct<DT>(fp, blk.run[r], aud, u - blk.run[r].unit0);
  }
  for (int r = 0; r < blk.n_runs; ++r) {
    const FRun& run = blk.run[r];
    if (run.nu != 0) continue;
    const double sig = (fp.noise_mode == ASR_NOISE_WHITE) ? __ldg(fp.sigma + run.clip) : 0.0;
    for (int i = tid; i < run.count; i += kFrThreads)
      aud[run.aud0 + i] = padded_at<DT>(fp, run.base, run.L, run.p0 + i, sig) * scale;
  }
}

// ------------------------------------------------------------------------------------------------
// Asynchronous staging (ASYNC kernel): the raw bytes of a block's samples (and of its float64 noise) are copied to
// shared memory with cp.async while the previous block is transformed; a short conversion phase then turns them into
// the float32 frame samples.  Nothing waits on global-memory latency any more.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, const int src_bytes) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int DT>
__device__ __forceinline__ void raw_issue(const FParams& fp, const FBlock& blk, char* __restrict__ raw, const int tid) {
  constexpr int esz = DT == ASR_I16 ? 2 : (DT == ASR_F32 ? 4 : 8);
  const int n_runs = blk.n_runs;
  for (int r = 0; r < n_runs; ++r) {
    // descriptor fields into registers first: the copies below write shared memory, which would force re-reads
    const int nu = blk.run[r].nu, ra = blk.run[r].ra, rb = blk.run[r].rb, raw_a = blk.run[r].raw_a, raw_z = blk.run[r].raw_z;
    const long long first = blk.run[r].base + ra;
    if (nu == 0) continue;
    const int n = rb - ra;
    {
      const char* src = reinterpret_cast<const char*>(fp.audio) + first * esz + tid * 16;
      char* dst = raw + raw_a + tid * 16;
      const int bytes = n * esz;
      for (int c = tid * 16; c < bytes; c += kFrThreads * 16, src += kFrThreads * 16, dst += kFrThreads * 16)
        cp_async16(dst, src, min(16, bytes - c));
    }
    if (fp.noise_mode != ASR_NOISE_NONE) {
      const char* src = reinterpret_cast<const char*>(fp.z + first) + tid * 16;
      char* dst = raw + raw_z + tid * 16;
      const int bytes = n * 8;
      for (int c = tid * 16; c < bytes; c += kFrThreads * 16, src += kFrThreads * 16, dst += kFrThreads * 16)
        cp_async16(dst, src, min(16, bytes - c));
    }
  }
  cp_async_commit();
}

// one sample from the raw copy: original index o must lie in [ra, rb)
template <int DT>
__device__ __forceinline__ float raw_sample(const FParams& fp, const FRun& run, const char* __restrict__ raw, const int o,
                                            const double sig, const float scale) {
  const int i = o - run.ra;
  float x;
  double xd;
  if constexpr (DT == ASR_I16) {
    x = static_cast<float>(reinterpret_cast<const short*>(raw + run.raw_a)[i]) * (1.0f / 32768.0f);
    xd = static_cast<double>(x);
  } else if constexpr (DT == ASR_F32) {
    x = reinterpret_cast<const float*>(raw + run.raw_a)[i];
    xd = static_cast<double>(x);
  } else {
    xd = reinterpret_cast<const double*>(raw + run.raw_a)[i];
    x = static_cast<float>(xd);
  }
  if (fp.noise_mode == ASR_NOISE_NONE) return x * scale;
  const double z = reinterpret_cast<const double*>(raw + run.raw_z)[i];
  return static_cast<float>(__dadd_rn(xd, __dmul_rn(sig, z)));
}

template <int DT>
__device__ __forceinline__ void raw_convert(const FParams& fp, const FBlock& blk, const char* __restrict__ raw,
                                            float* __restrict__ aud, const int tid) {
  constexpr int esz = DT == ASR_I16 ? 2 : (DT == ASR_F32 ? 4 : 8);
  // clean int16 is staged UNSCALED (the window table carries the exact 2^-15); scaling by 2^15 is exact
  const float scale = (DT == ASR_I16 && fp.noise_mode == ASR_NOISE_NONE) ? 32768.0f : 1.0f;
  const int n_runs = blk.n_runs;
  for (int r = 0; r < n_runs; ++r) {
    const FRun run = blk.run[r];                        // a register copy: the stores below go to shared memory too
    const double sig = (fp.noise_mode == ASR_NOISE_WHITE) ? __ldg(fp.sigma + run.clip) : 0.0;
    float* dst = aud + run.aud0;
    if (run.nu == 0) {        // odd alignment / pre-emphasis / mixture noise: straight from global memory
      for (int i = tid; i < run.count; i += kFrThreads) dst[i] = padded_at<DT>(fp, run.base, run.L, run.p0 + i, sig) * scale;
      continue;
    }
    const int orig0 = run.p0 - fp.pad;
    const char* base_a = raw + run.raw_a - run.ra * esz;       // sample o of the clip at base_a + o*esz
    const char* base_z = raw + run.raw_z - run.ra * 8;
    for (int u = tid; u < run.nu; u += kFrThreads) {
      const int orig = orig0 + 4 * u;
      float4 v;
      if (orig >= 0 && orig + 4 <= run.L) {                          // inside the clip (hence inside [ra, rb))
        UnitRaw<DT> ur;
        UnitZ uz;
        const char* pa = base_a + orig * esz;
        if constexpr (DT == ASR_I16) ur.a = *reinterpret_cast<const int2*>(pa);
        else if constexpr (DT == ASR_F32) ur.a = *reinterpret_cast<const float4*>(pa);
        else { ur.a[0] = *reinterpret_cast<const double2*>(pa); ur.a[1] = *reinterpret_cast<const double2*>(pa + 16); }
        if (fp.noise_mode != ASR_NOISE_NONE) {
          const char* pz = base_z + orig * 8;
          uz.z[0] = *reinterpret_cast<const double2*>(pz);
          uz.z[1] = *reinterpret_cast<const double2*>(pz + 16);
        }
        v = unit_convert<DT>(fp, ur, uz, uz, sig);
      } else {                                                       // reflect / zero padding at the clip edges
        float e[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int o = orig + j;
          bool zero = false;
          if (o < 0) { zero = fp.pad_mode != ASR_PAD_REFLECT; o = -o; }
          else if (o >= run.L) { zero = fp.pad_mode != ASR_PAD_REFLECT; o = 2 * (run.L - 1) - o; }
          e[j] = zero ? 0.0f : raw_sample<DT>(fp, run, raw, o, sig, scale);
        }
        v = make_float4(e[0], e[1], e[2], e[3]);
      }
      *reinterpret_cast<float4*>(dst + 4 * u) = v;
    }
  }
}

__device__ __forceinline__ float fast_log2(const float x) {     // x >= amin > 0: no denormal handling needed
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------------------------------------------
template <int DT, bool ASYNC>
__global__ void __launch_bounds__(kFrAllThreads, 1) frames512_kernel(const __grid_constant__ FParams fp) {
  extern __shared__ __align__(16) float smem[];
  __shared__ FBlock ring[kFrRing];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- this CTA's range of the flattened frame list ----
  const int total = __ldg(fp.fstart + fp.n_clips);
  const int n_blocks = (total + kFrBlock - 1) / kFrBlock;
  const int per = (n_blocks + gridDim.x - 1) / gridDim.x;
  const long long gb = static_cast<long long>(blockIdx.x) * per * kFrBlock;
  if (gb >= total) return;
  const int g_begin = static_cast<int>(gb);
  const int g_end = static_cast<int>(min(static_cast<long long>(total), gb + static_cast<long long>(per) * kFrBlock));

  // ---- tables: global blob -> shared ----
  {
    float4* dst = reinterpret_cast<float4*>(smem);
    for (int i = tid; i < fp.blob_f4; i += kFrAllThreads) dst[i] = __ldg(fp.blob + i);
  }
  const float2* s_win2 = reinterpret_cast<const float2*>(smem + fp.off_window);
  const float* s_twp = smem + fp.off_twp;
  const float2* s_twu = reinterpret_cast<const float2*>(smem + fp.off_twu);
  const float4* s_wtab = reinterpret_cast<const float4*>(smem + fp.off_wtab);
  const int4* s_pieces = reinterpret_cast<const int4*>(smem + fp.off_pieces);
  float* s_aud = smem + fp.sm_aud;          // [2][aud_cap] (ASYNC: one buffer)
  char* s_raw = reinterpret_cast<char*>(smem + fp.sm_raw);   // ASYNC: raw bytes of the block being copied in
  float* s_S = smem + fp.sm_S;              // [2][32][s_pitch]
  float* s_xb = smem + fp.sm_xb;            // [16 warps][2][xb_stride]
  float* s_part = smem + fp.sm_part;        // [2][n_refs][33]
  for (int i = tid; i < 2 * kFrBlock * fp.s_pitch; i += kFrAllThreads) s_S[i] = 0.0f;    // incl. the zero tail of every row

  // ---- block cursor (lane 0 of the last warp) over a shared-memory cache of the clips' metadata ----
  constexpr int kAsmWarp = kFrMelWarps;               // this warp assembles the block descriptors instead of taking mel segments
  __shared__ ClipMeta s_meta[kClipCache];
  int b_cur = 0, t_cur = 0, g_cur = g_begin, cache_base = 0;
  {
    int lo = 0, hi = fp.n_clips;                       // largest b with fstart[b] <= g_begin (every thread: uniform)
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(fp.fstart + mid) <= g_begin) lo = mid; else hi = mid;
    }
    b_cur = lo;
    t_cur = g_begin - __ldg(fp.fstart + lo);
    cache_base = lo;
  }
  auto load_meta = [&](const int base) {               // all threads
    if (tid < kClipCache) {
      const int b = min(base + tid, fp.n_clips - 1);
      s_meta[tid].off = __ldg(fp.offsets + b);
      s_meta[tid].L = __ldg(fp.lengths + b);
      s_meta[tid].T = __ldg(fp.nframes + b);
    }
  };
  load_meta(cache_base);
  __syncthreads();
  auto assemble_runs = [&](FBlock& blk) {              // cursor thread
    int n_slots = 0, n_runs = 0, aud = 0, n_units = 0, raw_off = 0;
#pragma unroll 1
    for (int c = 0; c < kFrMaxRuns; ++c) {
      const int b = b_cur;
      if (b >= fp.n_clips || b >= cache_base + kClipCache || n_slots >= kFrBlock || n_runs >= fp.max_runs || g_cur >= g_end) break;
      const ClipMeta cm = s_meta[b - cache_base];
      if (t_cur < cm.T) {
        const int n = min(min(kFrBlock - n_slots, cm.T - t_cur), g_end - g_cur);
        FRun& run = blk.run[n_runs];
        run.base = cm.off;
        run.L = cm.L;
        run.clip = b;
        run.p0 = t_cur * fp.hop;
        run.count = (n - 1) * fp.hop + 512;
        const long long e0 = run.base + static_cast<long long>(run.p0) - fp.pad;
        const bool vec = fp.vec_ok && (e0 & (ASYNC ? 7 : 3)) == 0;
        run.unit0 = n_units;
        run.nu = vec ? run.count >> 2 : 0;
        n_units += run.nu;
        run.aud0 = (aud + 3) & ~3;
        aud = run.aud0 + run.count;
        if (ASYNC && vec) {
          constexpr int esz = DT == ASR_I16 ? 2 : (DT == ASR_F32 ? 4 : 8);
          const int o0 = run.p0 - fp.pad, o1 = o0 + run.count;
          int ra = max(0, o0), rb = min(run.L, o1);
          if (fp.pad_mode == ASR_PAD_REFLECT) {         // sources of the reflected samples (single-frame runs at a clip edge)
            if (o0 < 0) rb = max(rb, min(run.L, 1 - o0));
            if (o1 > run.L) ra = min(ra, max(0, 2 * (run.L - 1) - (o1 - 1)));
          }
          ra &= ~7;                                     // keeps the copy 16-byte aligned for every dtype
          run.ra = ra; run.rb = rb;
          run.raw_a = raw_off;
          raw_off += ((rb - ra) * esz + 15) & ~15;
          run.raw_z = raw_off;
          if (fp.noise_mode != ASR_NOISE_NONE) raw_off += ((rb - ra) * 8 + 15) & ~15;
        }
        // slot tables of this run are filled by the lanes of the warp afterwards: stash what they need
        blk.slot_g[n_runs] = g_cur; blk.slot_clip[n_runs] = n_slots;       // (temporarily: g0 and slot0 of run r)
        n_slots += n; t_cur += n; g_cur += n; ++n_runs;
      }
      if (t_cur >= cm.T) { ++b_cur; t_cur = 0; }       // clip finished (or empty): the next one continues the block
    }
    blk.n_runs = n_runs; blk.n_slots = n_slots; blk.fin = g_cur >= g_end ? 1 : 0; blk.n_units = n_units;
    blk.refill = -1;
    if (g_cur < g_end && b_cur + kFrMaxRuns > cache_base + kClipCache) { blk.refill = b_cur; cache_base = b_cur; }
  };
  auto fill_slots = [&](FBlock& blk) {                 // all lanes of the cursor warp
    int g0[kFrMaxRuns], s0[kFrMaxRuns];
#pragma unroll
    for (int r = 0; r < kFrMaxRuns; ++r) { g0[r] = blk.slot_g[r]; s0[r] = blk.slot_clip[r]; }
    const int n_runs = blk.n_runs, n_slots = blk.n_slots;
    __syncwarp();
    int aud = 0, g = 0, clip = -1;
    if (lane < n_slots) {
#pragma unroll
      for (int r = 0; r < kFrMaxRuns; ++r)
        if (r < n_runs && lane >= s0[r]) {
          aud = blk.run[r].aud0 + (lane - s0[r]) * fp.hop;
          g = g0[r] + lane - s0[r];
          clip = blk.run[r].clip;
        }
    }
    blk.slot_aud[lane] = aud; blk.slot_g[lane] = g; blk.slot_clip[lane] = clip;
  };
  for (int i = 0; i < 2; ++i) {
    if (warp == kAsmWarp) {
      if (lane == 0) assemble_runs(ring[i]);
      __syncwarp();
      fill_slots(ring[i]);
    }
    __syncthreads();                                   // also: tables and the zeroed S
    if (ring[i].refill >= 0) { load_meta(ring[i].refill); __syncthreads(); }
  }
  if constexpr (ASYNC) {
    raw_issue<DT>(fp, ring[0], s_raw, tid);
    cp_async_wait_all();
    __syncthreads();
    raw_convert<DT>(fp, ring[0], s_raw, s_aud, tid);
    __syncthreads();
    raw_issue<DT>(fp, ring[1], s_raw, tid);
  } else {
    stage_block<DT>(fp, ring[0], s_aud, tid, kFrAllThreads);
    __syncthreads();
  }

  // ---- per-thread constants of the phases ----
  const int part_buf = fp.n_refs * 33;                 // floats per partial buffer
  const bool comb_lane = lane < fp.n_mels;
  const float* comb_pj = s_part + (comb_lane ? (2 * lane + 1) * 33 : 0);   // rise partial of filter `lane`; fall of it is 33 further
  const int2 mel_range = warp < kFrMelWarps ? reinterpret_cast<const int2*>(smem + fp.off_wrange)[warp] : make_int2(0, 0);
  const int fft_h = lane >> 4, fft_l = lane & 15;
  const int fft_slot = 8 * (warp >> 2) + (warp & 3) + 4 * fft_h;   // half-warps 4 slots apart: complementary bank halves of S
  float* const fft_xb = s_xb + (warp * 2 + fft_h) * fp.xb_stride;
  const int s_buf = kFrBlock * fp.s_pitch;
  int cm_clip[2] = {-1, -1};                           // combine: clip whose maximum is being accumulated (per slot position)
  float cm_max[2] = {-3.0e38f, -3.0e38f};
  auto flush_max = [&](const int clip, float mx) {
    if (clip >= 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (lane == 0) atomic_max_float(fp.clipmax + clip, mx);
    }
  };
  auto combine_slot = [&](const FBlock& blk, const float* part, const int slot, int& cm_clip, float& cm_max) {
    const int clip = blk.slot_clip[slot];
    if (clip != cm_clip) { flush_max(cm_clip, cm_max); cm_clip = clip; cm_max = -3.0e38f; }
    float* row = fp.lm + static_cast<long long>(blk.slot_g[slot]) * fp.lm_pitch;
    if (fp.lm_pitch <= 32) {                            // one pass: lane j <-> filter j (lanes >= n_mels write the zero pad columns)
      float db = 0.0f;
      if (comb_lane) {
        const float m = comb_pj[(part - s_part) + slot] + comb_pj[(part - s_part) + 33 + slot];
        db = 3.01029995663981195f * fast_log2(fmaxf(fp.amin, m));
        cm_max = fmaxf(cm_max, db);
      }
      if (lane < fp.lm_pitch) row[lane] = db;
    } else {
#pragma unroll 1
      for (int j = lane; j < fp.lm_pitch; j += 32) {
        float db = 0.0f;
        if (j < fp.n_mels) {
          const float m = part[(2 * j + 1) * 33 + slot] + part[(2 * j + 2) * 33 + slot];
          db = 3.01029995663981195f * fast_log2(fmaxf(fp.amin, m));
          cm_max = fmaxf(cm_max, db);
        }
        row[j] = db;
      }
    }
  };

  const FBlock* b_comb = &ring[kFrRing - 2];           // block it-2 (not live before it = 2)
  const FBlock* b_mel = &ring[kFrRing - 1];            // block it-1
  const FBlock* b_fft = &ring[0];                      // block it
  const FBlock* b_stage = &ring[1];                    // block it+1
  for (int it = 0;; ++it) {
    const int par = it & 1;
    const int n_comb = it >= 2 ? b_comb->n_slots : 0;
    const int n_mel = it >= 1 ? b_mel->n_slots : 0;
    const int n_fft = b_fft->n_slots;
    if (n_comb == 0 && n_mel == 0 && n_fft == 0 && b_fft->fin) break;      // uniform over the CTA
    FBlock* nb;
    // ---- descriptor of block it+2 (one warp, in place of its mel share) ----
    if (warp == kAsmWarp) {
      nb = &ring[(it + 2) % kFrRing];
      if (lane == 0) assemble_runs(*nb);
      __syncwarp();
      fill_slots(*nb);
    }
    // ---- stage (block it+1), part 1: issue the loads of this thread's samples (consumed after the mel phase) ----
    StagePf<DT> pf;
    const bool do_stage = !ASYNC && b_stage->n_slots > 0 && !(fp.dbg_skip & 1);
    if (do_stage) stage_issue<DT>(fp, *b_stage, tid, pf);

    // ---- combine (block it-2): lanes <-> filters; warp w takes slots w and w+16.
    //      filter j = rising slope over segment j + falling slope over segment j+1 ----
    if (n_comb > 0 && !(fp.dbg_skip & 2)) {
      const float* part = s_part + par * part_buf;
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (warp + kFrWarps * h < n_comb) combine_slot(*b_comb, part, warp + kFrWarps * h, cm_clip[h], cm_max[h]);
    }

    // ---- mel (block it-1): lanes <-> frames, this warp's segments ----
    if (n_mel > 0 && !(fp.dbg_skip & 4)) {
      const float* S = s_S + (par ^ 1) * s_buf + lane * fp.s_pitch;      // S buffer of block it-1
      float* part = s_part + (par ^ 1) * part_buf + lane;
#pragma unroll 1
      for (int pi = mel_range.x; pi < mel_range.x + mel_range.y; ++pi) {
        const int4 pc = s_pieces[pi];                 // (first bin (multiple of 4), PAIRS of float4 groups, weight offset (float4), fall-partial offset)
        const float4* sp = reinterpret_cast<const float4*>(S + pc.x);
        const float4* wt = s_wtab + pc.z;
        float a = 0.0f, b = 0.0f;
#pragma unroll 1
        for (int q = 0; q < pc.y; ++q) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float4 sv = sp[2 * q + u];
            const float4 w01 = wt[4 * q + 2 * u], w23 = wt[4 * q + 2 * u + 1];
            a = fmaf(w01.x, sv.x, a); b = fmaf(w01.y, sv.x, b);
            a = fmaf(w01.z, sv.y, a); b = fmaf(w01.w, sv.y, b);
            a = fmaf(w23.x, sv.z, a); b = fmaf(w23.y, sv.z, b);
            a = fmaf(w23.z, sv.w, a); b = fmaf(w23.w, sv.w, b);
          }
        }
        part[pc.w] = a;                               // falling slope of filter seg-1
        part[pc.w + 33] = b;                          // rising slope of filter seg
      }
    }

    // ---- stage (block it+1), part 2: convert (noise mix in float64), store to shared memory, once per sample ----
    if (do_stage) stage_finish<DT>(fp, *b_stage, s_aud + (par ^ 1) * fp.aud_cap, tid, pf);

    // ---- fft (block it) ----
    if (n_fft > 0 && !(fp.dbg_skip & 8)) {
      const float* xs = s_aud + (ASYNC ? 0 : par * fp.aud_cap) + b_fft->slot_aud[fft_slot];
      float re[16], im[16];
      fft512_load(xs, s_win2, fft_l, re, im);
      frame_power_fft<512, false, ASYNC>(re, im, s_twp, s_twu, fft_xb, s_S + par * s_buf + fft_slot * fp.s_pitch, fft_l);
    }

    nb = &ring[(it + 2) % kFrRing];
    b_comb = b_mel; b_mel = b_fft; b_fft = b_stage; b_stage = nb;
    if constexpr (ASYNC) cp_async_wait_all();          // this thread's share of the next block's raw bytes has landed
    __syncthreads();
    if (nb->refill >= 0) { load_meta(nb->refill); __syncthreads(); }   // rare: the cursor ran past the cached clips
    if constexpr (ASYNC) {
      // the FFTs are done with the sample buffer: convert the next block into it, then start copying the one after
      if (b_fft->n_slots > 0 && !(fp.dbg_skip & 1)) raw_convert<DT>(fp, *b_fft, s_raw, s_aud, tid);
      __syncthreads();
      if (b_stage->n_slots > 0 && !(fp.dbg_skip & 1)) raw_issue<DT>(fp, *b_stage, s_raw, tid);
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) flush_max(cm_clip[h], cm_max[h]);
}

// ------------------------------------------------------------------------------------------------
// Cepstra: grid (clip, time tile).  Thread <-> frame: clamp the log-mel row, DCT-II x lifter from a
// transposed table (one float4 = 4 coefficients of one mel band, broadcast), [Savitzky-Golay deltas
// through shared memory], coalesced stores along time.
constexpr int kCepThreads = 128;

template <int NC4>   // NC4 = ceil(n_mfcc / 4), 1..16
__global__ void __launch_bounds__(kCepThreads) cepstra_kernel(const __grid_constant__ FParams fp) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const int T = __ldg(fp.nframes + b);
  const int rows = fp.logmel_only ? fp.n_mels : fp.out_rows;
  const long long out_base = static_cast<long long>(b) * rows * fp.out_frames;
  const int half = (fp.delta_orders > 0 && !fp.logmel_only) ? fp.delta_width / 2 : 0;
  const int tile = kCepThreads - 2 * half;             // output frames per CTA
  const int t_lo = blockIdx.y * tile;                   // first output frame of this tile
  auto store = [&](const long long idx, const float v) {
    if (fp.out_f64) reinterpret_cast<double*>(fp.out)[idx] = static_cast<double>(v);
    else reinterpret_cast<float*>(fp.out)[idx] = v;
  };
  const int t_out = min(T, fp.out_frames);              // frames carrying data; the rest is zero padding
  // ---- zero padding in the feature domain (VDR/extract...py:36-37), also clips that cannot be framed ----
  {
    const int z_lo = max(t_out, t_lo), z_hi = min(fp.out_frames, t_lo + tile);
    const int w = z_hi - z_lo;
    for (int e = tid; w > 0 && e < rows * w; e += kCepThreads)
      store(out_base + static_cast<long long>(e / w) * fp.out_frames + z_lo + e % w, 0.0f);
  }
  if (t_lo >= t_out) return;

  float* s_dct = smem;                                  // [n_mels][4*NC4] transposed, lifter folded in
  float* s_cep = smem + fp.cep_off_cbuf;                // [n_mfcc][kCepThreads+1]
  const float* s_taps = smem + fp.cep_off_taps;
  if (!fp.logmel_only) {
    const float4* src = fp.blob + fp.cep_blob_f4;
    float4* dst = reinterpret_cast<float4*>(smem);
    for (int i = tid; i < fp.cep_tab_f4; i += kCepThreads) dst[i] = __ldg(src + i);
  }
  const float thr = (fp.top_db >= 0.0f) ? __ldg(fp.clipmax + b) - fp.top_db : -3.0e38f;
  const int g0 = __ldg(fp.fstart + b);
  // frame handled by this thread: centre frames of the tile plus the delta halo, clamped into the clip
  const int c_lo = half > 0 ? min(max(t_lo, half), T - 1 - half) - half : t_lo;   // first frame whose cepstrum the tile needs
  const int t = c_lo + tid;
  const bool live = t < T && t >= 0;
  __syncthreads();

  if (fp.logmel_only) {
    if (live && t < t_out && tid < tile) {
      const float* row = fp.lm + static_cast<long long>(g0 + t) * fp.lm_pitch;
      for (int j = 0; j < fp.n_mels; ++j)
        store(out_base + static_cast<long long>(j) * fp.out_frames + t, fmaxf(__ldcg(row + j), thr));
    }
    return;
  }

  float acc[4 * NC4];
#pragma unroll
  for (int c = 0; c < 4 * NC4; ++c) acc[c] = 0.0f;
  if (live) {
    const float4* row4 = reinterpret_cast<const float4*>(fp.lm + static_cast<long long>(g0 + t) * fp.lm_pitch);
    const float4* d4 = reinterpret_cast<const float4*>(s_dct);
    for (int jq = 0; jq < fp.lm_pitch / 4; ++jq) {
      const float4 v4 = __ldcg(row4 + jq);
      const float v[4] = {fmaxf(v4.x, thr), fmaxf(v4.y, thr), fmaxf(v4.z, thr), fmaxf(v4.w, thr)};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = 4 * jq + u;                        // rows j >= n_mels of the table are zero
#pragma unroll
        for (int c4 = 0; c4 < NC4; ++c4) {
          const float4 d = d4[j * NC4 + c4];
          acc[4 * c4] = fmaf(v[u], d.x, acc[4 * c4]);
          acc[4 * c4 + 1] = fmaf(v[u], d.y, acc[4 * c4 + 1]);
          acc[4 * c4 + 2] = fmaf(v[u], d.z, acc[4 * c4 + 2]);
          acc[4 * c4 + 3] = fmaf(v[u], d.w, acc[4 * c4 + 3]);
        }
      }
    }
  }
  if (half == 0) {
    if (live && t < t_out) {
#pragma unroll
      for (int c = 0; c < 4 * NC4; ++c)
        if (c < fp.n_mfcc) store(out_base + static_cast<long long>(c) * fp.out_frames + t, acc[c]);
    }
    return;
  }
  // ---- deltas: librosa.feature.delta = savgol_filter(width, polyorder=order, deriv=order, mode='interp'):
  //      interior taps everywhere, with the window centre clamped to [half, T-1-half] at the edges ----
  const int cp = kCepThreads + 1;
#pragma unroll
  for (int c = 0; c < 4 * NC4; ++c)
    if (c < fp.n_mfcc) s_cep[c * cp + tid] = acc[c];
  __syncthreads();
  const int to = t_lo + tid;                             // output frame of this thread
  if (tid < tile && to < t_out) {
    const int tc = min(max(to, half), T - 1 - half);
    const int i0 = tc - half - c_lo;                     // first tap position in s_cep
    for (int c = 0; c < fp.n_mfcc; ++c) {
      const float* cr = s_cep + c * cp;
      store(out_base + static_cast<long long>(c) * fp.out_frames + to, cr[to - c_lo]);
      for (int o = 1; o <= fp.delta_orders; ++o) {
        const float* taps = s_taps + (o - 1) * fp.delta_width;
        float v = 0.0f;
        for (int j = 0; j < fp.delta_width; ++j) v = fmaf(taps[j], cr[i0 + j], v);
        store(out_base + static_cast<long long>(o * fp.n_mfcc + c) * fp.out_frames + to, v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
template <int DT, bool ASYNC>
static cudaError_t launch_frames_dt(const FParams& fp, int grid, int smem_bytes, cudaStream_t stream) {
  static int granted[kMaxDevices] = {0};       // the kernel also has static shared memory: always ask for what is needed
  const cudaError_t eg = ensure_dyn_smem(frames512_kernel<DT, ASYNC>, smem_bytes, 0, granted);
  if (eg != cudaSuccess) return eg;
  frames512_kernel<DT, ASYNC><<<grid, kFrAllThreads, smem_bytes, stream>>>(fp);
  return cudaGetLastError();
}

template <int NC4>
static cudaError_t launch_cep_n(const FParams& fp, dim3 grid, int smem_bytes, cudaStream_t stream) {
  static int granted[kMaxDevices] = {0};
  const cudaError_t eg = ensure_dyn_smem(cepstra_kernel<NC4>, smem_bytes, 48 * 1024, granted);
  if (eg != cudaSuccess) return eg;
  cepstra_kernel<NC4><<<grid, kCepThreads, smem_bytes, stream>>>(fp);
  return cudaGetLastError();
}

cudaError_t launch_frame_prefix(const FParams& fp, cudaStream_t stream) {
  frame_prefix_kernel<<<1, 1024, 0, stream>>>(fp);
  return cudaGetLastError();
}

cudaError_t launch_frames_path(const FParams& fp, int sm_count, int frames_smem_bytes, int cep_smem_bytes, int max_frames,
                               cudaStream_t stream) {
  cudaError_t e = launch_frame_prefix(fp, stream);
  if (e != cudaSuccess) return e;
  if (fp.async_stage) {
    switch (fp.dtype) {
      case ASR_I16: e = launch_frames_dt<ASR_I16, true>(fp, sm_count, frames_smem_bytes, stream); break;
      case ASR_F32: e = launch_frames_dt<ASR_F32, true>(fp, sm_count, frames_smem_bytes, stream); break;
      default: e = launch_frames_dt<ASR_F64, true>(fp, sm_count, frames_smem_bytes, stream); break;
    }
  } else {
    switch (fp.dtype) {
      case ASR_I16: e = launch_frames_dt<ASR_I16, false>(fp, sm_count, frames_smem_bytes, stream); break;
      case ASR_F32: e = launch_frames_dt<ASR_F32, false>(fp, sm_count, frames_smem_bytes, stream); break;
      default: e = launch_frames_dt<ASR_F64, false>(fp, sm_count, frames_smem_bytes, stream); break;
    }
  }
  if (e != cudaSuccess) return e;
  const int half = (fp.delta_orders > 0 && !fp.logmel_only) ? fp.delta_width / 2 : 0;
  const int tile = kCepThreads - 2 * half;
  const int span = max(max_frames, fp.out_frames);
  const dim3 grid(fp.n_clips, (span + tile - 1) / tile);
  switch ((fp.n_mfcc + 3) / 4) {
    case 1: return launch_cep_n<1>(fp, grid, cep_smem_bytes, stream);
    case 2: return launch_cep_n<2>(fp, grid, cep_smem_bytes, stream);
    case 3: return launch_cep_n<3>(fp, grid, cep_smem_bytes, stream);
    case 4: return launch_cep_n<4>(fp, grid, cep_smem_bytes, stream);
    case 5: return launch_cep_n<5>(fp, grid, cep_smem_bytes, stream);
    case 6: return launch_cep_n<6>(fp, grid, cep_smem_bytes, stream);
    case 7: return launch_cep_n<7>(fp, grid, cep_smem_bytes, stream);
    case 8: return launch_cep_n<8>(fp, grid, cep_smem_bytes, stream);
    case 9: return launch_cep_n<9>(fp, grid, cep_smem_bytes, stream);
    case 10: return launch_cep_n<10>(fp, grid, cep_smem_bytes, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace asr
