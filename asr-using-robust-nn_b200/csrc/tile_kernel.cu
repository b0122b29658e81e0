// TMA-staged block pipeline for n_fft = 512 on sm_100a ("tiles" path): frame prefix -> tile512_kernel -> cepstra_t_kernel.
//
//   tile512_kernel   persistent, one 768-thread CTA per SM.  The frames of ALL clips form one flat list; a CTA owns a
//                    contiguous range of it and walks it in blocks of 32 frames.  Two kinds of warps:
//                    8 HELPER warps (producer side, two blocks ahead of the FFT)
//                      descriptors  which clips / frames make up block i+2, where their samples go
//                      copy         one lane per run issues cp.async.bulk (TMA) copies of the block's raw samples and of
//                                   their float64 noise into shared memory; completion is an mbarrier transaction count
//                      convert      raw -> float32 frame samples of block i+1, ONCE per sample, into the other of two
//                                   sample buffers: dtype decode, [noise mix], reflect / zero padding
//                    16 MAIN warps
//                      fft          2 frames per warp: window, 256-point complex FFT in registers, unpack, |X|^2 ->
//                                   S[slot][bin] (the exchange buffer of a slot is reused as its spectrum row)
//                      mel          lanes <-> frames, warps <-> bin ranges: pieces of mel segments, every bin feeds the
//                                   falling slope of one Slaney triangle and the rising slope of the next; partial sums
//                                   per (segment, piece) go to shared memory
//                      combine      lanes <-> frames, warps <-> filters: partials -> 10*log10 -> lm[filter][flat frame]
//                                   (transposed: stores and the cepstra kernel's loads are coalesced along time)
//                    Per block: main {combine(i-1), fft(i)} | named barrier | {mel(i)} ; helpers {descriptors(i+2),
//                    convert(i+1), copies(i+2)} ; one __syncthreads of all 24 warps.  The conversion (shared-memory and
//                    conversion-pipe work) runs beside the FFTs (FP32 issue slots) instead of between them.
//   cepstra_t_kernel per clip tile: clip maximum of the log-mel matrix (power_to_db's top_db clamp is clip-wide),
//                    clamp, DCT-II (ortho) x lifter, [delta, delta-delta], truncate / zero-pad to out_frames.
//
// Arithmetic restated from librosa.feature.mfcc (oracle/librosa_ref.py); call sites replaced:
// VDR/extract_features_construct_dataset.py:30, VDR/attacks.py:114,267 (and the SR twins for even n_fft).
#include <cstring>
#include "common.cuh"
#include "fft_core.cuh"
#include "sample_access.cuh"

// -DASR_TILE_DEBUG (make debug -> libasr_b200_dbg.so, loaded with ASR_B200_LIB): device-side checks of the staging
// protocol - raw-buffer bounds of every bulk copy, bytes announced to the mbarrier against bytes issued, sample-buffer
// bounds of every run, descriptor ring / slot ranges.  A failed check traps (the launch fails, nothing hangs).
#ifdef ASR_TILE_DEBUG
#define TL_CHECK(cond) do { if (!(cond)) { printf("tile512_kernel check failed: %s (line %d, CTA %d, thread %d)\n", #cond, __LINE__, blockIdx.x, threadIdx.x); __trap(); } } while (0)
#else
#define TL_CHECK(cond) do { } while (0)
#endif

namespace asr {

constexpr int kTlHelpDiv = 16;    // helper threads per main warp: 8 -> NW/4 helper warps, 16 -> NW/2
constexpr int kTlCache = 32;      // clips whose metadata the descriptor warp caches in shared memory
constexpr int kTlRing = 4;        // block descriptors alive at once: i-1 (combine) .. i+2 (being assembled)

struct __align__(16) TRun {     // 64 bytes, read back as four int4
  long long base;   // element offset of the clip
  double sig;       // sigma of the clip (white noise)
  int L;            // clip length
  int o0;           // ORIGINAL sample index of the first staged sample (t0 * hop - pad, may be negative)
  int count;        // staged samples: (n - 1) * hop + 512
  int aud0;         // staged position of the run's first sample
  int nu;           // 4-sample units of the run (count / 4); 0: the run is staged sample by sample from global memory
  int ra, rb;       // original samples [ra, rb) are copied raw (ra is a multiple of 8)
  int raw_a;        // byte offset of the raw audio inside the raw buffer (16-byte aligned)
  int raw_z;        // byte offset of the raw noise
  int slot0;        // first frame slot of the run
  int pad0, pad1;
};
static_assert(sizeof(TRun) == 64, "TRun is read as four int4");
struct __align__(16) TBlock {
  int n_runs, n_slots, g0, tx_bytes;   // g0: flattened index of slot 0; tx_bytes: bytes the TMA copies deliver
  TRun run[kTlMaxRuns];
  int slot_aud[32];                    // staged position of the slot's first sample (0 for empty slots)
};
struct TMeta { long long off; double sig; int L; int T; };

// ---- mbarrier / TMA (1-D bulk copy) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, const int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, const unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, const unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, const unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float tl_log2(const float x) {     // x >= amin > 0
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------------------------------------------
// 2 consecutive samples from the raw copy -> float32 frame samples (lanes take consecutive pairs: every shared-memory
// access of the conversion is a dense, conflict-free row).
// int16 stays UNSCALED (the window table carries the exact 2^-15): bits(2^23 + (s + 32768)) - (2^23 + 32768) = s.
// Noise (VDR/attacks.py:241-244): the reference forms float64(x) + sigma*z with two float64 roundings and hands that
// signal to librosa; the frame sample is its float32 rounding.  The kernel performs the same two float64 operations
// (__dmul_rn, __dadd_rn: no FMA contraction) and one conversion, so the staged sample equals float32(reference signal)
// bit for bit (asr_plan_set_stage_probe reads them back; tests/test_tiles_gpu.py).  For int16 the whole chain runs in
// int16 units (`sig` = sigma * 2^15, an exact scaling, so fl64(s + fl64(sig z)) = 2^15 fl64(x + fl64(sigma z))).
template <int DT, bool NOISE>
__device__ __forceinline__ float2 convert2(const char* __restrict__ pa, const char* __restrict__ pz, const double sig) {
  if constexpr (DT == ASR_I16) {
    const unsigned w = *reinterpret_cast<const unsigned*>(pa) ^ 0x80008000u;      // biased halves u = s + 32768
    if constexpr (NOISE) {
      // exact double(s) without a conversion instruction: bits(2^52 + u) - (2^52 + 32768); `sig` carries the 2^15
      const double2 z = *reinterpret_cast<const double2*>(pz);
      const double s0 = __hiloint2double(0x43300000, static_cast<int>(w & 0xFFFFu)) - 4503599627403264.0;
      const double s1 = __hiloint2double(0x43300000, static_cast<int>(w >> 16)) - 4503599627403264.0;
      return make_float2(static_cast<float>(__dadd_rn(s0, __dmul_rn(sig, z.x))), static_cast<float>(__dadd_rn(s1, __dmul_rn(sig, z.y))));
    }
    return make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7410)) - 8421376.0f,
                       __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7432)) - 8421376.0f);
  } else if constexpr (DT == ASR_F32) {
    const float2 v = *reinterpret_cast<const float2*>(pa);
    if constexpr (NOISE) {
      const double2 z = *reinterpret_cast<const double2*>(pz);
      return make_float2(static_cast<float>(__dadd_rn(static_cast<double>(v.x), __dmul_rn(sig, z.x))),
                         static_cast<float>(__dadd_rn(static_cast<double>(v.y), __dmul_rn(sig, z.y))));
    }
    return v;
  } else {
    const double2 a = *reinterpret_cast<const double2*>(pa);
    if constexpr (NOISE) {
      const double2 z = *reinterpret_cast<const double2*>(pz);
      return make_float2(static_cast<float>(__dadd_rn(a.x, __dmul_rn(sig, z.x))), static_cast<float>(__dadd_rn(a.y, __dmul_rn(sig, z.y))));
    }
    return make_float2(static_cast<float>(a.x), static_cast<float>(a.y));
  }
}

// one sample (clip edges): original index o, pa/pz point at original sample 0 of the raw copy
template <int DT, bool NOISE>
__device__ __forceinline__ float convert1(const char* __restrict__ pa, const char* __restrict__ pz, const int o, const double sig) {
  double xd;
  if constexpr (DT == ASR_I16) {
    const short s = reinterpret_cast<const short*>(pa)[o];                  // unscaled
    if constexpr (!NOISE) return static_cast<float>(s);
    xd = static_cast<double>(s);
  } else if constexpr (DT == ASR_F32) {
    const float x = reinterpret_cast<const float*>(pa)[o];
    if constexpr (!NOISE) return x;
    xd = static_cast<double>(x);
  } else {
    xd = reinterpret_cast<const double*>(pa)[o];
    if constexpr (!NOISE) return static_cast<float>(xd);
  }
  return static_cast<float>(__dadd_rn(xd, __dmul_rn(sig, reinterpret_cast<const double*>(pz)[o])));
}

// ------------------------------------------------------------------------------------------------
// Second pass of the 16 x 16 decomposition with the inter-pass twiddle merged into the butterflies ("twisted" DIT):
//   Z[l + 16 k1] = sum_n1 A[n1] W256^(n1 (l + 16 k1))     (A = pass-1 output of residue k2 = l, no twiddle applied)
// is a radix-2 DIT whose stage-m butterfly j uses  W_{16m}^(l + 16 j) = W_{16m}^l * W_m^j  in place of W_m^j: one rounded
// table value per butterfly instead of a twiddle multiply followed by a butterfly, and j -> j + m/4 is a factor -i, so
// a lane keeps 1 + 1 + 2 + 4 = 8 complex values in registers for the whole kernel (tw[0]: m = 2, tw[1]: m = 4,
// tw[2..3]: m = 8, tw[4..7]: m = 16).  32 butterflies x 6 FMA; no table loads.  Input bit-reversed, output natural.
__device__ __forceinline__ void dft16_twisted(float (&re)[16], float (&im)[16], const float (&twr)[8], const float (&twi)[8]) {
#pragma unroll
  for (int m = 2; m <= 16; m *= 2) {
    const int h = m / 2, base = (m == 2) ? 0 : (m == 4 ? 1 : (m == 8 ? 2 : 4)), nq = (h > 1) ? h / 2 : 1;
#pragma unroll
    for (int g = 0; g < 16; g += m) {
#pragma unroll
      for (int j = 0; j < h; ++j) {
        const int a = g + j, b = a + h;
        const float ar = re[a], ai = im[a], br = re[b], bi = im[b];
        float nr, ni;
        if (j < nq) {                               // w = tw
          const float wr = twr[base + j], wi = twi[base + j];
          nr = fmaf(-wi, bi, fmaf(wr, br, ar));
          ni = fmaf(wi, br, fmaf(wr, bi, ai));
        } else {                                    // w = -i tw = (wi, -wr)
          const float wr = twr[base + j - nq], wi = twi[base + j - nq];
          nr = fmaf(wr, bi, fmaf(wi, br, ar));
          ni = fmaf(-wr, br, fmaf(wi, bi, ai));
        }
        re[a] = nr; im[a] = ni;
        re[b] = fmaf(2.0f, ar, -nr);
        im[b] = fmaf(2.0f, ai, -ni);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Power spectrum of one real frame of 512 staged samples by a group of 16 lanes: 256-point complex FFT of
// z[n] = x[2n] + i x[2n+1] (n = n1 + 16 n2, k = k2 + 16 k1; lane = n1 in pass 1, k2 in pass 2), then the real-input
// unpack.  Differences from frame_power_fft<512> (fft_core.cuh): the window multiply is fused into the first
// radix-2 stage (p = xa*wa; p +- xb*wb), the inter-pass twiddles live in the second pass's butterflies (dft16_twisted),
// and the unpack works on 2X: X' = (A + conj B) + w'(A - conj B), Y' = 2(A + conj B) - X', with w' = 2w; the spectrum
// row holds 4|X|^2 and the mel weights of this path carry the 1/4.
__device__ __forceinline__ void tile_fft512(const float2* __restrict__ xs2, const float2* __restrict__ win2,
                                            const float (&twr)[8], const float (&twi)[8], const float2* __restrict__ twu2,
                                            float* buf, const int l) {
  constexpr int M = 256, G = 16, P = 16;
  float re[P], im[P];
#pragma unroll
  for (int g = 0; g < P; g += 2) {                  // first stage: bit-reversed inputs g, g+1 <-> n2 = brev(g), brev(g) + 8
    const int na = l + 16 * brev<16>(g), nb = na + 128;
    const float2 xa = xs2[na], xb = xs2[nb];
    const float2 wa = win2[na], wb = win2[nb];
    const float pr = xa.x * wa.x, pi = xa.y * wa.y;
    re[g] = fmaf(xb.x, wb.x, pr); re[g + 1] = fmaf(-xb.x, wb.x, pr);
    im[g] = fmaf(xb.y, wb.y, pi); im[g + 1] = fmaf(-xb.y, wb.y, pi);
  }
  dft_dit_from<P, 4>(re, im);
  float2* xb2 = reinterpret_cast<float2*>(buf);
  float ur[G], ui[G];
  __syncwarp();
#pragma unroll
  for (int k2 = 0; k2 < P; ++k2) xb2[l * (P + 1) + k2] = make_float2(re[k2], im[k2]);
  __syncwarp();
#pragma unroll
  for (int n1 = 0; n1 < G; ++n1) {
    const float2 a = xb2[n1 * (P + 1) + l];
    ur[brev<G>(n1)] = a.x;
    ui[brev<G>(n1)] = a.y;
  }
  dft16_twisted(ur, ui, twr, twi);
  __syncwarp();                                    // exchange data consumed; buf becomes the spectrum row
  const int partner = (G - l) & (G - 1);
#pragma unroll
  for (int k1 = 0; k1 < G / 2; ++k1) {
    const int k = l + P * k1;
    // Z[M-k] lives in lane G-l at G-1-k1; lane 0 pairs with itself at (G-k1) mod G.  Every lane offers what its requester needs.
    const int j0 = (G - k1) & (G - 1);
    const float give_r = (l == 0) ? ur[j0] : ur[G - 1 - k1];
    const float give_i = (l == 0) ? ui[j0] : ui[G - 1 - k1];
    const float br = __shfl_sync(0xffffffffu, give_r, partner, G);
    const float bi = __shfl_sync(0xffffffffu, give_i, partner, G);
    const float2 w = twu2[k];                      // (-sin(2 pi k/N), -cos(2 pi k/N))
    const float ar = ur[k1], ai = ui[k1];
    const float sr = ar + br, si = ai - bi;        // A + conj(B)
    const float dr = ar - br, di = ai + bi;        // A - conj(B)
    const float xr = fmaf(-w.y, di, fmaf(w.x, dr, sr));
    const float xi = fmaf(w.y, dr, fmaf(w.x, di, si));
    const float yr = fmaf(2.0f, sr, -xr), yi = fmaf(2.0f, si, -xi);
    buf[k] = fmaf(xr, xr, xi * xi);
    buf[M - k] = fmaf(yr, yr, yi * yi);
  }
  if (l == 0) buf[M / 2] = 4.0f * fmaf(ur[G / 2], ur[G / 2], ui[G / 2] * ui[G / 2]);
}

// ------------------------------------------------------------------------------------------------
// NW main warps per CTA (FFT, mel, combine; FR = 2*NW frames per block) + NW/4 helper warps (descriptors, TMA copies,
// sample conversion into the other of two sample buffers), so that the conversion - shared-memory and conversion-pipe
// work - runs beside the FFTs - FP32 issue slots - instead of between them.  NW = 16: one CTA per SM; NW = 8: two.
template <int DT, bool NOISE, int NW>
__global__ void __launch_bounds__(NW * 32 + NW * kTlHelpDiv, NW == 8 ? 2 : 1) tile512_kernel(const __grid_constant__ FParams fp) {
  constexpr int kTlThreads = NW * 32, kTlBlock = 2 * NW, kTlWarps = NW;
  constexpr int kHelpWarps = NW * kTlHelpDiv / 32, kAllThreads = kTlThreads + 32 * kHelpWarps;
  extern __shared__ __align__(16) float smem[];
  __shared__ TBlock ring[kTlRing];
  __shared__ TMeta s_meta[kTlCache];
  __shared__ __align__(8) unsigned long long s_bar;
  constexpr int esz = DT == ASR_I16 ? 2 : (DT == ASR_F32 ? 4 : 8);
  constexpr int kAsmWarp = kTlWarps;                // first helper warp: assembles descriptors and issues the copies
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);   // the same value, known to be warp-uniform
  const bool helper = warp_u >= kTlWarps;
  const int hw = warp_u - kTlWarps, ht = tid - kTlThreads;   // helper warp / thread index

  // ---- this CTA's range of the flattened frame list ----
  const int total = __ldg(fp.fstart + fp.n_clips);
  const int n_blocks = (total + kTlBlock - 1) / kTlBlock;
  const int per = (n_blocks + gridDim.x - 1) / gridDim.x;
  const long long gb = static_cast<long long>(blockIdx.x) * per * kTlBlock;
  if (gb >= total) return;
  const int g_begin = static_cast<int>(gb);
  const int g_end = static_cast<int>(min(static_cast<long long>(total), gb + static_cast<long long>(per) * kTlBlock));

  // ---- tables: global blob -> shared (common tables, then the window that matches the staged sample scale) ----
  {
    float4* dst = reinterpret_cast<float4*>(smem);
    for (int i = tid; i < fp.blob_f4; i += kAllThreads) dst[i] = __ldg(fp.blob + i);
    if (tid < 128) dst[fp.blob_f4 + tid] = __ldg(fp.blob + fp.off_window / 4 + tid);
  }
  const float2* s_win2 = reinterpret_cast<const float2*>(smem + 4 * fp.blob_f4);
  const float2* s_twu = reinterpret_cast<const float2*>(smem + fp.off_twu);
  const float4* s_wtab = reinterpret_cast<const float4*>(smem + fp.off_wtab);
  const int4* s_pieces = reinterpret_cast<const int4*>(smem + fp.off_steps);
  float* s_aud = smem + fp.sm_aud;                   // [2][aud_cap]
  float* s_S = smem + fp.sm_S;                       // [32 slots][kTlRS]
  float* s_part = smem + fp.sm_part;                 // [t_npart][FR]
  char* s_raw = reinterpret_cast<char*>(smem + fp.sm_raw);
  for (int i = tid; i < fp.t_npart * kTlBlock; i += kAllThreads) s_part[i] = 0.0f;   // rows of pieces that do not exist stay 0
  if (tid == 0) mbar_init(&s_bar, 1);

  // ---- block cursor (warp kAsmWarp; lane 0 holds the live copy) ----
  int b_cur = 0, t_cur = 0, g_cur = g_begin, cache_base = 0;
  auto load_meta = [&](const int base) {               // lanes of the descriptor warp
#pragma unroll
    for (int e = lane; e < kTlCache; e += 32) {
      const int b = min(base + e, fp.n_clips - 1);
      TMeta m;
      m.off = __ldg(fp.offsets + b);
      m.sig = NOISE ? __ldg(fp.sigma + b) : 0.0;
      m.L = __ldg(fp.lengths + b);
      m.T = __ldg(fp.nframes + b);
      s_meta[e] = m;
    }
  };
  auto assemble = [&](TBlock& blk) {                   // whole descriptor warp
    int n_slots = 0, n_runs = 0, aud = 0, raw_off = 0, tx = 0;
    const int g0 = g_cur;
    for (;;) {
      int need = 0;
      if (lane == 0) {
        while (n_runs < fp.max_runs && n_slots < kTlBlock && g_cur < g_end && b_cur < fp.n_clips) {
          if (b_cur >= cache_base + kTlCache) { need = 1; break; }
          const TMeta cm = s_meta[b_cur - cache_base];
          if (t_cur < cm.T) {
            const int n = min(min(kTlBlock - n_slots, cm.T - t_cur), g_end - g_cur);
            TRun& run = blk.run[n_runs];
            run.base = cm.off;
            run.sig = cm.sig;
            run.L = cm.L;
            run.o0 = t_cur * fp.hop - fp.pad;
            run.count = (n - 1) * fp.hop + 512;
            run.aud0 = (aud + 3) & ~3;
            aud = run.aud0 + run.count;
            run.slot0 = n_slots;
            const bool fast = fp.vec_ok && (cm.off & 7) == 0;
            run.nu = fast ? run.count >> 2 : 0;
            run.ra = 0; run.rb = 0; run.raw_a = 0; run.raw_z = 0;
            if (fast) {
              const int o0 = run.o0, o1 = o0 + run.count;
              int ra = max(0, o0), rb = min(cm.L, o1);
              if (fp.pad_mode == ASR_PAD_REFLECT) {       // sources of the reflected samples (edge frames)
                if (o0 < 0) rb = max(rb, min(cm.L, 1 - o0));
                if (o1 > cm.L) ra = min(ra, max(0, 2 * (cm.L - 1) - (o1 - 1)));
              }
              ra &= ~7;                                   // 16-byte aligned source for every dtype
              run.ra = ra; run.rb = rb;
              run.raw_a = raw_off;
              raw_off += ((rb - ra) * esz + 15) & ~15;
              run.raw_z = raw_off;
              if (NOISE) raw_off += ((rb - ra) * 8 + 15) & ~15;
              tx += max(0, (rb & ~7) - ra) * (esz + (NOISE ? 8 : 0));
            }
            n_slots += n; t_cur += n; g_cur += n; ++n_runs;
          }
          if (t_cur >= cm.T) { ++b_cur; t_cur = 0; }      // clip finished (or without frames): the next one continues the block
        }
      }
      need = __shfl_sync(0xffffffffu, need, 0);
      if (!need) break;
      cache_base = __shfl_sync(0xffffffffu, b_cur, 0);
      __syncwarp();
      load_meta(cache_base);
      __syncwarp();
    }
    if (lane == 0) {
      blk.n_runs = n_runs; blk.n_slots = n_slots; blk.g0 = g0; blk.tx_bytes = tx;
      TL_CHECK(n_runs <= fp.max_runs && n_runs <= kTlMaxRuns && n_slots <= kTlBlock);
      TL_CHECK(aud <= fp.aud_cap);                                   // the block's samples fit one sample buffer
      TL_CHECK(fp.sm_raw * 4 + raw_off <= fp.t_smem_bytes);          // the block's raw bytes fit the raw buffer
      TL_CHECK(g0 >= g_begin && g0 + n_slots <= g_end);
    }
    __syncwarp();
    const int nr = blk.n_runs, ns = blk.n_slots;
    int a = 0;                                          // (lanes >= kTlBlock: ns <= kTlBlock keeps them at 0)
    if (lane < ns) {
#pragma unroll
      for (int r = 0; r < kTlMaxRuns; ++r)
        if (r < nr && lane >= blk.run[r].slot0) a = blk.run[r].aud0 + (lane - blk.run[r].slot0) * fp.hop;
    }
    blk.slot_aud[lane] = a;
  };
  // raw copies of a block: one bulk copy per array and run (lanes <-> runs), element copies for the < 8 samples past
  // the last 16-byte boundary, then ONE arrival carrying the byte count
  auto issue_copies = [&](const TBlock& blk) {         // whole descriptor warp
    const int nr = blk.n_runs;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#ifdef ASR_TILE_DEBUG
    {
      int issued = 0;
      if (lane < nr && blk.run[lane].nu > 0) issued = max(0, (blk.run[lane].rb & ~7) - blk.run[lane].ra) * (esz + (NOISE ? 8 : 0));
      for (int o = 16; o > 0; o >>= 1) issued += __shfl_xor_sync(0xffffffffu, issued, o);
      TL_CHECK(issued == blk.tx_bytes);                              // what the mbarrier expects is what the copies deliver
    }
#endif
    if (lane < nr && blk.run[lane].nu > 0) {
      const TRun& run = blk.run[lane];
      const int nfull = (run.rb & ~7) - run.ra;
      TL_CHECK(run.raw_a >= 0 && (run.raw_a & 15) == 0 && (run.ra & 7) == 0 && run.rb <= run.L && run.ra <= run.rb);
      TL_CHECK(fp.sm_raw * 4 + run.raw_a + max(nfull, 0) * esz <= fp.t_smem_bytes);
      if (nfull > 0) {
        tma_bulk_g2s(s_raw + run.raw_a, reinterpret_cast<const char*>(fp.audio) + (run.base + run.ra) * esz,
                     static_cast<unsigned>(nfull * esz), &s_bar);
        if (NOISE) tma_bulk_g2s(s_raw + run.raw_z, fp.z + run.base + run.ra, static_cast<unsigned>(nfull * 8), &s_bar);
      }
    }
    for (int r = 0; r < nr; ++r) {
      const TRun& run = blk.run[r];
      const int i = max(run.ra, run.rb & ~7) + lane;
      if (run.nu > 0 && lane < 8 && i < run.rb) {
        const long long e = run.base + i;
        if constexpr (DT == ASR_I16) reinterpret_cast<short*>(s_raw + run.raw_a)[i - run.ra] = __ldg(reinterpret_cast<const short*>(fp.audio) + e);
        else if constexpr (DT == ASR_F32) reinterpret_cast<float*>(s_raw + run.raw_a)[i - run.ra] = __ldg(reinterpret_cast<const float*>(fp.audio) + e);
        else reinterpret_cast<double*>(s_raw + run.raw_a)[i - run.ra] = __ldg(reinterpret_cast<const double*>(fp.audio) + e);
        if (NOISE) reinterpret_cast<double*>(s_raw + run.raw_z)[i - run.ra] = __ldg(fp.z + e);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_expect_tx(&s_bar, static_cast<unsigned>(blk.tx_bytes));
  };
  // raw -> float32 frame samples of a block, once per sample
  const float scale = DT == ASR_I16 ? 32768.0f : 1.0f;   // int16 is staged unscaled (the slow path computes scaled values)
  const float inv_scale = DT == ASR_I16 ? (1.0f / 32768.0f) : 1.0f;
  float* const probe = fp.stage_probe;                   // parity probe: staged samples back to global memory, packed like the audio
  auto convert_block = [&](const TBlock& blk, float* aud) {       // helper warps
    const int nr = blk.n_runs;
    for (int r = 0; r < nr; ++r) {
      const int4* rq = reinterpret_cast<const int4*>(&blk.run[r]);   // the descriptor in four 16-byte loads
      const int4 q0 = rq[0], q1 = rq[1], q2 = rq[2], q3 = rq[3];
      const long long base = (static_cast<long long>(q0.y) << 32) | static_cast<unsigned>(q0.x);
      const double sig = __hiloint2double(q0.w, q0.z);
      const double sigs = sig * scale;                   // int16 is mixed in int16 units (exact scaling)
      const int L = q1.x, o0 = q1.y, count = q1.z, nu = q2.x, ra = q2.y;
      float* dst = aud + q1.w;
      if (nu == 0) {                                    // clip not on a 16-byte boundary: sample by sample from global memory
        for (int i = ht; i < count; i += 32 * kHelpWarps) {
          const float v = padded_at<DT>(fp, base, L, o0 + fp.pad + i, sig);
          dst[i] = v * scale;
          if (probe && o0 + i >= 0 && o0 + i < L) probe[base + o0 + i] = v;
        }
        continue;
      }
      const char* pa = s_raw + q2.w - ra * esz;          // original sample o at pa + o*esz
      const char* pz = s_raw + q3.x - ra * 8;
      // a warp takes 128 consecutive samples per round: lane l the pairs at 2l and 64 + 2l
      for (int c0 = 128 * hw; c0 < count; c0 += 128 * kHelpWarps) {
        const int og = o0 + c0;                         // original index of the round's first sample
        if (og >= 0 && og + 128 <= L && c0 + 128 <= count) {       // (warp-uniform) all 128 samples inside the clip
          const char* qa = pa + (og + 2 * lane) * esz;
          const char* qz = pz + (og + 2 * lane) * 8;
          const float2 v0 = convert2<DT, NOISE>(qa, qz, sigs);
          const float2 v1 = convert2<DT, NOISE>(qa + 64 * esz, qz + 512, sigs);
          float2* qd = reinterpret_cast<float2*>(dst + c0) + lane;
          qd[0] = v0;
          qd[32] = v1;
          if (probe) {
            float* pp = probe + base + og + 2 * lane;
            pp[0] = v0.x * inv_scale; pp[1] = v0.y * inv_scale;
            pp[64] = v1.x * inv_scale; pp[65] = v1.y * inv_scale;
          }
          continue;
        }
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {                   // clip edges and the last round of the run
          const int i = c0 + 64 * h + 2 * lane;
          if (i < count) {
            const int orig = o0 + i;
            float e[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              int o = orig + j;
              bool zero = false;
              if (o < 0) { zero = fp.pad_mode != ASR_PAD_REFLECT; o = -o; }
              else if (o >= L) { zero = fp.pad_mode != ASR_PAD_REFLECT; o = 2 * (L - 1) - o; }
              e[j] = zero ? 0.0f : convert1<DT, NOISE>(pa, pz, o, sigs);
              if (probe && o == orig + j) probe[base + o] = e[j] * inv_scale;     // samples of the clip itself (not their reflections)
            }
            *reinterpret_cast<float2*>(dst + i) = make_float2(e[0], e[1]);
          }
        }
      }
    }
  };

  // named barriers: 1 = main warps (between FFT and mel), 2 = helper warps; __syncthreads = everybody, once per block
  auto bar_main = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kTlThreads) : "memory"); };
  auto bar_help = [&]() { asm volatile("bar.sync 2, %0;" ::"n"(32 * kHelpWarps) : "memory"); };

  // ---- prologue: descriptors of blocks 0 and 1, samples of block 0 ----
  if (warp_u == kAsmWarp) {
    int lo = 0, hi = fp.n_clips;                       // largest b with fstart[b] <= g_begin
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(fp.fstart + mid) <= g_begin) lo = mid; else hi = mid;
    }
    b_cur = lo;
    t_cur = g_begin - __ldg(fp.fstart + lo);
    cache_base = lo;
    load_meta(cache_base);
    __syncwarp();
    assemble(ring[0]);
    assemble(ring[1]);
  }
  __syncthreads();                                     // tables, zeroed partials, mbarrier, descriptors 0 and 1
  if (helper) {
    if (warp_u == kAsmWarp) issue_copies(ring[0]);
    mbar_wait(&s_bar, 0);
    convert_block(ring[0], s_aud);
    bar_help();
    if (warp_u == kAsmWarp && ring[1].n_slots > 0) issue_copies(ring[1]);
  }
  __syncthreads();

  if (helper) {
    // ---- helper warps: block it+1 raw -> sample buffer (it+1)&1 while the main warps transform block it;
    //      descriptor of block it+2, then its copies as soon as the raw buffer is free again ----
    for (int it = 0;; ++it) {
      if (ring[it & (kTlRing - 1)].n_slots == 0) break;
      const TBlock& nxt = ring[(it + 1) & (kTlRing - 1)];
      TBlock& nn = ring[(it + 2) & (kTlRing - 1)];
      if (warp_u == kAsmWarp) assemble(nn);
      if (nxt.n_slots > 0) {
        mbar_wait(&s_bar, static_cast<unsigned>((it + 1) & 1));
        convert_block(nxt, s_aud + ((it + 1) & 1) * fp.aud_cap);
      }
      bar_help();                                      // every helper is done with the raw buffer
      if (warp_u == kAsmWarp && nn.n_slots > 0) issue_copies(nn);
      __syncthreads();
    }
    return;
  }

  // ---- main warps: per-thread constants of the phases ----
  const int fft_h = lane >> 4, fft_l = lane & 15;
  float twr[8], twi[8];                                            // this lane's second-pass twiddles (dft16_twisted)
  {
    const float4* t4 = reinterpret_cast<const float4*>(smem + fp.off_twp + 16 * fft_l);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 t = t4[q];
      twr[2 * q] = t.x; twi[2 * q] = t.y; twr[2 * q + 1] = t.z; twi[2 * q + 1] = t.w;
    }
  }
  const int fft_slot0 = 8 * (warp >> 2) + (warp & 3);            // half-warps 4 slots apart: complementary bank halves of S
  const int fft_slot = fft_slot0 + 4 * fft_h;
  float* const fft_buf = s_S + fft_slot * kTlRS;
  const int fr = tid & (kTlBlock - 1), vw = tid / kTlBlock;      // frame slot and "virtual warp" of the mel / combine phases
  const int2 mel_range = reinterpret_cast<const int2*>(smem + fp.off_wrange)[vw];   // (first piece, pieces)
  const float4* mel_S = reinterpret_cast<const float4*>(s_S + fr * kTlRS);
  const int npc = fp.t_npc;
  float* const lm_fr = fp.lm + fr;
  char* const part_fr = reinterpret_cast<char*>(s_part + fr);

  for (int it = 0;; ++it) {
    const TBlock& cur = ring[it & (kTlRing - 1)];
    // ---- combine (block it-1): lanes <-> frames, virtual warps <-> filters ----
    if (it > 0) {
      const TBlock& prv = ring[(it - 1) & (kTlRing - 1)];
      const int ns = prv.n_slots, g0 = prv.g0;
      for (int j = vw; j < fp.n_mels; j += kTlThreads / kTlBlock) {
        // rise rows of segment j: p[c * 2FR]; fall rows of segment j+1: p[(2 npc - 1 + 2c) * FR]
        const float* p = s_part + ((j * npc) * 2 + 1) * kTlBlock + fr;
        const float* q = p + (2 * npc - 1) * kTlBlock;
        float m = p[0] + q[0];
        if (npc > 1) m += p[2 * kTlBlock] + q[2 * kTlBlock];
        if (npc > 2) m += p[4 * kTlBlock] + q[4 * kTlBlock];
        if (npc > 3) m += p[6 * kTlBlock] + q[6 * kTlBlock];
        const float db = 3.01029995663981195f * tl_log2(fmaxf(fp.amin, m));
        if (fr < ns) lm_fr[static_cast<long long>(j) * fp.lm_stride + g0] = db;
      }
    }
    if (cur.n_slots == 0) break;                       // uniform over the CTA
    // ---- fft (block it) ----
    TL_CHECK(cur.n_slots <= kTlBlock && cur.slot_aud[fft_slot] >= 0 && cur.slot_aud[fft_slot] + 512 <= fp.aud_cap);
    if (fft_slot0 < cur.n_slots)
      tile_fft512(reinterpret_cast<const float2*>(s_aud + (it & 1) * fp.aud_cap + cur.slot_aud[fft_slot]), s_win2, twr, twi, s_twu,
                  fft_buf, fft_l);
    bar_main();
    // ---- mel (block it): lanes <-> frames, this virtual warp's pieces ----
    {
      const int4* pp = s_pieces + mel_range.x;
      const char* Sb = reinterpret_cast<const char*>(mel_S);
#pragma unroll 1
      for (int n = mel_range.y; n > 0; --n, ++pp) {
        const int4 pc = *pp;                           // (byte offset of the first 4 bins in the S row, steps, byte offset of the fall partial row, first weight float4)
        const float4* sq = reinterpret_cast<const float4*>(Sb + pc.x);
        const float4* wp = s_wtab + pc.w;
        float a0 = 0.0f, a1 = 0.0f, b0 = 0.0f, b1 = 0.0f;
#pragma unroll 1
        for (int k = 0; k < pc.y; ++k) {
          const float4 sv = sq[k];
          const float4 w01 = wp[2 * k], w23 = wp[2 * k + 1];   // (fall, rise) of bins 0,1 and 2,3
          a0 = fmaf(w01.x, sv.x, a0); b0 = fmaf(w01.y, sv.x, b0);
          a1 = fmaf(w01.z, sv.y, a1); b1 = fmaf(w01.w, sv.y, b1);
          a0 = fmaf(w23.x, sv.z, a0); b0 = fmaf(w23.y, sv.z, b0);
          a1 = fmaf(w23.z, sv.w, a1); b1 = fmaf(w23.w, sv.w, b1);
        }
        float* pr = reinterpret_cast<float*>(part_fr + pc.z);
        pr[0] = a0 + a1;                               // falling slope of filter seg-1
        pr[kTlBlock] = b0 + b1;                        // rising slope of filter seg
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Cepstra from the transposed log-mel workspace: grid (clip, time tile).  Every CTA first takes the clip-wide maximum
// (power_to_db(top_db): max over the whole (n_mels, T) matrix of the call), then thread <-> frame: clamp, DCT-II x
// lifter from a transposed table, [Savitzky-Golay deltas through shared memory], coalesced stores along time.
constexpr int kCepTThreads = 128;

template <int NC4>   // NC4 = ceil(n_mfcc / 4), 1..10
__global__ void __launch_bounds__(kCepTThreads) cepstra_t_kernel(const __grid_constant__ FParams fp) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float s_red[kCepTThreads / 32];
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const int T = __ldg(fp.nframes + b);
  const int rows = fp.logmel_only ? fp.n_mels : fp.out_rows;
  const long long out_base = static_cast<long long>(b) * rows * fp.out_frames;
  const int half = (fp.delta_orders > 0 && !fp.logmel_only) ? fp.delta_width / 2 : 0;
  const int tile = kCepTThreads - 2 * half;             // output frames per CTA
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
    if (w > 0)                                           // threads <-> columns, rows in the loop: no division per element
      for (int r = 0; r < rows; ++r)
        for (int x = tid; x < w; x += kCepTThreads) store(out_base + static_cast<long long>(r) * fp.out_frames + z_lo + x, 0.0f);
  }
  if (t_lo >= t_out) return;

  float* s_dct = smem;                                  // [n_mels][4*NC4] transposed, lifter folded in
  float* s_cep = smem + fp.cep_off_cbuf;                // [n_mfcc][kCepTThreads+1]
  const float* s_taps = smem + fp.cep_off_taps;
  if (!fp.logmel_only) {
    const float4* src = fp.cep_blob;
    float4* dst = reinterpret_cast<float4*>(smem);
    for (int i = tid; i < fp.cep_tab_f4; i += kCepTThreads) dst[i] = __ldg(src + i);
  }
  const int g0 = __ldg(fp.fstart + b);
  const float* lm0 = fp.lm + g0;
  // frame handled by this thread: centre frames of the tile plus the delta halo, clamped into the clip
  const int c_lo = half > 0 ? min(max(t_lo, half), T - 1 - half) - half : t_lo;   // first frame whose cepstrum the tile needs
  const int t = c_lo + tid;
  const bool live = t < T && t >= 0;
  // ---- this thread's log-mel column -> shared memory (one pass over the workspace), clip maximum on the way ----
  float* s_col = smem + fp.cep_off_col + tid;            // [n_mels][kCepTThreads]
  float mx = -3.0e38f;
  if (live) {
    const float* col = lm0 + t;
#pragma unroll 8
    for (int j = 0; j < fp.n_mels; ++j) {
      const float v = __ldcg(col + static_cast<long long>(j) * fp.lm_stride);
      s_col[j * kCepTThreads] = v;
      mx = fmaxf(mx, v);
    }
  }
  float thr = -3.0e38f;
  if (fp.top_db >= 0.0f) {
    // frames of the clip outside this CTA's window (clips longer than one tile)
    for (int t2 = tid; t2 < T; t2 += kCepTThreads) {
      if (t2 >= c_lo && t2 < c_lo + kCepTThreads) continue;
      for (int j = 0; j < fp.n_mels; ++j) mx = fmaxf(mx, __ldcg(lm0 + static_cast<long long>(j) * fp.lm_stride + t2));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = mx;
    __syncthreads();
    mx = s_red[0];
#pragma unroll
    for (int w = 1; w < kCepTThreads / 32; ++w) mx = fmaxf(mx, s_red[w]);
    thr = mx - fp.top_db;
  } else {
    __syncthreads();                                     // the DCT table
  }

  if (fp.logmel_only) {
    if (live && t < t_out && tid < tile) {
      for (int j = 0; j < fp.n_mels; ++j)
        store(out_base + static_cast<long long>(j) * fp.out_frames + t, fmaxf(s_col[j * kCepTThreads], thr));
    }
    return;
  }

  float acc[4 * NC4];
#pragma unroll
  for (int c = 0; c < 4 * NC4; ++c) acc[c] = 0.0f;
  if (live) {
    const float4* d4 = reinterpret_cast<const float4*>(s_dct);
#pragma unroll 2
    for (int j = 0; j < fp.n_mels; ++j) {
      const float v = fmaxf(s_col[j * kCepTThreads], thr);
#pragma unroll
      for (int c4 = 0; c4 < NC4; ++c4) {
        const float4 d = d4[j * NC4 + c4];
        acc[4 * c4] = fmaf(v, d.x, acc[4 * c4]);
        acc[4 * c4 + 1] = fmaf(v, d.y, acc[4 * c4 + 1]);
        acc[4 * c4 + 2] = fmaf(v, d.z, acc[4 * c4 + 2]);
        acc[4 * c4 + 3] = fmaf(v, d.w, acc[4 * c4 + 3]);
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
  const int cp = kCepTThreads + 1;
#pragma unroll
  for (int c = 0; c < 4 * NC4; ++c)
    if (c < fp.n_mfcc) s_cep[c * cp + tid] = acc[c];
  __syncthreads();
  const int to = t_lo + tid;                             // output frame of this thread
  if (tid < tile && to < t_out) {
    const int tc = min(max(to, half), T - 1 - half);
    const int i0 = tc - half - c_lo;                     // first tap position in s_cep
    if (fp.delta_width == 9 && fp.delta_orders <= 2) {
      // librosa's default width: taps of both orders in registers, the nine cepstra of the window loaded once per
      // coefficient and used for both orders (same products in the same order as the general loop below)
      float tp[2][9];
#pragma unroll
      for (int o = 0; o < 2; ++o)
#pragma unroll
        for (int j = 0; j < 9; ++j) tp[o][j] = (o < fp.delta_orders) ? s_taps[o * 9 + j] : 0.0f;
      for (int c = 0; c < fp.n_mfcc; ++c) {
        const float* cr = s_cep + c * cp;
        float x[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) x[j] = cr[i0 + j];
        store(out_base + static_cast<long long>(c) * fp.out_frames + to, cr[to - c_lo]);
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          if (o < fp.delta_orders) {
            float v = 0.0f;
#pragma unroll
            for (int j = 0; j < 9; ++j) v = fmaf(tp[o][j], x[j], v);
            store(out_base + static_cast<long long>((o + 1) * fp.n_mfcc + c) * fp.out_frames + to, v);
          }
        }
      }
    } else {
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
}

// ------------------------------------------------------------------------------------------------
// The same for n_mels <= 32 without deltas (BASELINE C1/C2/C4): the frame's log-mel column stays in registers and the
// DCT table comes from the kernel parameters, i.e. as constant-bank operands of the FMAs - no shared memory, no table loads.
template <int NC4>
__global__ void __launch_bounds__(kCepTThreads) cepstra_small_kernel(const __grid_constant__ FParams fp) {
  __shared__ float s_red[kCepTThreads / 32];
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const int T = __ldg(fp.nframes + b);
  const long long out_base = static_cast<long long>(b) * fp.out_rows * fp.out_frames;
  const int t_lo = blockIdx.y * kCepTThreads;           // first output frame of this tile
  auto store = [&](const long long idx, const float v) {
    if (fp.out_f64) reinterpret_cast<double*>(fp.out)[idx] = static_cast<double>(v);
    else reinterpret_cast<float*>(fp.out)[idx] = v;
  };
  const int t_out = min(T, fp.out_frames);              // frames carrying data; the rest is zero padding
  {                                                     // zero padding in the feature domain (VDR/extract...py:36-37)
    const int z_lo = max(t_out, t_lo), z_hi = min(fp.out_frames, t_lo + kCepTThreads);
    const int w = z_hi - z_lo;
    if (w > 0)
      for (int r = 0; r < fp.out_rows; ++r)
        for (int x = tid; x < w; x += kCepTThreads) store(out_base + static_cast<long long>(r) * fp.out_frames + z_lo + x, 0.0f);
  }
  if (t_lo >= t_out) return;
  const int g0 = __ldg(fp.fstart + b);
  const float* lm0 = fp.lm + g0;
  const int t = t_lo + tid;
  const bool live = t < T;
  float v[32];
  float mx = -3.0e38f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    v[j] = (live && j < fp.n_mels) ? __ldcg(lm0 + static_cast<long long>(j) * fp.lm_stride + t) : -3.0e38f;
    mx = fmaxf(mx, v[j]);
  }
  float thr = -3.0e38f;
  if (fp.top_db >= 0.0f) {
    for (int t2 = tid; t2 < T; t2 += kCepTThreads) {   // frames of the clip outside this CTA's tile (clips longer than one tile)
      if (t2 >= t_lo && t2 < t_lo + kCepTThreads) continue;
      for (int j = 0; j < fp.n_mels; ++j) mx = fmaxf(mx, __ldcg(lm0 + static_cast<long long>(j) * fp.lm_stride + t2));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = mx;
    __syncthreads();
    mx = s_red[0];
#pragma unroll
    for (int w = 1; w < kCepTThreads / 32; ++w) mx = fmaxf(mx, s_red[w]);
    thr = mx - fp.top_db;
  }
  if (!live || t >= t_out) return;
  float acc[4 * NC4];
#pragma unroll
  for (int c = 0; c < 4 * NC4; ++c) acc[c] = 0.0f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < fp.n_mels) {
      const float x = fmaxf(v[j], thr);
#pragma unroll
      for (int c = 0; c < 4 * NC4; ++c) acc[c] = fmaf(x, fp.cep_dct[j * 4 * NC4 + c], acc[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < 4 * NC4; ++c)
    if (c < fp.n_mfcc) store(out_base + static_cast<long long>(c) * fp.out_frames + t, acc[c]);
}

// ------------------------------------------------------------------------------------------------
template <int DT, bool NOISE, int NW>
static cudaError_t launch_tile_nw(const FParams& fp, int sm_count, int smem_bytes, cudaStream_t stream) {
  static int granted[kMaxDevices] = {0};       // the kernel also has static shared memory: always ask for what is needed
  const cudaError_t eg = ensure_dyn_smem(tile512_kernel<DT, NOISE, NW>, smem_bytes, 0, granted, true);
  if (eg != cudaSuccess) return eg;
  tile512_kernel<DT, NOISE, NW><<<(NW == 8 ? 2 : 1) * sm_count, NW * 32 + NW * kTlHelpDiv, smem_bytes, stream>>>(fp);
  return cudaGetLastError();
}
template <int DT, bool NOISE>
static cudaError_t launch_tile_dt(const FParams& fp, int sm_count, int smem_bytes, cudaStream_t stream) {
  return fp.t_nw == 8 ? launch_tile_nw<DT, NOISE, 8>(fp, sm_count, smem_bytes, stream)
                      : launch_tile_nw<DT, NOISE, 16>(fp, sm_count, smem_bytes, stream);
}

template <int NC4>
static cudaError_t launch_cep_t(const FParams& fp, dim3 grid, int smem_bytes, cudaStream_t stream) {
  static int granted[kMaxDevices] = {0};
  const cudaError_t eg = ensure_dyn_smem(cepstra_t_kernel<NC4>, smem_bytes, 48 * 1024, granted);
  if (eg != cudaSuccess) return eg;
  cepstra_t_kernel<NC4><<<grid, kCepTThreads, smem_bytes, stream>>>(fp);
  return cudaGetLastError();
}

// cepstra stage of the transposed log-mel workspace (shared by the TILES and TC paths)
cudaError_t launch_cepstra_tail(const FParams& fp, int cep_smem_bytes, int max_frames, cudaStream_t stream) {
  const int half = (fp.delta_orders > 0 && !fp.logmel_only) ? fp.delta_width / 2 : 0;
  const int tile = kCepTThreads - 2 * half;
  const int span = max(max_frames, fp.out_frames);
  const dim3 grid(fp.n_clips, (span + tile - 1) / tile);
  if (fp.cep_small) {
    switch ((fp.n_mfcc + 3) / 4) {
      case 1: cepstra_small_kernel<1><<<grid, kCepTThreads, 0, stream>>>(fp); return cudaGetLastError();
      case 2: cepstra_small_kernel<2><<<grid, kCepTThreads, 0, stream>>>(fp); return cudaGetLastError();
      case 3: cepstra_small_kernel<3><<<grid, kCepTThreads, 0, stream>>>(fp); return cudaGetLastError();
      case 4: cepstra_small_kernel<4><<<grid, kCepTThreads, 0, stream>>>(fp); return cudaGetLastError();
      case 5: cepstra_small_kernel<5><<<grid, kCepTThreads, 0, stream>>>(fp); return cudaGetLastError();
      case 6: cepstra_small_kernel<6><<<grid, kCepTThreads, 0, stream>>>(fp); return cudaGetLastError();
      case 7: cepstra_small_kernel<7><<<grid, kCepTThreads, 0, stream>>>(fp); return cudaGetLastError();
      case 8: cepstra_small_kernel<8><<<grid, kCepTThreads, 0, stream>>>(fp); return cudaGetLastError();
      default: break;                                   // wider tables take the general kernel
    }
  }
  switch ((fp.n_mfcc + 3) / 4) {
    case 1: return launch_cep_t<1>(fp, grid, cep_smem_bytes, stream);
    case 2: return launch_cep_t<2>(fp, grid, cep_smem_bytes, stream);
    case 3: return launch_cep_t<3>(fp, grid, cep_smem_bytes, stream);
    case 4: return launch_cep_t<4>(fp, grid, cep_smem_bytes, stream);
    case 5: return launch_cep_t<5>(fp, grid, cep_smem_bytes, stream);
    case 6: return launch_cep_t<6>(fp, grid, cep_smem_bytes, stream);
    case 7: return launch_cep_t<7>(fp, grid, cep_smem_bytes, stream);
    case 8: return launch_cep_t<8>(fp, grid, cep_smem_bytes, stream);
    case 9: return launch_cep_t<9>(fp, grid, cep_smem_bytes, stream);
    case 10: return launch_cep_t<10>(fp, grid, cep_smem_bytes, stream);
    default: return cudaErrorInvalidValue;
  }
}


cudaError_t launch_tiles_path(const FParams& fp, int sm_count, int tile_smem_bytes, int cep_smem_bytes, int max_frames,
                              cudaStream_t stream) {
  cudaError_t e = launch_frame_prefix(fp, stream);
  if (e != cudaSuccess) return e;
  const bool noise = fp.noise_mode == ASR_NOISE_WHITE;
  switch (fp.dtype) {
    case ASR_I16: e = noise ? launch_tile_dt<ASR_I16, true>(fp, sm_count, tile_smem_bytes, stream)
                            : launch_tile_dt<ASR_I16, false>(fp, sm_count, tile_smem_bytes, stream); break;
    case ASR_F32: e = noise ? launch_tile_dt<ASR_F32, true>(fp, sm_count, tile_smem_bytes, stream)
                            : launch_tile_dt<ASR_F32, false>(fp, sm_count, tile_smem_bytes, stream); break;
    default: e = noise ? launch_tile_dt<ASR_F64, true>(fp, sm_count, tile_smem_bytes, stream)
                       : launch_tile_dt<ASR_F64, false>(fp, sm_count, tile_smem_bytes, stream); break;
  }
  if (e != cudaSuccess) return e;
  return launch_cepstra_tail(fp, cep_smem_bytes, max_frames, stream);
}

}  // namespace asr
