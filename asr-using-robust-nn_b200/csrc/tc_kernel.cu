// Tensor-core (tcgen05 + TMEM) path of the n_fft = 512 front end ("tc" path), and the tcgen05 self-test.
//
// tc512_kernel: persistent, one 640-thread CTA per SM; the frames of ALL clips form one flat list, a CTA owns a
// contiguous range of it and walks it in TILES of 128 frames = the 128 rows of one UMMA accumulator (one frame per TMEM lane).
//
//   The 512-point real FFT of a frame is the 256-point complex FFT of z[n] = x[2n] + i x[2n+1], n = b + 16 a,
//   k = c + 16 d:   Z[c + 16 d] = sum_b W16^(b d) * Y_b[c],   Y_b[c] = W256^(b c) * sum_a W16^(a c) * (w z)[b + 16 a].
//   PASS 1 (window, the stride-16 DFT over a, the inter-pass twiddle) is a real 32 x 32 linear map per residue b with
//   CONSTANT coefficients - it runs on the tensor cores: D_b[128 frames][32] = A_b[128][32] * M_b^T, float16 operands,
//   float32 accumulation in TMEM.  Float32 accuracy comes from a two-term split of both operands:
//     sample v (int16 units, after the exact float64 noise mix) = 32 * (hi + lo), hi = fp16(v / 32),
//     lo = fp16(v/32 - hi)  - exact for int16 audio, 2^-22 of the sample (3e-8 absolute below 0.25) otherwise;
//     M * 16 = MH1 + MH2 (fp16 + fp16 residual);   D = hi*MH1 + hi*MH2 + lo*MH1  = 2^14 * Y
//   (3 MMAs of K = 32 per b; the dropped terms are below 2^-22 of the frame's scale).
//   PASS 2 (16-point DFT over b per column c), the real-input unpack, |X|^2 run on the CUDA cores with lanes <-> frames:
//   a thread reads ITS frame's row from TMEM (tcgen05.ld), so every twiddle is a compile-time or constant-bank operand -
//   no table loads, no shuffles, no shared-memory exchange.  The power spectrum goes back into the TMEM columns the
//   thread has just consumed (bin k -> column 2k), the mel stage reads it from there (lanes <-> frames, the 4 warps of a
//   lane quarter share the bins), so the spectrum never touches shared memory.
//
//   Warp roles: 4 HELPER warps convert the samples of tile t+1 (dtype decode, float64 two-rounding noise mix,
//   reflect / zero padding, hi/lo split) into a residue-major staging array HL[b][pair row] while the 16 MAIN warps work on
//   tile t: phase A = per residue b: 128 rows x 64 B copied from HL[b] into the UMMA operand tile (K-major, unswizzled,
//   chunk-major so that the copy is conflict-free), one elected thread issues the 6 MMAs against the matrices of b (all 16 sets stay resident in
//   shared memory, 64 KB), a 2-slot operand ring keeps the tensor core busy while the next operand tile is built;
//   phase B = pass 2 / unpack / power / mel / log -> transposed log-mel workspace (then cepstra_*_kernel, tile_kernel.cu).
//
// Arithmetic restated from librosa.feature.mfcc (oracle/librosa_ref.py); call sites replaced:
// VDR/extract_features_construct_dataset.py:30, VDR/attacks.py:114,267 (BASELINE configs with n_fft = 512, int16 audio).
#include <cuda_fp16.h>
#include <cmath>
#include <cstring>
#include "common.cuh"
#include "fft_core.cuh"
#include "sample_access.cuh"
#include "tc_common.cuh"

namespace asr {

// ---- layout constants ----
constexpr int kTcRows = 128;                 // frames per tile
constexpr int kTcSub = 16;                   // frames per sub-block (descriptor / staging unit)
constexpr int kTcNSub = kTcRows / kTcSub;    // 8
constexpr int kTcMaxRuns = 4;                // runs of consecutive frames of one clip per sub-block
constexpr int kTcMainWarps = 16, kTcHelpWarps = 4;
constexpr int kTcMain = 32 * kTcMainWarps, kTcHelp = 32 * kTcHelpWarps, kTcThreads = kTcMain + kTcHelp;
constexpr int kTcRing = 32;                  // sub-block descriptors alive: tiles t-1 (log-mel stores) .. t+1 (being assembled)
constexpr int kTcCache = 32;
constexpr int kTcNA = 2;                     // operand-tile ring
constexpr int kTcATile = 16 * kTcRows * 4;   // bytes of one A tile (hi or lo): 4 chunks x 128 rows x 16 B
constexpr int kTcBMat = 16 * 32 * 4;         // bytes of one 32 x 32 float16 matrix
constexpr int kTcBSet = 2 * kTcBMat;         // MH1, MH2 of one residue b; all 16 sets stay resident in shared memory (64 KB)

struct __align__(16) TcRun {      // 48 bytes
  long long base;   // element offset of the clip
  double sig;       // sigma of the clip (white noise)
  int L;            // clip length
  int o0;           // ORIGINAL sample index of the first staged sample (t0 * hop - pad, may be negative)
  int count;        // staged samples: (n - 1) * hop + 512
  int aud0;         // staged position (samples, tile-local, multiple of 32) of the run's first sample
  int slot0;        // first frame slot of the run
  int pad0, pad1, pad2;
};
struct __align__(16) TcBlock {
  int n_runs, n_slots, g0, done;   // g0: flattened index of slot 0; done: no frame of this CTA's range is left
  TcRun run[kTcMaxRuns];
  int slot_row[kTcSub];            // pair row (staged position / 32) of the slot's first sample
};
struct TcMeta { long long off; double sig; int L; int T; };

// unpack twiddles of the 512-point real FFT on 2X: (-sin(2 pi k / 512), -cos(2 pi k / 512)), k = 0..255
__constant__ float2 c_tc_twu[256];

__device__ __forceinline__ void tc_tma_g2s(void* smem_dst, const void* gsrc, const unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_st1(const uint32_t addr, const float v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(addr), "r"(__float_as_uint(v)) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(const uint32_t addr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(addr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tc_log2(const float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Column c of pass 1 for this thread's frame: Y_b[c], b = 0..15, in bit-reversed order (the DIT input order).
__device__ __forceinline__ void tc_load_col(const uint32_t tm_lane, const int c, float (&ur)[16], float (&ui)[16]) {
#pragma unroll
  for (int b = 0; b < 16; ++b) {
    float v[2];
    tc::tmem_ld2(tm_lane + 32 * b + 2 * c, v);
    ur[brev<16>(b)] = v[0];
    ui[brev<16>(b)] = v[1];
  }
  tc::tmem_ld_wait();
}

// One unpack step on 2X (see tile_fft512): A = Z[k], B = Z[256 - k]  ->  4|X[k]|^2, 4|X[256 - k]|^2
__device__ __forceinline__ void tc_unpack(const float ar, const float ai, const float br, const float bi, const float2 w,
                                          float& pk, float& pm) {
  const float sr = ar + br, si = ai - bi;          // A + conj(B)
  const float dr = ar - br, di = ai + bi;          // A - conj(B)
  const float xr = fmaf(-w.y, di, fmaf(w.x, dr, sr));
  const float xi = fmaf(w.y, dr, fmaf(w.x, di, si));
  const float yr = fmaf(2.0f, sr, -xr), yi = fmaf(2.0f, si, -xi);
  pk = fmaf(xr, xr, xi * xi);
  pm = fmaf(yr, yr, yi * yi);
}

// columns c and 16 - c (1 <= c <= 7) of this thread's frame: bins c + 16 d and their mirrors -> TMEM column 2 * bin
__device__ __forceinline__ void tc_pair(const uint32_t tm_lane, const int c) {
  float ur[16], ui[16], vr[16], vi[16];
  tc_load_col(tm_lane, c, ur, ui);
  dft_dit<16>(ur, ui);
  const int cp = 16 - c;
  tc_load_col(tm_lane, cp, vr, vi);
  dft_dit<16>(vr, vi);
#pragma unroll
  for (int d = 0; d < 16; ++d) {
    float pk, pm;
    tc_unpack(ur[d], ui[d], vr[15 - d], vi[15 - d], c_tc_twu[c + 16 * d], pk, pm);   // k = c + 16 d, mirror 256 - k
    tmem_st1(tm_lane + 32 * d + 2 * c, pk);                                          // bin k
    tmem_st1(tm_lane + 32 * (15 - d) + 2 * cp, pm);                                  // bin 256 - k = cp + 16 (15 - d)
  }
}

// column 8 pairs with itself: bins 8 + 16 d and 8 + 16 (15 - d)
__device__ __forceinline__ void tc_col8(const uint32_t tm_lane) {
  float ur[16], ui[16];
  tc_load_col(tm_lane, 8, ur, ui);
  dft_dit<16>(ur, ui);
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    float pk, pm;
    tc_unpack(ur[d], ui[d], ur[15 - d], ui[15 - d], c_tc_twu[8 + 16 * d], pk, pm);
    tmem_st1(tm_lane + 32 * d + 16, pk);
    tmem_st1(tm_lane + 32 * (15 - d) + 16, pm);
  }
}
// column 0 pairs with itself: bins 16 d and 16 (16 - d); bin 0 pairs with itself and yields bins 0 and 256 (-> column 1)
__device__ __forceinline__ void tc_col0(const uint32_t tm_lane) {
  float ur[16], ui[16];
  tc_load_col(tm_lane, 0, ur, ui);
  dft_dit<16>(ur, ui);
#pragma unroll
  for (int d = 0; d <= 8; ++d) {
    float pk, pm;
    tc_unpack(ur[d], ui[d], ur[(16 - d) & 15], ui[(16 - d) & 15], c_tc_twu[16 * d], pk, pm);
    tmem_st1(tm_lane + 32 * d, pk);
    if (d == 0) tmem_st1(tm_lane + 1, pm);                       // bin 256
    else if (d < 8) tmem_st1(tm_lane + 32 * (16 - d), pm);       // bin 16 (16 - d)
  }
}

// ------------------------------------------------------------------------------------------------
template <bool NOISE>
__global__ void __launch_bounds__(kTcThreads, 1) tc512_kernel(const __grid_constant__ FParams fp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ TcBlock ring[kTcRing];
  __shared__ TcMeta s_meta[kTcCache];
  __shared__ __align__(8) uint64_t bar_d_full, bar_a_free[kTcNA];
  __shared__ uint32_t s_tmem;
  __shared__ int s_ready, s_free;      // tiles converted by the helper warps / tiles whose operand copies are done (monotonic)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const bool helper = warp_u >= kTcMainWarps;

  // ---- this CTA's range of the flattened frame list (whole tiles) ----
  const int total = __ldg(fp.fstart + fp.n_clips);
  const int n_tiles = (total + kTcRows - 1) / kTcRows;
  const int per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const long long gb = static_cast<long long>(blockIdx.x) * per * kTcRows;
  if (gb >= total) return;
  const int g_begin = static_cast<int>(gb);
  const int g_end = static_cast<int>(min(static_cast<long long>(total), gb + static_cast<long long>(per) * kTcRows));

  // ---- shared-memory carve-up ----
  float* s_tab = reinterpret_cast<float*>(smem_raw);                       // mel tables
  uint2* s_hl = reinterpret_cast<uint2*>(smem_raw + fp.tc_sm_hl);          // [16][tc_hl_stride] (hi pair, lo pair)
  unsigned char* s_a = smem_raw + fp.tc_sm_a;                              // [kTcNA][hi tile | lo tile]
  unsigned char* s_b = smem_raw + fp.tc_sm_b;                              // [16 b][MH1 | MH2], resident
  float* s_slots = reinterpret_cast<float*>(smem_raw + fp.tc_sm_slots);    // [tc_n_slots][128] boundary subtotals
  {
    float4* dst = reinterpret_cast<float4*>(s_tab);
    for (int i = tid; i < fp.blob_f4; i += kTcThreads) dst[i] = __ldg(fp.blob + i);
    uint4* mb = reinterpret_cast<uint4*>(s_b);
    const uint4* gm = reinterpret_cast<const uint4*>(fp.tc_mats);
    for (int i = tid; i < 16 * kTcBSet / 16; i += kTcThreads) mb[i] = __ldg(gm + i);
  }
  tc::fence_async_smem();                               // the matrices are read by the tensor core (async proxy)
  const float4* s_wtab = reinterpret_cast<const float4*>(s_tab + fp.off_wtab);
  const int4* s_pieces = reinterpret_cast<const int4*>(s_tab + fp.off_steps);
  const int2* s_wrange = reinterpret_cast<const int2*>(s_tab + fp.off_wrange);
  const int4* s_bnd = reinterpret_cast<const int4*>(s_tab + fp.tc_off_bnd);
  if (tid == 0) {
    s_ready = 0; s_free = 0;
    tc::mbar_init(&bar_d_full, 1);
    for (int i = 0; i < kTcNA; ++i) tc::mbar_init(&bar_a_free[i], 1);
    tc::mbar_init_fence();
  }
  if (warp_u == 0) tc::tmem_alloc(&s_tmem, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, s_tmem, 0);      // warp-uniform copy (uniform-register operand of the tcgen05 instructions)

  // bounded wait with a breadcrumb: on a protocol error the waiting thread leaves (site, CTA, thread, parity) in mapped
  // host memory and traps, so the launch fails with a diagnosis instead of hanging
  auto dwait = [&](const int site, uint64_t* bar, const uint32_t parity) {
    for (uint32_t spin = 0; !tc::mbar_try_wait(bar, parity); ++spin)
      if (spin > (1u << 22)) {
        if (fp.tc_dbg) {
          volatile int* d = fp.tc_dbg;
          if (atomicCAS(const_cast<int*>(fp.tc_dbg), 0, site) == 0) {
            d[1] = blockIdx.x; d[2] = threadIdx.x; d[3] = static_cast<int>(parity);
            __threadfence_system();
          }
        }
        __trap();
      }
  };
  // coarse hand-shakes between the warp roles (once per tile): monotonic counters, release / acquire at CTA scope
  auto post = [&](int* ctr, const int v) { asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(tc::smem_u32(ctr)), "r"(v) : "memory"); };
  auto await = [&](const int site, int* ctr, const int v) {
    for (uint32_t spin = 0;; ++spin) {
      int cur;
      asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(cur) : "r"(tc::smem_u32(ctr)) : "memory");
      if (cur >= v) break;
      __nanosleep(64);
      if (spin > (1u << 24)) {
        if (fp.tc_dbg && atomicCAS(const_cast<int*>(fp.tc_dbg), 0, site) == 0) {
          volatile int* d = fp.tc_dbg;
          d[1] = blockIdx.x; d[2] = threadIdx.x; d[3] = v; d[4] = s_ready; d[5] = s_free;
          __threadfence_system();
        }
        __trap();
      }
    }
  };
  auto bar_main = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kTcMain) : "memory"); };
  auto bar_help = [&]() { asm volatile("bar.sync 2, %0;" ::"n"(kTcHelp) : "memory"); };

  if (helper) {
    // =============================== helper warps: descriptors + sample conversion ===============================
    const int ht = tid - kTcMain, hw = warp_u - kTcMainWarps;
    int b_cur = 0, t_cur = 0, g_cur = g_begin, cache_base = 0;       // block cursor (warp hw == 0; lane 0 holds the live copy)
    int aud = 0;                                                     // staged samples used in the tile being assembled
    auto load_meta = [&](const int base) {
#pragma unroll
      for (int e = lane; e < kTcCache; e += 32) {
        const int b = min(base + e, fp.n_clips - 1);
        TcMeta m;
        m.off = __ldg(fp.offsets + b);
        m.sig = NOISE ? __ldg(fp.sigma + b) : 0.0;
        m.L = __ldg(fp.lengths + b);
        m.T = __ldg(fp.nframes + b);
        s_meta[e] = m;
      }
    };
    // one sub-block descriptor: up to 16 frames in up to kTcMaxRuns runs; staged positions continue inside the tile
    auto assemble = [&](TcBlock& blk, const bool first_of_tile) {     // whole descriptor warp
      if (first_of_tile) aud = 0;
      int n_slots = 0, n_runs = 0;
      const int g0 = g_cur;
      for (;;) {
        int need = 0;
        if (lane == 0) {
          while (n_runs < kTcMaxRuns && n_slots < kTcSub && g_cur < g_end && b_cur < fp.n_clips) {
            if (b_cur >= cache_base + kTcCache) { need = 1; break; }
            const TcMeta cm = s_meta[b_cur - cache_base];
            if (t_cur < cm.T) {
              int n = min(min(kTcSub - n_slots, cm.T - t_cur), g_end - g_cur);
              const int a0 = (aud + 31) & ~31;
              const int room = fp.tc_hl_rows * 32 - a0 - 512;          // samples left for further hops of this run
              if (room < 0) break;                                     // staging array full: the rest goes to the next tile
              n = min(n, 1 + room / fp.hop);
              TcRun& run = blk.run[n_runs];
              run.base = cm.off;
              run.sig = cm.sig;
              run.L = cm.L;
              run.o0 = t_cur * fp.hop - fp.pad;
              run.count = (n - 1) * fp.hop + 512;
              run.aud0 = a0;
              run.slot0 = n_slots;
              aud = a0 + run.count;
              n_slots += n; t_cur += n; g_cur += n; ++n_runs;
            }
            if (t_cur >= cm.T) { ++b_cur; t_cur = 0; }
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
        blk.n_runs = n_runs; blk.n_slots = n_slots; blk.g0 = g0;
        blk.done = (n_slots == 0 && (g_cur >= g_end || b_cur >= fp.n_clips)) ? 1 : 0;
      }
      __syncwarp();
      const int nr = blk.n_runs, ns = blk.n_slots;
      if (lane < kTcSub) {
        int a = 0;
        if (lane < ns) {
#pragma unroll
          for (int r = 0; r < kTcMaxRuns; ++r)
            if (r < nr && lane >= blk.run[r].slot0) a = (blk.run[r].aud0 + (lane - blk.run[r].slot0) * fp.hop) >> 5;
        }
        blk.slot_row[lane] = a;
      }
      __syncwarp();
    };
    // samples of one sub-block -> HL[b][pair row]: decode, [exact noise mix], padding, hi/lo split; once per sample
    auto convert_sub = [&](const TcBlock& blk) {
      const int nr = blk.n_runs;
      for (int r = 0; r < nr; ++r) {
        const TcRun& run = blk.run[r];
        const long long base = run.base;
        const double sig = run.sig;
        const int L = run.L, o0 = run.o0, np = run.count >> 1, p0 = run.aud0 >> 1;
        for (int i = ht; i < np; i += kTcHelp) {
          // value_at / padded_at (sample_access.cuh): int16 / 32768, float64 two-rounding mix, reflect / zero padding
          const float x0 = padded_at<ASR_I16>(fp, base, L, o0 + fp.pad + 2 * i, sig) * 1024.0f;      // = v / 32, v in int16 units
          const float x1 = padded_at<ASR_I16>(fp, base, L, o0 + fp.pad + 2 * i + 1, sig) * 1024.0f;
          const __half2 hi = __floats2half2_rn(x0, x1);
          const float2 hf = __half22float2(hi);
          const __half2 lo = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
          const int P = p0 + i;
          uint2 w;
          w.x = *reinterpret_cast<const unsigned*>(&hi);
          w.y = *reinterpret_cast<const unsigned*>(&lo);
          s_hl[(P & 15) * fp.tc_hl_stride + (P >> 4)] = w;
        }
      }
    };

    if (hw == 0) {
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
      for (int j = 0; j < kTcNSub; ++j) assemble(ring[j], j == 0);
    }
    bar_help();
    for (int t = 0;; ++t) {
      const int k0 = t * kTcNSub;
      if (ring[k0 & (kTcRing - 1)].done) {               // nothing left: tell the main warps (they read the same flag)
        if (ht == 0) post(&s_ready, t + 1);
        break;
      }
      if (t > 0) await(1, &s_free, t);                   // tile t-1's operand copies are done
      if (!(fp.dbg_skip & 2))                             // (timing experiments: main warps alone)
        for (int j = 0; j < kTcNSub; ++j) convert_sub(ring[(k0 + j) & (kTcRing - 1)]);
      if (hw == 0)                                        // descriptors of tile t+1 (tile t-1's are still read by the main warps)
        for (int j = 0; j < kTcNSub; ++j) assemble(ring[(k0 + kTcNSub + j) & (kTcRing - 1)], j == 0);
      bar_help();
      if (ht == 0) post(&s_ready, t + 1);
    }
  } else {
    // =============================== main warps ===============================
    const int row = tid & (kTcRows - 1), cq = tid >> 7;               // operand copy: row, 16-byte chunk (4 values of a)
    const int q = warp_u & 3, grp = warp_u >> 2;                      // TMEM lane quarter, column / bin group
    const int frow = 32 * q + lane;                                   // this thread's frame row in phase B
    const uint32_t tm_lane = tmem + (static_cast<uint32_t>(32 * q) << 16);
    const uint32_t idesc = tc::idesc_f16_f32(kTcRows, 32);
    const uint64_t desc_a = tc::smem_desc(tc::smem_u32(s_a), 16 * kTcRows, 128);     // hi tile of slot 0, K step 0
    const uint64_t desc_b = tc::smem_desc(tc::smem_u32(s_b), 512, 128);              // MH1 of b = 0, K step 0
    const int2 my_pieces = s_wrange[grp];
    uint32_t gbc = 0;                                                 // running count of (tile, b) steps
    long long tk[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};                // (profiling breadcrumbs, CTA 0 thread 0: cycles per phase)
    const bool prof = fp.tc_dbg != nullptr && (fp.dbg_skip & 128) && blockIdx.x == 0 && tid == 0;
    long long tlast = prof ? clock64() : 0;
    auto lap = [&](const int i) { if (prof) { const long long now = clock64(); tk[i] += now - tlast; tlast = now; } };
    for (int t = 0;; ++t) {
      const int k0 = t * kTcNSub;
      await(2, &s_ready, t + 1);
      lap(0);
      if (ring[k0 & (kTcRing - 1)].done) break;
      // ---------------- phase A: operand tiles + MMAs ----------------
      const int rowp = ring[(k0 + (row >> 4)) & (kTcRing - 1)].slot_row[row & 15];
      const uint2* src_row = s_hl + rowp + 4 * cq;
      unsigned char* const a_dst = s_a + cq * (16 * kTcRows) + row * 16;
#pragma unroll 1
      for (int b = 0; b < 16; ++b, ++gbc) {
        const uint32_t sa = gbc & (kTcNA - 1);
        if (gbc >= kTcNA) dwait(3, &bar_a_free[sa], ((gbc / kTcNA) - 1) & 1);   // MMAs of step gbc - kTcNA have read this slot
        lap(1);
        if (!(fp.dbg_skip & 32)) {
          const uint2* src = src_row + b * fp.tc_hl_stride;
          const uint2 w0 = src[0], w1 = src[1], w2 = src[2], w3 = src[3];
          unsigned char* at = a_dst + sa * (2 * kTcATile);
          *reinterpret_cast<uint4*>(at) = make_uint4(w0.x, w1.x, w2.x, w3.x);
          *reinterpret_cast<uint4*>(at + kTcATile) = make_uint4(w0.y, w1.y, w2.y, w3.y);
        }
        lap(2);
        tc::fence_async_smem();
        lap(3);
        bar_main();
        lap(4);
        if (warp_u == 0 && tc::elect_one()) {
          tc::tc_fence_after();
          // descriptors differ in their start address only (bits 0..13, in 16-byte units)
          const uint64_t dah = desc_a + (sa * (2 * kTcATile) >> 4), dal = dah + (kTcATile >> 4);
          const uint64_t d1 = desc_b + (b * kTcBSet >> 4), d2 = d1 + (kTcBMat >> 4);
          const uint32_t dcol = tmem + 32 * b;
          constexpr uint32_t kA = 2 * (16 * kTcRows) >> 4, kB = 2 * 512 >> 4;    // second K step: two chunks further
          if (fp.dbg_skip & 256) {                        // (timing experiment: six independent accumulators)
          tc::mma_f16(tmem + 0, dah, d1, idesc, 1);
          tc::mma_f16(tmem + 64, dah + kA, d1 + kB, idesc, 1);
          tc::mma_f16(tmem + 128, dah, d2, idesc, 1);
          tc::mma_f16(tmem + 192, dah + kA, d2 + kB, idesc, 1);
          tc::mma_f16(tmem + 256, dal, d1, idesc, 1);
          tc::mma_f16(tmem + 320, dal + kA, d1 + kB, idesc, 1);
          } else if (!(fp.dbg_skip & 16)) {
          tc::mma_f16(dcol, dah, d1, idesc, 0);
          tc::mma_f16(dcol, dah + kA, d1 + kB, idesc, 1);
          if (!(fp.dbg_skip & 64)) {
          tc::mma_f16(dcol, dah, d2, idesc, 1);
          tc::mma_f16(dcol, dah + kA, d2 + kB, idesc, 1);
          tc::mma_f16(dcol, dal, d1, idesc, 1);
          tc::mma_f16(dcol, dal + kA, d1 + kB, idesc, 1);
          }
          }
          tc::mma_commit(&bar_a_free[sa]);
        }
        lap(5);
      }
      if (tid == 0) {
        tc::mma_commit(&bar_d_full);
        post(&s_free, t + 1);                                         // every operand copy of this tile is behind the last bar_main
      }
      // ---------------- phase B: pass 2, unpack, power -> TMEM; mel; log -> workspace ----------------
      const TcBlock& myb = ring[(k0 + (frow >> 4)) & (kTcRing - 1)];
      const int gidx = (frow & 15) < myb.n_slots ? myb.g0 + (frow & 15) : -1;
      dwait(6, &bar_d_full, static_cast<uint32_t>(t & 1));
      tc::tc_fence_after();
      lap(6);
      if (!(fp.dbg_skip & 4)) {                         // (timing experiments)
      if (grp == 0) { tc_pair(tm_lane, 1); tc_pair(tm_lane, 2); }
      else if (grp == 1) { tc_pair(tm_lane, 3); tc_pair(tm_lane, 4); }
      else if (grp == 2) { tc_pair(tm_lane, 5); tc_pair(tm_lane, 6); }
      else { tc_pair(tm_lane, 7); tc_col0(tm_lane); tc_col8(tm_lane); }
      }
      tmem_st_wait();
      tc::tc_fence_before();
      bar_main();
      tc::tc_fence_after();
      lap(7);
      {
        // mel: this group's pieces in ascending-bin order, one piece per segment; filter seg-1 = rise(seg-1) + fall(seg)
        auto emit = [&](const int code, const float m) {
          const int f = (code & 0xFFFF) - 1, mode = code >> 16;
          if (f < 0) return;
          if (mode == 0) {
            if (gidx >= 0) fp.lm[static_cast<long long>(f) * fp.lm_stride + gidx] = 3.01029995663981195f * tc_log2(fmaxf(fp.amin, m));
          } else {
            s_slots[(mode - 1) * kTcRows + frow] = m;
          }
        };
        float pending = 0.0f;
        const int4* pp = s_pieces + my_pieces.x;
#pragma unroll 1
        for (int n = (fp.dbg_skip & 8) ? 0 : my_pieces.y; n > 0; --n, ++pp) {
          const int4 pc = *pp;                          // (first float4 of bins, steps, first weight float4, emit code)
          const float4* wp = s_wtab + pc.z;
          float a0 = 0.0f, a1 = 0.0f, b0 = 0.0f, b1 = 0.0f;
#pragma unroll 1
          for (int k = 0; k < pc.y; ++k) {
            const int qd = pc.x + k;                    // bins 4 qd .. 4 qd + 3 live in columns 8 qd, +2, +4, +6
            float v[8];
            if (qd < 64) {
              tmem_ld8(tm_lane + 8 * qd, v);
              tc::tmem_ld_wait();
            } else {                                    // bin 256 (column 1); 257.. do not exist (zero weights)
              float u[2];
              tc::tmem_ld2(tm_lane, u);
              tc::tmem_ld_wait();
              v[0] = u[1]; v[2] = 0.0f; v[4] = 0.0f; v[6] = 0.0f;
            }
            const float4 w01 = wp[2 * k], w23 = wp[2 * k + 1];   // (fall, rise) of bins 0,1 and 2,3
            a0 = fmaf(w01.x, v[0], a0); b0 = fmaf(w01.y, v[0], b0);
            a1 = fmaf(w01.z, v[2], a1); b1 = fmaf(w01.w, v[2], b1);
            a0 = fmaf(w23.x, v[4], a0); b0 = fmaf(w23.y, v[4], b0);
            a1 = fmaf(w23.z, v[6], a1); b1 = fmaf(w23.w, v[6], b1);
          }
          emit(pc.w, pending + (a0 + a1));
          pending = b0 + b1;
        }
      }
      tc::tc_fence_before();
      bar_main();                                       // TMEM is free again; boundary subtotals are visible
      tc::tc_fence_after();
      lap(8);
      for (int idx = tid; idx < fp.tc_n_bnd * kTcRows; idx += kTcMain) {
        const int4 e = s_bnd[idx >> 7];                 // (filter, first slot, slots, -)
        const int r = idx & (kTcRows - 1);
        float m = 0.0f;
        for (int i = 0; i < e.z; ++i) m += s_slots[(e.y + i) * kTcRows + r];
        const TcBlock& rb = ring[(k0 + (r >> 4)) & (kTcRing - 1)];
        if ((r & 15) < rb.n_slots)
          fp.lm[static_cast<long long>(e.x) * fp.lm_stride + rb.g0 + (r & 15)] = 3.01029995663981195f * tc_log2(fmaxf(fp.amin, m));
      }
      bar_main();                                       // slots may be rewritten by the next tile's mel
      lap(9);
    }
    if (prof) {
      volatile int* d = fp.tc_dbg;
      for (int i = 0; i < 10; ++i) d[6 + i] = static_cast<int>(tk[i] >> 4);
      __threadfence_system();
    }
    // every asynchronous arrival this CTA has asked for must have landed before its shared memory is released: the
    // commits of the last two steps are the only ones nobody has waited for
    if (tid == 0 && !(fp.dbg_skip & 1)) {
      for (uint32_t i = 1; i <= kTcNA && i <= gbc; ++i) dwait(7, &bar_a_free[(gbc - i) & (kTcNA - 1)], ((gbc - i) / kTcNA) & 1);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp_u == 0) tc::tmem_dealloc(tmem, 512);
}

cudaError_t tc_upload_constants() {
  float2 tw[256];
  const double kPi = 3.141592653589793238462643383279502884;
  for (int k = 0; k < 256; ++k) {
    const double ang = 2.0 * kPi * k / 512.0;
    tw[k] = make_float2(static_cast<float>(-std::sin(ang)), static_cast<float>(-std::cos(ang)));
  }
  return cudaMemcpyToSymbol(c_tc_twu, tw, sizeof(tw));
}

int tc_static_smem_bytes() { return static_cast<int>(sizeof(TcBlock) * kTcRing + sizeof(TcMeta) * kTcCache + 256); }

template <bool NOISE>
static cudaError_t launch_tc_n(const FParams& fp, int sm_count, int smem_bytes, cudaStream_t stream) {
  static int granted[kMaxDevices] = {0};
  const cudaError_t eg = ensure_dyn_smem(tc512_kernel<NOISE>, smem_bytes, 0, granted, true);
  if (eg != cudaSuccess) return eg;
  tc512_kernel<NOISE><<<sm_count, kTcThreads, smem_bytes, stream>>>(fp);
  return cudaGetLastError();
}

cudaError_t launch_tc_path(const FParams& fp, int sm_count, int tc_smem_bytes, int cep_smem_bytes, int max_frames,
                           cudaStream_t stream) {
  cudaError_t e = launch_frame_prefix(fp, stream);
  if (e != cudaSuccess) return e;
  e = fp.noise_mode == ASR_NOISE_NONE ? launch_tc_n<false>(fp, sm_count, tc_smem_bytes, stream)
                                      : launch_tc_n<true>(fp, sm_count, tc_smem_bytes, stream);
  if (e != cudaSuccess) return e;
  return launch_cepstra_tail(fp, cep_smem_bytes, max_frames, stream);
}

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
