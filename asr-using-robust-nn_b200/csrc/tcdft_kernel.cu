// Tensor-core dense DFT front end for FFT sizes without a register FFT (the reference's speaker-corpus parameters:
// n_fft = win_length = 441, hop 220, 128 mel - SR/extract_features_construct_dataset.py:227-228).
//
// tcdft_kernel: persistent, one 512-thread CTA per SM; the frames of ALL clips form one flat list, a CTA owns a contiguous
// range of it and walks it in tiles of 128 frames = the 128 rows of one UMMA accumulator (one frame per TMEM lane).
//   The windowed real DFT of a frame is ONE dense contraction with constant coefficients,
//       [Re X | Im X][128 frames][2 * NH] = A[128][K] * B[2 * NH][K]^T,  B = 16 w[n] (cos, -sin)(2 pi k n / n_fft),
//   K = n_fft rounded up to 16 (28 steps of 16 for 441), NH = bins rounded up to 16 (224): 2 x 224 float32 accumulator
//   columns in TMEM.  Float32-level accuracy from float16 operands by the two-term split of both sides,
//       x = hi + lo (fp16 + fp16 residual),  B = B1 + B2,   D = hi*B1 + hi*B2 + lo*B1
//   (6 MMAs of 128 x 224 x 16 per K step; the dropped lo*B2 term is 2^-22 of the frame's scale).
//   Per K step: the 512 threads gather their frames' 16 samples straight from global memory (the staging of the next step
//   is in flight while this one is converted), split them and write the operand tile (K-major, unswizzled, chunk-major);
//   one bulk copy (cp.async.bulk) brings the step's slab of B from L2; one elected thread issues the MMAs.  4 stages.
//   Then, with lanes <-> frames: |X|^2 back into the Re columns, the mel stage reads it from TMEM (the 4 warps of a lane
//   quarter share the bins), 10 log10 -> transposed log-mel workspace -> cepstra_t_kernel (tile_kernel.cu).
// Why tensor cores here and not for n_fft = 512: a dense DFT costs n_fft * bins MACs per frame against ~n log n for an FFT,
// but 441 = 3^2 7^2 has no radix-2 FFT, the FP32 direct DFT of this library runs at 0.16 M windows/s, and the contraction is
// a large-N GEMM with a constant matrix - the shape tcgen05 is built for (profiles/: ncu A/B).
#include <cuda_fp16.h>
#include <cmath>
#include <cstring>
#include "common.cuh"
#include "tc_common.cuh"

namespace asr {

constexpr int kDfRows = 128;
constexpr int kDfProducers = 512;              // 16 warps: operand staging, then pass-2 work with lanes <-> frames
constexpr int kDfThreads = kDfProducers + 32;  // + the issuer warp (B slabs by TMA, MMAs)
constexpr int kDfStages = 4;
constexpr int kDfATile = 2 * kDfRows * 16;     // one operand tile (hi or lo) of a K step: 2 chunks x 128 rows x 16 B

__device__ __forceinline__ void df_tma_g2s(void* smem_dst, const void* gsrc, const unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void df_ld8(const uint32_t addr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(addr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void df_ld4(const uint32_t addr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void df_st8(const uint32_t addr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ float df_log2(const float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// raw sample as loaded (the conversion to float32 happens when the value is consumed, K steps later: a conversion right
// behind the load would wait for the load and serialise the staging)
template <int DT> struct DfRaw { using type = double; };
template <> struct DfRaw<ASR_I16> { using type = short; };
template <> struct DfRaw<ASR_F32> { using type = float; };
template <int DT>
__device__ __forceinline__ typename DfRaw<DT>::type df_load(const void* __restrict__ audio, const long long i) {
  return __ldg(reinterpret_cast<const typename DfRaw<DT>::type*>(audio) + i);
}
template <int DT>
__device__ __forceinline__ float df_value(const typename DfRaw<DT>::type r) {
  if (DT == ASR_I16) return static_cast<float>(r) * (1.0f / 32768.0f);
  return static_cast<float>(r);
}

template <int DT>
__global__ void __launch_bounds__(kDfThreads, 1) tcdft_kernel(const __grid_constant__ FParams fp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ long long s_base[kDfRows];            // element offset of the row's clip
  __shared__ int s_o0[kDfRows], s_len[kDfRows], s_g[kDfRows];   // original index of the frame's first sample, clip length, flat frame index (-1: no frame)
  __shared__ __align__(8) uint64_t bar_free[kDfStages], bar_full[kDfStages], bar_d;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);

  const int total = __ldg(fp.fstart + fp.n_clips);
  const int n_tiles = (total + kDfRows - 1) / kDfRows;
  const int per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * per, t_end = min(n_tiles, t_begin + per);
  if (t_begin >= t_end) return;

  float* s_tab = reinterpret_cast<float*>(smem_raw);
  unsigned char* s_stage = smem_raw + fp.df_sm_stage;                      // [kDfStages][A hi | A lo | B slab]
  float* s_slots = reinterpret_cast<float*>(smem_raw + fp.tc_sm_slots);
  {
    float4* dst = reinterpret_cast<float4*>(s_tab);
    for (int i = tid; i < fp.blob_f4; i += kDfThreads) dst[i] = __ldg(fp.blob + i);
  }
  const float4* s_wtab = reinterpret_cast<const float4*>(s_tab + fp.off_wtab);
  const int4* s_pieces = reinterpret_cast<const int4*>(s_tab + fp.off_steps);
  const int2* s_wrange = reinterpret_cast<const int2*>(s_tab + fp.off_wrange);
  const int4* s_bnd = reinterpret_cast<const int4*>(s_tab + fp.tc_off_bnd);
  if (tid == 0) {
    for (int i = 0; i < kDfStages; ++i) { tc::mbar_init(&bar_free[i], 1); tc::mbar_init(&bar_full[i], 1); }
    tc::mbar_init(&bar_d, 1);
    tc::mbar_init_fence();
  }
  if (warp_u == 0) tc::tmem_alloc(&s_tmem, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, s_tmem, 0);

  auto dwait = [&](const int site, uint64_t* bar, const uint32_t parity) {
    for (uint32_t spin = 0; !tc::mbar_try_wait(bar, parity); ++spin)
      if (spin > (1u << 22)) {
        if (fp.tc_dbg && atomicCAS(const_cast<int*>(fp.tc_dbg), 0, 100 + site) == 0) {
          volatile int* d = fp.tc_dbg;
          d[1] = blockIdx.x; d[2] = threadIdx.x; d[3] = static_cast<int>(parity);
          __threadfence_system();
        }
        __trap();
      }
  };

  const int n_fft = fp.n_fft, NH = fp.df_nh, KS = fp.df_ksteps;
  const int stage_bytes = 2 * kDfATile + fp.df_bslab;
  const int row = tid >> 2, q4 = tid & 3;                                   // operand gather: 4 consecutive lanes read the 16 consecutive samples of a row
  const int q = warp_u & 3, grp = warp_u >> 2;                              // TMEM lane quarter, bin group
  const int frow = 32 * q + lane;
  const uint32_t tm_lane = tmem + (static_cast<uint32_t>(32 * q) << 16);
  const uint32_t idesc = tc::idesc_f16_f32(kDfRows, NH);
  const uint64_t desc_a = tc::smem_desc(tc::smem_u32(s_stage), 16 * kDfRows, 128);
  const uint64_t desc_b = tc::smem_desc(tc::smem_u32(s_stage + 2 * kDfATile), 32 * NH, 128);   // chunk stride: 2 NH rows x 16 B
  const unsigned char* g_mats = reinterpret_cast<const unsigned char*>(fp.tc_mats);
  const int2 my_pieces = s_wrange[grp];
  const bool reflect = fp.pad_mode == ASR_PAD_REFLECT;

  // the 4 samples n0 .. n0+3 of this thread's row for a K step (reflect / zero padding at the clip edges, zeros past n_fft)
  long long r_base = 0;
  int r_o0 = 0, r_len = 0;
  using raw_t = typename DfRaw<DT>::type;
  auto gather = [&](const int ks, raw_t (&v)[4]) {
    const int n0 = 16 * ks + 4 * q4;
    const long long base = r_base;
    const int o0 = r_o0, L = r_len;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = n0 + e;
      int o = o0 + n;
      bool zero = n >= n_fft || L <= 0;
      if (o < 0) { zero = zero || !reflect; o = -o; }
      else if (o >= L) { zero = zero || !reflect; o = 2 * (L - 1) - o; }
      v[e] = zero ? static_cast<raw_t>(0) : df_load<DT>(fp.audio, base + o);
    }
  };

  // named barriers: 1..4 = "operand tile of stage s written" (512 producer threads arrive, the issuer warp syncs),
  // 5 = producers only, 6 = "TMEM read out" (producers arrive, issuer syncs)
  auto bar_stage_arrive = [&](const uint32_t s) { asm volatile("bar.arrive %0, %1;" ::"r"(1 + s), "n"(kDfThreads) : "memory"); };
  auto bar_stage_sync = [&](const uint32_t s) { asm volatile("bar.sync %0, %1;" ::"r"(1 + s), "n"(kDfThreads) : "memory"); };
  auto bar_prod = [&]() { asm volatile("bar.sync 5, %0;" ::"n"(kDfProducers) : "memory"); };

  uint32_t ksg = 0;                                                         // running count of K steps (stage ring)
  if (warp_u == kDfProducers / 32) {
    // =============================== issuer warp: B slabs (TMA) and MMAs ===============================
    for (int t = t_begin; t < t_end; ++t) {
      for (int k = 0; k < 2 && k < KS; ++k) {                               // slabs of the first two steps
        const uint32_t s = (ksg + k) & (kDfStages - 1);
        if (ksg + k >= kDfStages) dwait(1, &bar_free[s], (((ksg + k) / kDfStages) - 1) & 1);
        if (tc::elect_one()) {
          tc::mbar_arrive_expect_tx(&bar_full[s], fp.df_bslab);
          df_tma_g2s(s_stage + s * stage_bytes + 2 * kDfATile, g_mats + static_cast<size_t>(k) * fp.df_bslab, fp.df_bslab, &bar_full[s]);
        }
        __syncwarp();
      }
      if (t > t_begin) {                                                    // the previous tile's spectrum has been read out of TMEM
        asm volatile("bar.sync 6, %0;" ::"n"(kDfThreads) : "memory");
        tc::tc_fence_after();
      }
#pragma unroll 1
      for (int ks = 0; ks < KS; ++ks, ++ksg) {
        const uint32_t s = ksg & (kDfStages - 1);
        bar_stage_sync(s);                                                  // the producers have written this step's operand tile
        dwait(3, &bar_full[s], (ksg / kDfStages) & 1);
        tc::tc_fence_after();
        if (tc::elect_one()) {
          const uint64_t dah = desc_a + (s * stage_bytes >> 4), dal = dah + (kDfATile >> 4);
          const uint64_t b1 = desc_b + (s * stage_bytes >> 4), b2 = b1 + (fp.df_bslab >> 5);   // second matrix: half a slab further
          const uint32_t acc = ks > 0 ? 1u : 0u;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t dcol = tmem + h * NH;
            const uint64_t bh = static_cast<uint64_t>(h * NH);               // rows h*NH .. : 16 bytes each
            tc::mma_f16(dcol, dah, b1 + bh, idesc, acc);
            tc::mma_f16(dcol, dah, b2 + bh, idesc, 1);
            tc::mma_f16(dcol, dal, b1 + bh, idesc, 1);
          }
          tc::mma_commit(&bar_free[s]);
          if (ks == KS - 1) tc::mma_commit(&bar_d);
        }
        __syncwarp();
        if (ks + 2 < KS) {                                                  // slab of step ks + 2 into the stage step ks - 2 used
          const uint32_t s2 = (ksg + 2) & (kDfStages - 1);
          if (ksg + 2 >= kDfStages) dwait(4, &bar_free[s2], (((ksg + 2) / kDfStages) - 1) & 1);
          if (tc::elect_one()) {
            tc::mbar_arrive_expect_tx(&bar_full[s2], fp.df_bslab);
            df_tma_g2s(s_stage + s2 * stage_bytes + 2 * kDfATile, g_mats + static_cast<size_t>(ks + 2) * fp.df_bslab, fp.df_bslab, &bar_full[s2]);
          }
          __syncwarp();
        }
      }
    }
    // every asynchronous arrival this CTA asked for must have landed before its shared memory is released
    for (uint32_t i = 1; i <= kDfStages && i <= ksg; ++i) dwait(6, &bar_free[(ksg - i) & (kDfStages - 1)], ((ksg - i) / kDfStages) & 1);
  } else {
    // =============================== producer / epilogue warps ===============================
    long long tk[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};                      // (profiling breadcrumbs, CTA 0 thread 0: cycles per phase)
    const bool prof = fp.tc_dbg != nullptr && (fp.dbg_skip & 128) && blockIdx.x == 0 && tid == 0;
    long long tlast = prof ? clock64() : 0;
    auto lap = [&](const int i) { if (prof) { const long long now = clock64(); tk[i] += now - tlast; tlast = now; } };
    for (int t = t_begin; t < t_end; ++t) {
      // ---- rows of this tile: flat frame index -> (clip, frame) ----
      if (tid < kDfRows) {
        const int g = t * kDfRows + tid;
        long long base = 0;
        int o0 = 0, L = 0, gi = -1;
        if (g < total) {
          int lo = 0, hi = fp.n_clips;                                      // largest b with fstart[b] <= g
          while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(fp.fstart + mid) <= g) lo = mid; else hi = mid;
          }
          base = __ldg(fp.offsets + lo);
          L = __ldg(fp.lengths + lo);
          o0 = (g - __ldg(fp.fstart + lo)) * fp.hop - fp.pad;
          gi = g;
        }
        s_base[tid] = base; s_o0[tid] = o0; s_len[tid] = L; s_g[tid] = gi;
      }
      bar_prod();
      r_base = s_base[row]; r_o0 = s_o0[row]; r_len = s_len[row];
      lap(0);
      // ---- phase A: operand tiles; the staging loads of three K steps are in flight while one is converted ----
      raw_t buf0[4], buf1[4], buf2[4], buf3[4];
      gather(0, buf0);
      if (1 < KS) gather(1, buf1);
      if (2 < KS) gather(2, buf2);
      auto kstep = [&](const int ks, raw_t (&raw)[4], raw_t (&ld)[4]) {
        const uint32_t s = ksg & (kDfStages - 1);
        float cur[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) cur[e] = df_value<DT>(raw[e]);
        if (ks + 3 < KS) gather(ks + 3, ld);
        lap(1);
        if (ksg >= kDfStages) dwait(2, &bar_free[s], ((ksg / kDfStages) - 1) & 1);
        lap(2);
        {
          const __half2 h01 = __floats2half2_rn(cur[0], cur[1]), h23 = __floats2half2_rn(cur[2], cur[3]);
          const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
          const __half2 l01 = __floats2half2_rn(cur[0] - f01.x, cur[1] - f01.y), l23 = __floats2half2_rn(cur[2] - f23.x, cur[3] - f23.y);
          unsigned char* at = s_stage + s * stage_bytes + (q4 >> 1) * (16 * kDfRows) + row * 16 + (q4 & 1) * 8;
          *reinterpret_cast<uint2*>(at) = make_uint2(*reinterpret_cast<const unsigned*>(&h01), *reinterpret_cast<const unsigned*>(&h23));
          *reinterpret_cast<uint2*>(at + kDfATile) = make_uint2(*reinterpret_cast<const unsigned*>(&l01), *reinterpret_cast<const unsigned*>(&l23));
        }
        tc::fence_async_smem();
        bar_stage_arrive(s);                                                // does not wait: the issuer warp picks the tile up
        lap(3);
        ++ksg;
      };
#pragma unroll 1
      for (int ks = 0; ks < KS; ks += 4) {
        kstep(ks, buf0, buf3);
        if (ks + 1 < KS) kstep(ks + 1, buf1, buf0);
        if (ks + 2 < KS) kstep(ks + 2, buf2, buf1);
        if (ks + 3 < KS) kstep(ks + 3, buf3, buf2);
      }
      // ---- phase B: power -> TMEM, mel, log -> workspace ----
      const int gidx = s_g[frow];
      dwait(5, &bar_d, static_cast<uint32_t>((t - t_begin) & 1));
      tc::tc_fence_after();
      lap(6);
      {
        const int per_grp = ((NH / 8 + 3) / 4) * 8;                         // bins per group, a multiple of 8
        for (int k0 = grp * per_grp; k0 < min(NH, (grp + 1) * per_grp); k0 += 8) {
          float re[8], im[8];
          df_ld8(tm_lane + k0, re);
          df_ld8(tm_lane + NH + k0, im);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) re[i] = fmaf(re[i], re[i], im[i] * im[i]);
          df_st8(tm_lane + k0, re);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      tc::tc_fence_before();
      bar_prod();
      tc::tc_fence_after();
      lap(7);
      {
        auto emit = [&](const int code, const float m) {
          const int f = (code & 0xFFFF) - 1, mode = code >> 16;
          if (f < 0) return;
          if (mode == 0) {
            if (gidx >= 0) fp.lm[static_cast<long long>(f) * fp.lm_stride + gidx] = 3.01029995663981195f * df_log2(fmaxf(fp.amin, m));
          } else {
            s_slots[(mode - 1) * kDfRows + frow] = m;
          }
        };
        float pending = 0.0f;
        const int4* pp = s_pieces + my_pieces.x;
#pragma unroll 1
        for (int n = my_pieces.y; n > 0; --n, ++pp) {
          const int4 pc = *pp;                          // (first float4 of bins, steps, first weight float4, emit code)
          const float4* wp = s_wtab + pc.z;
          float a0 = 0.0f, a1 = 0.0f, b0 = 0.0f, b1 = 0.0f;
#pragma unroll 1
          for (int k = 0; k < pc.y; ++k) {
            float v[4];
            df_ld4(tm_lane + 4 * (pc.x + k), v);
            tc::tmem_ld_wait();
            const float4 w01 = wp[2 * k], w23 = wp[2 * k + 1];
            a0 = fmaf(w01.x, v[0], a0); b0 = fmaf(w01.y, v[0], b0);
            a1 = fmaf(w01.z, v[1], a1); b1 = fmaf(w01.w, v[1], b1);
            a0 = fmaf(w23.x, v[2], a0); b0 = fmaf(w23.y, v[2], b0);
            a1 = fmaf(w23.z, v[3], a1); b1 = fmaf(w23.w, v[3], b1);
          }
          emit(pc.w, pending + (a0 + a1));
          pending = b0 + b1;
        }
      }
      tc::tc_fence_before();
      if (t + 1 < t_end) asm volatile("bar.arrive 6, %0;" ::"n"(kDfThreads) : "memory");   // TMEM may be overwritten
      bar_prod();                                       // boundary subtotals are visible
      lap(8);
      for (int idx = tid; idx < fp.tc_n_bnd * kDfRows; idx += kDfProducers) {
        const int4 e = s_bnd[idx >> 7];                 // (filter, first slot, slots, -)
        const int r = idx & (kDfRows - 1);
        float m = 0.0f;
        for (int i = 0; i < e.z; ++i) m += s_slots[(e.y + i) * kDfRows + r];
        if (s_g[r] >= 0) fp.lm[static_cast<long long>(e.x) * fp.lm_stride + s_g[r]] = 3.01029995663981195f * df_log2(fmaxf(fp.amin, m));
      }
      bar_prod();                                       // rows / slots may be rewritten by the next tile
      lap(9);
    }
    if (prof) {
      volatile int* d = fp.tc_dbg;
      for (int i = 0; i < 10; ++i) d[6 + i] = static_cast<int>(tk[i] >> 4);
      __threadfence_system();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp_u == 0) tc::tmem_dealloc(tmem, 512);
}

template <int DT>
static cudaError_t launch_df(const FParams& fp, int sm_count, int smem_bytes, cudaStream_t stream) {
  static int granted[kMaxDevices] = {0};
  const cudaError_t eg = ensure_dyn_smem(tcdft_kernel<DT>, smem_bytes, 0, granted, true);
  if (eg != cudaSuccess) return eg;
  tcdft_kernel<DT><<<sm_count, kDfThreads, smem_bytes, stream>>>(fp);
  return cudaGetLastError();
}

cudaError_t launch_tcdft_path(const FParams& fp, int sm_count, int smem_bytes, int cep_smem_bytes, int max_frames, cudaStream_t stream) {
  cudaError_t e = launch_frame_prefix(fp, stream);
  if (e != cudaSuccess) return e;
  switch (fp.dtype) {
    case ASR_I16: e = launch_df<ASR_I16>(fp, sm_count, smem_bytes, stream); break;
    case ASR_F32: e = launch_df<ASR_F32>(fp, sm_count, smem_bytes, stream); break;
    default: e = launch_df<ASR_F64>(fp, sm_count, smem_bytes, stream); break;
  }
  if (e != cudaSuccess) return e;
  return launch_cepstra_tail(fp, cep_smem_bytes, max_frames, stream);
}

int tcdft_static_smem_bytes() { return kDfRows * (8 + 3 * 4) + 2 * kDfStages * 8 + 64; }
int tcdft_stages() { return kDfStages; }

}  // namespace asr
