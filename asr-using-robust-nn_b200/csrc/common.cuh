// Shared declarations of the asr_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <new>
#include <string>
#include <vector>
#include "../../include/asr_b200.h"

namespace asr {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kMaxSmemBytes = 227 * 1024;   // opt-in dynamic shared memory per CTA on sm_100a
constexpr int kMelChunkQuads = 4;           // a mel task covers at most 4 float4 = 16 bins

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
#define ASR_CUDA_TRY(expr)                                   \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return ::asr::cuda_fail(_e, #expr); \
  } while (0)

// One mel task = a run of <= kMelChunkQuads float4 weight groups of one filter.
struct MelTask {
  int filter;    // mel filter index
  int k_start;   // first FFT bin (multiple of 4)
  int n_quads;   // number of float4 groups
  int w_off;     // offset into the weight array, in float4 units
};

// Parameters of one fused-MFCC launch, passed by value (__grid_constant__).
struct KParams {
  // ---- batch ----
  const void* audio;
  const long long* offsets;
  const int* lengths;
  int dtype;
  int n_clips;
  // ---- fused noise ----
  int noise_mode;
  const double* z;
  const double* z2;
  const double* sigma;
  double mix_p, mix_s0, mix_s1;
  // ---- output ----
  void* out;
  int out_f64;
  int out_frames;
  int out_rows;
  int logmel_only;   // 1: write the clamped log-mel matrix instead of the MFCC rows
  int* status;
  // ---- plan scalars ----
  int n_fft, hop, pad, pad_mode, n_bins, n_mels, n_mfcc, delta_orders, delta_width;
  float top_db, amin, preemph;
  int lm_pitch, dct_pitch;
  int fft_path;       // 1: register FFT kernel for n_fft in {512,1024,2048}; 0: direct DFT
  int fb;             // frames per batch (frame buffers per CTA)
  int frame_stride;   // floats between frame buffers (frame_stride/4 odd)
  int chunk_cap;      // floats of staged audio per batch
  int n_tasks, n_tasks_padded, part_pitch, fixed_slots;
  // ---- table blob (global) and section offsets in floats ----
  const float4* blob;
  int blob_f4;
  int off_window, off_window_i16, off_twp, off_twu, off_tasks, off_melw, off_ftasks, off_fslots, off_dct, off_taps;
  // ---- dynamic shared-memory layout, offsets in floats ----
  int sm_frames, sm_part, sm_lm, sm_cbuf, sm_red;
  int t_cap;          // frame capacity of one CTA's log-mel buffer
  int cluster_size;   // CTAs per clip (thread-block cluster), 1..16
  int vec_ok;         // audio (and noise) pointers are 16-byte aligned: vector staging allowed
  int cbuf_pitch;
};

// host-side launcher (mfcc_kernel.cu)
cudaError_t launch_mfcc(const KParams& kp, int smem_bytes, cudaStream_t stream);
cudaError_t mfcc_kernel_init();   // opt-in shared memory attributes, once per process/device

}  // namespace asr

struct asr_plan {
  asr_mfcc_params prm;
  int win_length;
  int pad;
  int n_bins;
  int fft_path;       // 1 = register FFT, 0 = direct DFT
  int fb;
  int frame_stride;
  int chunk_cap;
  int lm_pitch, dct_pitch;
  int n_tasks, n_tasks_padded, part_pitch, fixed_slots;
  int off_window, off_window_i16, off_twp, off_twu, off_tasks, off_melw, off_ftasks, off_fslots, off_dct, off_taps;
  int blob_floats;
  float* blob_dev;
  int device;
  // host copies for table-level parity tests
  std::vector<float> h_window, h_mel_dense, h_dct, h_taps;
};
