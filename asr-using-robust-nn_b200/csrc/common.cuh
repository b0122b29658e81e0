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

// Opt-in dynamic shared memory above `floor_bytes`, remembered per device: cudaFuncSetAttribute applies to the current
// device only, and one process may drive several.
constexpr int kMaxDevices = 64;
template <typename K>
inline cudaError_t ensure_dyn_smem(K kernel, int bytes, int floor_bytes, int (&granted)[kMaxDevices], bool max_carveout = false) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  if (bytes > floor_bytes && bytes > granted[dev]) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && max_carveout)
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    granted[dev] = bytes;
  }
  return cudaSuccess;
}
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
  // chunk mode (clips whose log-mel matrix exceeds one CTA's shared memory): grid (clip, chunk of `t_cap` frames), no cluster;
  // a CTA writes its chunk's log-mel rows to the transposed workspace lm_global[filter][fstart[clip] + frame] and the
  // cepstra kernels of the block-pipelined paths finish the clip (clip maximum, clamp, DCT, deltas, padding)
  int chunk_mode;
  int n_chunks;       // chunks per clip of the longest clip
  float* lm_global;
  int lm_stride;
  const int* fstart;
  int vec_ok;         // audio (and noise) pointers are 16-byte aligned: vector staging allowed
  int cbuf_pitch;
};

// ---- block-pipelined path for n_fft = 512 (frames_kernel.cu) ----
constexpr int kFrWarps = 16;        // every warp: staging share, combine (2 slots), FFT (2 frames) per block
constexpr int kFrThreads = kFrWarps * 32;
constexpr int kFrMelWarps = 15;     // warps 0..14 share the mel segments; warp 15 assembles the block descriptors instead
constexpr int kFrAllThreads = kFrThreads;
constexpr int kFrBlock = 32;        // frames per block
constexpr int kFrMaxRuns = 4;       // runs of consecutive frames (of one clip) per block, at most

// ---- TMA-staged block pipeline for n_fft = 512 (tile_kernel.cu) ----
// Two shapes: 16 warps / 32 frames per block / one CTA per SM, or 8 warps / 16 frames per block / two CTAs per SM.
// The last warp of a CTA assembles descriptors and issues the copies; the other warps share the mel steps.
constexpr int kTlMaxRuns = 4;
constexpr int kTlRS = 548;          // floats per frame slot: 16 x 17 float2 exchange buffer, reused as the spectrum row
                                    // (RS/4 odd: float4 rows conflict-free over lanes; RS = 4 mod 8: slots 4 apart sit in complementary bank halves)
constexpr int kTlMaxPieces = 4;     // a mel segment is cut into at most this many per-warp pieces

constexpr int kCepSmallTab = 512;   // floats of DCT table carried in the kernel parameters (constant bank)

struct FParams {
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
  int out_f64, out_frames, out_rows, logmel_only;
  int* status;
  // ---- plan scalars ----
  int n_fft, hop, pad, pad_mode, n_mels, n_mfcc, delta_orders, delta_width;
  float top_db, amin, preemph;
  // ---- tables (global blob; the first blob_f4 float4 are copied to shared memory by the frames kernel) ----
  const float4* blob;
  int blob_f4;
  int off_window, off_twp, off_twu, off_wtab, off_pieces, off_wrange, off_frange, off_refs;
  // ---- frames kernel shared-memory layout (floats) ----
  int sm_aud, sm_S, sm_xb, sm_part, sm_raw;
  int async_stage;    // 1: samples are copied raw to shared memory with cp.async one block ahead (frames512_kernel<.., true>)
  int aud_cap, s_pitch, xb_stride, n_refs, max_runs, vec_ok;
  // ---- workspace ----
  float* lm;          // [total frames][lm_pitch] log-mel rows (unclamped)
  int lm_pitch;
  float* clipmax;     // [n_clips]
  int* fstart;        // [n_clips + 1] flattened index of each clip's first frame
  int* nframes;       // [n_clips]
  // ---- cepstra kernel tables: float4 offset inside the blob, size, shared-memory offsets (floats) ----
  int cep_blob_f4, cep_tab_f4, cep_off_cbuf, cep_off_taps;
  int dbg_skip;       // timing experiments only (ASR_B200_DBG_SKIP): bit 0 stage, 1 combine, 2 mel, 3 fft are skipped
  // ---- tiles path (tile_kernel.cu) ----
  int off_steps;      // int2 per mel step: (float4 index inside the spectrum row, partial row to flush into or -1)
  int t_npart;        // rows of 32 floats in the partial buffer: (n_mels + 1) * t_npc * 2
  int t_npc;          // pieces per mel segment
  int t_nw;           // warps per CTA: 16 (one CTA per SM) or 8 (two CTAs per SM)
  int t_smem_bytes;   // dynamic shared memory of the launch (bounds checks of the debug build)
  int lm_stride;      // tiles path: lm is [n_mels][lm_stride] (transposed), lm_stride >= total frames
  const float4* cep_blob;   // cepstra tables (global)
  int cep_off_col;    // cepstra_t_kernel: shared-memory offset (floats) of the [n_mels][128] log-mel column buffer
  int cep_small;      // 1: n_mels <= 32 and the transposed DCT table fits cep_dct: columns in registers, table from the constant bank
  float cep_dct[kCepSmallTab];   // [n_mels][4*NC4] transposed DCT x lifter (cep_small only)
  // ---- tensor-core path (tc_kernel.cu) ----
  const void* tc_mats;   // [16 b][MH1 | MH2][4 chunks][32 rows] x 16 bytes of float16: pass-1 matrices
  int tc_sm_hl, tc_sm_a, tc_sm_b, tc_sm_slots;   // byte offsets inside the dynamic shared memory
  int tc_hl_stride;   // uint2 entries between the residue rows of the staging array (1 mod 16: conflict-free both ways)
  int tc_hl_rows;     // capacity of the staging array in pair rows (32 samples each) per tile
  int tc_off_bnd;     // float offset of the boundary-filter list (int4: filter, first slot, slots, -) in the table blob
  int tc_n_bnd;       // entries of that list
  int* tc_dbg;        // mapped host memory (4 ints): breadcrumb of a wait that timed out
  // ---- tensor-core dense DFT (tcdft_kernel.cu); shares tc_mats, tc_sm_slots, tc_off_bnd, tc_n_bnd, tc_dbg ----
  int df_sm_stage;    // byte offset of the stage ring inside the dynamic shared memory
  int df_nh;          // accumulator columns per half (bins rounded up to 16)
  int df_ksteps;      // K steps of 16 samples (n_fft rounded up)
  int df_bslab;       // bytes of one K step's slab of B: [B1 | B2][2 chunks][2 * nh rows][16]
  float* stage_probe; // parity probe (asr_plan_set_stage_probe): staged samples written back, packed like the audio; or null
};

cudaError_t launch_frames_path(const FParams& fp, int sm_count, int frames_smem_bytes, int cep_smem_bytes, int max_frames,
                               cudaStream_t stream);
cudaError_t launch_frame_prefix(const FParams& fp, cudaStream_t stream);          // frames_kernel.cu
cudaError_t launch_tiles_path(const FParams& fp, int sm_count, int tile_smem_bytes, int cep_smem_bytes, int max_frames,
                              cudaStream_t stream);                              // tile_kernel.cu
cudaError_t launch_cepstra_tail(const FParams& fp, int cep_smem_bytes, int max_frames, cudaStream_t stream);   // tile_kernel.cu
cudaError_t launch_tc_path(const FParams& fp, int sm_count, int tc_smem_bytes, int cep_smem_bytes, int max_frames,
                           cudaStream_t stream);                                 // tc_kernel.cu
cudaError_t tc_upload_constants();                                               // tc_kernel.cu: unpack twiddles -> constant memory
int tc_static_smem_bytes();                                                      // tc_kernel.cu
cudaError_t launch_tcdft_path(const FParams& fp, int sm_count, int smem_bytes, int cep_smem_bytes, int max_frames,
                              cudaStream_t stream);                              // tcdft_kernel.cu
int tcdft_static_smem_bytes();
int tcdft_stages();

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
  // ---- block-pipelined path (n_fft = 512): tables and layout; fr_ok = 0 -> the per-clip kernel is used ----
  int fr_ok;
  int path;                // asr_path
  int sm_count;
  float* fr_blob_dev;
  int fr_blob_f4;          // part copied to shared memory by the frames kernel
  int fr_off_window, fr_off_window_i16, fr_off_twp, fr_off_twu, fr_off_wtab, fr_off_pieces, fr_off_wrange, fr_off_frange,
      fr_off_refs;
  int fr_n_refs, fr_s_pitch, fr_xb_stride, fr_lm_pitch;
  int cep_blob_f4, cep_tab_f4, cep_off_cbuf, cep_off_taps, cep_smem_bytes;
  std::vector<float> h_dct_t;   // transposed DCT x lifter [lm_pitch][4*NC4], as in the blob
  // ---- tiles path (n_fft = 512, TMA staging): tables; tl_ok = 0 -> not available for this plan ----
  int tl_ok;
  struct TileTables {   // one set per kernel shape (index 0: 16 warps, 1: 8 warps): the mel steps are dealt out differently
    float* blob_dev;
    int blob_f4;        // common tables, copied to shared memory; the two windows follow in the global blob
    int off_window, off_window_i16, off_twp, off_twu, off_wtab, off_steps, off_wrange;
    int npc, npart, nsteps;
  } tl[2];
  // plan-owned workspace for callers that pass none (grown on demand; not safe for concurrent launches)
  void* ws_dev;
  size_t ws_bytes;
  float* stage_probe;   // asr_plan_set_stage_probe
  void* host_state = nullptr;   // asr_mfcc_batch_host: persistent streams / device buffers / pinned descriptors (capi.cu: HostState)
  // ---- tensor-core path (n_fft = 512, int16 audio): pass-1 matrices and mel tables; tc_ok = 0 -> not available ----
  int tc_ok;
  void* tc_mats_dev;
  float* tc_blob_dev;
  // ---- tensor-core dense DFT (FFT sizes without a register FFT, e.g. 441): B slabs and mel tables; df_ok = 0 -> not available ----
  int df_ok, df_nh, df_ksteps, df_bslab;
  void* df_mats_dev;
  float* df_blob_dev;
  int df_blob_f4, df_off_wtab, df_off_pieces, df_off_wrange, df_off_bnd, df_n_bnd, df_n_slots;
  float* cep_dev;       // cepstra tables (transposed DCT x lifter, delta taps) for the cepstra_*_kernel of tile_kernel.cu
  int* tc_dbg_host;     // cudaHostAlloc(mapped): breadcrumb of a timed-out wait (asr_plan_debug_word)
  int* tc_dbg_dev;
  int tc_blob_f4, tc_off_wtab, tc_off_pieces, tc_off_wrange, tc_off_bnd, tc_n_bnd, tc_n_slots;
};
