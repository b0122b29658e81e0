/*
 * asr_b200.h - C-ABI of the B200-native (sm_100a) MFCC / noise-mix / standardisation path.
 *
 * The reference (fmazilu/ASR-using-robust-NN) is pure Python and has no FFI of its
 * own; its boundary for this path is a handful of module-level Python functions.
 * Each entry point below names the reference function (file:line) whose arithmetic it
 * replaces.  INTEGRATION.md shows the ctypes stub a maintainer adds on the reference
 * side.  Abbreviations: VDR = "Voice digit recogniton/", SR = "Speaker recognition/".
 *
 * Conventions
 *   - plain C, no torch / CUDA types in signatures: device and host buffers are `void*`
 *     or typed pointers, the stream is a `void*` holding a `cudaStream_t` (NULL = default).
 *   - every buffer is CALLER-allocated; the library never frees caller memory.
 *   - `*_dev` arguments are device pointers on the current CUDA device; launches are
 *     asynchronous and ordered on the given stream.  `asr_*_host` entry points take
 *     host pointers and are synchronous.
 *   - every function returns an `asr_status` (0 = ok, negative = error); the message
 *     of the last error on the calling thread is `asr_last_error()`.
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point
 *     returns ASR_ERR_CUDA.
 */
#ifndef ASR_B200_H_
#define ASR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASR_B200_VERSION 100 /* 0.1.0 */

typedef enum asr_status {
  ASR_OK = 0,
  ASR_ERR_INVALID = -1,     /* bad argument / unsupported parameter combination */
  ASR_ERR_CUDA = -2,        /* CUDA runtime error (see asr_last_error) */
  ASR_ERR_TOO_LARGE = -3,   /* a clip needs more shared memory than one SM has */
  ASR_ERR_ALLOC = -4
} asr_status;

/* per-clip status written by asr_mfcc_batch (the reference raises a Python exception) */
enum {
  ASR_CLIP_OK = 0,
  ASR_CLIP_TOO_SHORT = 1,        /* reflect padding needs len > n_fft/2 (np.pad raises) */
  ASR_CLIP_TOO_FEW_FRAMES = 2    /* deltas need >= delta_width frames (librosa.feature.delta raises) */
};

typedef enum asr_dtype { ASR_I16 = 0, ASR_F32 = 1, ASR_F64 = 2 } asr_dtype;
typedef enum asr_window { ASR_WIN_HANN = 0, ASR_WIN_HAMMING = 1 } asr_window;
typedef enum asr_pad_mode { ASR_PAD_REFLECT = 0, ASR_PAD_CONSTANT = 1 } asr_pad_mode;
typedef enum asr_fftfreq_mode { ASR_FFTFREQ_LINSPACE = 0, ASR_FFTFREQ_RFFTFREQ = 1 } asr_fftfreq_mode;

/*
 * Every keyword of librosa.feature.mfcc / melspectrogram / stft that the reference's
 * call sites fix (VDR/extract_features_construct_dataset.py:30 - all defaults;
 * SR/extract_features_construct_dataset.py:227-228 - n_fft = win_length = 441,
 * hop_length = 220) or BASELINE.json's configs vary.  htk=False, mel norm='slaney',
 * power=2, ref=1, dct_type=2, dct norm='ortho' are fixed (the reference never
 * changes them).
 */
typedef struct asr_mfcc_params {
  int32_t sr;            /* sampling rate the mel scale is built for (22050) */
  int32_t n_fft;         /* 2048 */
  int32_t win_length;    /* 0 -> n_fft */
  int32_t hop_length;    /* 512 */
  int32_t window;        /* asr_window */
  int32_t center;        /* 1 */
  int32_t pad_mode;      /* asr_pad_mode; librosa 0.9: reflect */
  int32_t fftfreq_mode;  /* asr_fftfreq_mode; librosa 0.9: linspace */
  int32_t n_mels;        /* 128 */
  int32_t n_mfcc;        /* 20 */
  float fmin;            /* 0 */
  float fmax;            /* 0 -> sr/2 */
  float top_db;          /* 80; < 0 -> no clamp */
  float amin;            /* 1e-10 */
  float lifter;          /* 0 -> none */
  float preemph;         /* 0 -> none (librosa.effects.preemphasis coef otherwise) */
  int32_t delta_orders;  /* 0, 1 (append delta) or 2 (append delta and delta-delta) */
  int32_t delta_width;   /* 9 */
} asr_mfcc_params;

/*
 * Additive noise fused in front of the MFCC, applied to the decoded audio exactly as
 * the reference does before calling librosa.feature.mfcc:
 *   WHITE    x + sigma[b]*z          VDR/attacks.py:73-86 (add_white_noise) and :222-245
 *                                    (add_white_noise_with_snr, sigma[b] from the SNR chain)
 *   MIXTURE  x + (|q|<p ? s1:s0)*g   VDR/attacks.py:145-183 (mixtgauss / add_noise)
 * z / q / g are float64 standard-normal streams laid out like the audio (same offsets).
 * The MFCC arithmetic is float32, so the fused mix delivers the float32 rounding of the reference's float64
 * signal: every path (CLIP, FRAMES, TILES) mixes in float64 with two separately rounded operations (no FMA) and
 * rounds once - bit-equal to float32(reference signal) (asr_plan_set_stage_probe reads the staged samples back).
 * The standalone asr_mix_* kernels, whose OUTPUT is the float64 noisy signal, are bit-exact as well.
 */
typedef enum asr_noise_mode { ASR_NOISE_NONE = 0, ASR_NOISE_WHITE = 1, ASR_NOISE_MIXTURE = 2 } asr_noise_mode;

typedef struct asr_noise {
  int32_t mode;              /* asr_noise_mode */
  int32_t reserved;
  const double* z_dev;       /* WHITE: z ; MIXTURE: selector stream q */
  const double* z2_dev;      /* MIXTURE: carrier stream g */
  const double* sigma_dev;   /* WHITE: per-clip sigma [n_clips] */
  double p;                  /* MIXTURE: probability threshold  */
  double sigma0;             /* MIXTURE: background sigma (alpha) */
  double sigma1;             /* MIXTURE: impulse sigma (10*alpha) */
} asr_noise;

typedef struct asr_plan asr_plan;

int asr_version(void);
const char* asr_last_error(void);
int asr_device_count(void);

/* ---- plan: immutable tables (window, FFT twiddles, sparse mel bank, DCT*lifter, delta taps)
 *      built on the host in float64, stored in float32 on the current device. ------------- */
int asr_plan_create(const asr_mfcc_params* params, asr_plan** plan_out);
void asr_plan_destroy(asr_plan* plan);
/* T = 1 + (len + 2*(n_fft/2) - n_fft)/hop for center=1 (0 when the clip cannot be framed) */
int32_t asr_plan_num_frames(const asr_plan* plan, int64_t length);
/* n_mfcc * (1 + delta_orders) */
int32_t asr_plan_feature_rows(const asr_plan* plan);
/* 1 = radix-2 register FFT (n_fft in {512,1024,2048}); 0 = direct DFT (any n_fft, e.g. 441) */
int32_t asr_plan_uses_fft(const asr_plan* plan);
/* copies of the float32 tables, for table-level parity tests (host pointers; NULL to skip) */
int asr_plan_get_tables(const asr_plan* plan, float* window /*n_fft*/, float* mel_dense /*n_mels*(1+n_fft/2)*/,
                        float* dct /*n_mfcc*n_mels, lifter folded in*/, float* delta_taps /*orders*width*/);

/*
 * Batched fused MFCC.  Replaces the per-file loop + librosa.feature.mfcc call of
 *   extract_features / compute_mfcc_all_files   VDR/extract_features_construct_dataset.py:24-39,144-150
 *   load_audio_dataset_and_labels               SR/extract_features_construct_dataset.py:224-232
 *   black_box_attack_on_audio[_snr]             VDR/attacks.py:89-121,248-274 ; SR/attacks.py:97-146,254-295
 * One launch: [noise mix] -> [pre-emphasis] -> reflect-pad framing -> window -> real FFT ->
 * power -> sparse mel -> 10*log10 -> clip-wide top_db clamp -> DCT-II(ortho) -> lifter ->
 * [delta, delta-delta] -> truncate / zero-pad to out_frames.
 *
 *   audio_dev    packed samples of dtype `dtype`; clip b = audio[offsets[b] .. offsets[b]+lengths[b])
 *   offsets_dev  int64 [n_clips] element offsets, lengths_dev int32 [n_clips]
 *   max_length   max over lengths (host knows it; sizes shared memory)
 *   noise        NULL or a descriptor (device pointers)
 *   out_dev      [n_clips][rows][out_frames] of out_dtype (ASR_F32 or ASR_F64), rows = feature_rows;
 *                flattening a clip's block row-major gives the reference's (n_mfcc*T,) row
 *                (VDR/extract...py:149).  Frames >= T are zero (VDR/extract...py:36-37).
 *   status_dev   int32 [n_clips] or NULL
 *   workspace_dev / workspace_bytes
 *                scratch of at least asr_mfcc_workspace_bytes(plan, n_clips, max_length) bytes (16-byte aligned):
 *                flattened frame index, clip maxima and the log-mel rows between the two stages of the n_fft = 512
 *                path.  NULL: the plan's own scratch is used (grown with cudaMalloc on demand - then calls on
 *                one plan must not overlap in time; pass a workspace per stream to run them concurrently).
 */
int asr_mfcc_batch(const asr_plan* plan, const void* audio_dev, int32_t dtype, const int64_t* offsets_dev,
                   const int32_t* lengths_dev, int32_t n_clips, int32_t max_length, const asr_noise* noise,
                   void* out_dev, int32_t out_dtype, int32_t out_frames, int32_t* status_dev,
                   void* workspace_dev, size_t workspace_bytes, void* stream);

/* Bytes of scratch the two entry points above/below need for such a batch (0: none needed). */
size_t asr_mfcc_workspace_bytes(const asr_plan* plan, int32_t n_clips, int32_t max_length);
/* Kernel path of asr_mfcc_batch / asr_logmel_batch.  Three implementations exist for n_fft = 512:
 *   ASR_PATH_CLIP    one CTA (or cluster) per clip, everything in one launch (the only path for other n_fft);
 *   ASR_PATH_FRAMES  block-pipelined: frame prefix -> frames (persistent, all clips' frames as one list) -> cepstra;
 *                    samples staged through registers or cp.async (any even hop, pre-emphasis, mixture noise);
 *   ASR_PATH_TILES   the same pipeline with the raw samples (and their float64 noise) brought in by TMA bulk copies
 *                    (cp.async.bulk + mbarrier), a transposed log-mel workspace and the clip maximum taken by the
 *                    cepstra kernel; needs hop and n_fft/2 multiples of 8, no pre-emphasis, no mixture noise and
 *                    16-byte aligned arrays (clips that do not start on a 16-byte boundary are staged from global memory).
 *   ASR_PATH_TC      tensor-core pipeline for int16 audio (hop a multiple of 32): tiles of 128 frames = the rows of a tcgen05
 *                    accumulator in TMEM.  The window, the first (stride-16) DFT pass and the inter-pass twiddle run as
 *                    float16 MMAs with a two-term split of both operands (float32-level accuracy), the second pass, the
 *                    unpack and the mel stage on the CUDA cores with one frame per thread straight out of TMEM.  Falls back
 *                    to TILES when its conditions do not hold.
 * ASR_PATH_AUTO (default) takes TILES when its conditions hold, else FRAMES when noise is fused into the launch
 * (each sample and its noise are then read and mixed once instead of once per overlapping frame), else CLIP.
 * Tests force every path. */
typedef enum asr_path { ASR_PATH_AUTO = 0, ASR_PATH_CLIP = 1, ASR_PATH_FRAMES = 2, ASR_PATH_TILES = 3, ASR_PATH_TC = 4 } asr_path;
int asr_plan_set_path(asr_plan* plan, int32_t path);
/* Path a call with 16-byte aligned arrays of `dtype` and noise mode `noise_mode` takes (ASR_PATH_CLIP/FRAMES/TILES). */
int32_t asr_plan_path_used(const asr_plan* plan, int32_t dtype, int32_t noise_mode);
/* Kernel launches one asr_mfcc_batch call makes with this plan: 1 (CLIP) or 3 (FRAMES, TILES); `noisy` as in the call. */
int32_t asr_plan_launches(const asr_plan* plan, int32_t noisy);

/* Parity probe of the fused noise mix (TILES path): while `staged_dev` is non-NULL, every launch of this plan on the
 * TILES path also writes the float32 frame samples it staged - decoded audio with the noise mixed in, BEFORE the window -
 * to staged_dev[offsets[b] + i] (float32, packed exactly like the audio; int16 audio is reported in [-1, 1) units).
 * With white noise these equal float32(add_white_noise_with_snr(...)) of VDR/attacks.py:222-245 bit for bit.  NULL = off. */
int asr_plan_set_stage_probe(asr_plan* plan, float* staged_dev);

/* Stage-level probe for parity tests: the clamped log-mel matrix [n_clips][n_mels][out_frames] (float32). */
int asr_logmel_batch(const asr_plan* plan, const void* audio_dev, int32_t dtype, const int64_t* offsets_dev,
                     const int32_t* lengths_dev, int32_t n_clips, int32_t max_length, const asr_noise* noise,
                     float* out_dev, int32_t out_frames, int32_t* status_dev,
                     void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- noise path (VDR/attacks.py:73-86,145-183,222-245 ; SR/attacks.py:81-94,149-189,228-251) ---- */

/* P[b] = np.mean(sample**2) in float32, in numpy's pairwise summation order - bit-exact
 * (VDR/attacks.py:234).  audio dtype ASR_I16 (value/32768) or ASR_F32. */
int asr_clip_power(const void* audio_dev, int32_t dtype, const int64_t* offsets_dev, const int32_t* lengths_dev,
                   int32_t n_clips, float* power_dev, void* stream);

/* A small transfer between MAPPED pinned host memory (cudaHostAlloc / cudaMallocHost: device-accessible at the same
 * address under unified addressing) and device memory, in either direction, done by a kernel instead of a copy-engine
 * operation.  The 4*B-byte power read-back and the 8*B-byte sigma upload between asr_clip_power and the noisy asr_mfcc_batch
 * (the host evaluates VDR/attacks.py:235-241 in between) otherwise queue behind the bulk audio upload of the next batch on
 * the same DMA engine.  Pointers and byte count must be multiples of 4; asynchronous on `stream`. */
int asr_copy_mapped(const void* src, void* dst, size_t bytes, void* stream);

/* sigma[b] = sqrt(10**((10*log10(P[b]) - snr_db)/10)) with every step rounded to float32
 * (VDR/attacks.py:235-241 under numpy >= 2 scalar rules), evaluated on the device in float64
 * and rounded; the host-exact alternative is to run those four numpy lines on P. */
int asr_snr_sigma(const float* power_dev, float target_snr_db, double* sigma_dev, int32_t n_clips, void* stream);

/* The same chain on the HOST, bit-exact with the reference text under numpy >= 2 (every step float32; `10 ** x` is
 * libm powf there).  HOST pointers, no CUDA call.  `log10_power_host` (may be NULL = glibc log10f) carries log10(P) as
 * the caller's numpy computed it: np.log10 is not glibc's log10f on AVX-512 hosts (numpy's own SIMD kernel), so a
 * Python caller passes np.log10(P) to stay bit-equal with the reference on ITS host.  This is the default of the
 * Python pipeline (asr_b200.pipeline): power on the device -> 4*B bytes to the host -> this call -> sigma back. */
int asr_snr_sigma_host(const float* power_host, const float* log10_power_host, float target_snr_db,
                       double* sigma_host, int32_t n_clips);

/* out = float64(x) + sigma[b]*z   (two roundings, no FMA) - VDR/attacks.py:84-85, :241-244 */
int asr_mix_white(const void* audio_dev, int32_t dtype, const int64_t* offsets_dev, const int32_t* lengths_dev,
                  int32_t n_clips, const double* z_dev, const double* sigma_dev, double* out_dev, void* stream);

/* Babble noise of BASELINE configs[1] ("white/babble").  The reference has NO babble implementation (parity unpinned);
 * the recipe is SURVEY.md 8(d): babble[b][n] = sum over k = 1..talkers of clip (b + k*stride) mod n_clips at sample n
 * (float64, exact), power[b] = mean(babble[b]^2) in float64 (fixed order: per 2048-sample chunk 256 thread sums of 8 samples and a binary tree, chunks ascending).  The mix is the white-noise formula with
 * z := babble and sigma[b] := sigma_snr[b] / sqrt(power[b]) (host), i.e. asr_mix_white / the fused launch take the
 * stream unchanged: noise power = P / 10^(snr/10) by the same sigma law as VDR/attacks.py:233-241.
 * babble_dev is packed like the audio (float64), power_dev is float64 [n_clips]. */
size_t asr_babble_workspace_bytes(int32_t n_clips, int32_t max_length);   /* scratch (kept in the ABI; since round 2 a clip's chunks
                                                                             are summed inside its CTA and the scratch is not touched) */
int asr_babble_stream(const void* audio_dev, int32_t dtype, const int64_t* offsets_dev, const int32_t* lengths_dev,
                      int32_t n_clips, int32_t max_length, int32_t stride, int32_t talkers, double* babble_dev,
                      double* power_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* out = float64(x) + (|q|<p ? sigma1 : sigma0)*g - VDR/attacks.py:159-181 */
int asr_mix_mixture(const void* audio_dev, int32_t dtype, const int64_t* offsets_dev, const int32_t* lengths_dev,
                    int32_t n_clips, const double* q_dev, const double* g_dev, double p, double sigma0,
                    double sigma1, double* out_dev, void* stream);

/* Feature-domain variants on an (n_rows, n_cols) float64 matrix, one sigma for all rows
 * (add_white_noise_on_dataset / add_noise_mixture_on_dataset, VDR/attacks.py:186-219). */
int asr_mix_rows_white(const double* x_dev, int64_t n, const double* z_dev, double sigma, double* out_dev, void* stream);
int asr_mix_rows_mixture(const double* x_dev, int64_t n, const double* q_dev, const double* g_dev, double p,
                         double sigma0, double sigma1, double* out_dev, void* stream);

/* Seeded standard-normal stream generated on the device (Philox4x32-10 + Box-Muller),
 * element i depends only on (seed, first_index + i): shard- and batch-size independent. */
int asr_randn_f64(uint64_t seed, uint64_t first_index, int64_t n, double* out_dev, void* stream);

/* ---- standardisation ("CMVN"): standardize_dataset, VDR/attacks.py:48-69 (StandardScaler) ----
 * Two passes like sklearn's _incremental_mean_and_var:
 *   pass 1  acc1[c]   += sum_r x[r][c]
 *   pass 2  acc2[c]   += sum_r (x[r][c]-mean[c]) ; acc2[n_cols+c] += sum_r (x[r][c]-mean[c])^2
 * The accumulators are float64 device vectors the caller zeroes and - when rows are sharded
 * over GPUs - all-reduces (sum) between the passes.  x is row-major with leading dimension ld. */
int asr_cmvn_colsum(const void* x_dev, int32_t dtype, int64_t n_rows, int32_t n_cols, int64_t ld,
                    double* acc1_dev, void* stream);
int asr_cmvn_mean(const double* acc1_dev, int64_t n_total, int32_t n_cols, double* mean_dev, void* stream);
int asr_cmvn_colsum_centered(const void* x_dev, int32_t dtype, int64_t n_rows, int32_t n_cols, int64_t ld,
                             const double* mean_dev, double* acc2_dev, void* stream);
/* var = acc2sq/n - (acc2/n)^2 ; scale = sqrt(var), 1 where sklearn's _is_constant_feature holds */
int asr_cmvn_finalize(const double* acc2_dev, const double* mean_dev, int64_t n_total, int32_t n_cols,
                      double* var_dev, double* scale_dev, void* stream);
/* out[r][c] = (x[r][c]-mean[c])/scale[c] ; out dtype ASR_F32 or ASR_F64, leading dimension n_cols */
int asr_cmvn_apply(const void* x_dev, int32_t dtype, int64_t n_rows, int32_t n_cols, int64_t ld,
                   const double* mean_dev, const double* scale_dev, void* out_dev, int32_t out_dtype, void* stream);

/* ---- standardisation, fused form: the same two passes and summation order in 3 launches on one GPU and with ONE
 * exchange when the rows are sharded (standardize_dataset, VDR/attacks.py:48-69; with row noise: the MFCC-domain attacks
 * add_white_noise_on_dataset / add_noise_mixture_on_dataset of :186-219 followed by standardize_dataset, :433-491).
 *   asr_cmvn_partial_sums   pass 1 (column sums) or pass 2 (sums of x-m and (x-m)^2 about the LOCAL mean m, which every CTA
 *                           finishes from the pass-1 partials) of one row block into slab partials of the workspace; returns
 *                           the number of slabs written (>= 0) or a negative asr_status.  `slab_base` = slabs the earlier
 *                           blocks of this pass wrote (train / dev / test are three blocks), `n_local_rows` = rows of ALL blocks
 *                           of this rank, `n_slabs_pass1` = total slabs of pass 1 (pass 2 only).
 *   asr_cmvn_local_message  this rank's message [n, S (D), C (D), Q (D)] (3*n_cols + 1 float64).
 *   <exchange>              the one collective of the path: an all-gather of the messages (NCCL / peer memory; the caller's).
 *   asr_cmvn_merge          messages of all ranks in rank order -> mean / var / scale of the whole dataset (Chan's update
 *                           of the centred sums to the global mean; with one rank the identity), identical on every rank.
 *   asr_cmvn_apply2         out = (x [+ noise] - mean) / scale.  workspace_dev != NULL (single rank): mean / var / scale are
 *                           FINISHED inside this launch from the slab partials (no local_message / merge launch) and stored.
 * `row_noise` (NULL = none): asr_noise with mode WHITE (z_dev = z, sigma0 = sigma: x + sigma*z) or MIXTURE (z_dev = q,
 * z2_dev = g, p, sigma0, sigma1); the streams are contiguous [n_rows][n_cols] float64; two roundings, no FMA.
 * The workspace (asr_cmvn_workspace_bytes(n_cols) bytes, 8-byte aligned) belongs to the caller. */
size_t asr_cmvn_workspace_bytes(int32_t n_cols);
int32_t asr_cmvn_partial_sums(const void* x_dev, int32_t dtype, int64_t n_rows, int32_t n_cols, int64_t ld,
                              const asr_noise* row_noise, int32_t pass, int32_t n_slabs_pass1, int64_t n_local_rows,
                              int32_t slab_base, void* workspace_dev, size_t workspace_bytes, void* stream);
int asr_cmvn_local_message(const void* workspace_dev, size_t workspace_bytes, int32_t n_slabs_pass1, int32_t n_slabs_pass2,
                           int64_t n_local_rows, int32_t n_cols, double* msg_dev, void* stream);
/* The exchange as a C-ABI call for callers without torch (a ctypes / cgo stub on the reference side that owns an NCCL
 * communicator): ncclAllGather of this rank's message (3*n_cols + 1 float64) into msgs_dev[world][3*n_cols + 1] in rank
 * order, on `stream`.  `nccl_comm` is the caller's ncclComm_t.  libasr_b200.so does not link NCCL: the symbol is resolved
 * at run time from the NCCL library ALREADY LOADED in the process (the one that created the communicator), so versions
 * cannot mix; ASR_ERR_INVALID if no NCCL is loaded.  Replaces the reduction hidden in StandardScaler().fit of
 * standardize_dataset (VDR/attacks.py:48-69) when the rows are sharded over GPUs. */
int asr_cmvn_exchange_nccl(void* nccl_comm, const double* msg_dev, double* msgs_dev, int32_t n_cols, void* stream);
/* The exchange without NCCL: a one-shot all-gather over PEER MEMORY (NVLink / NVSwitch), one small kernel, capturable in a
 * CUDA graph (it takes no per-step argument).  Every rank allocates a region of asr_cmvn_p2p_region_bytes(world, n_cols) bytes,
 * ZEROED once, and maps it into its peers (cudaIpc*, cuMem fabric handles, or torch's symmetric memory - the caller's
 * plumbing); `peer_regions_dev` is a DEVICE array [world] with every rank's region as addressed from this process (entry
 * `rank` = the own region).  `state_dev`: this rank's LOCAL uint32[world] step counters, zeroed once together with the regions
 * (all ranks must have zeroed before the first call: one barrier at set-up).  Every rank calls this once per step; on return
 * (in stream order) msgs_dev[world][3*n_cols+1] holds all messages in rank order, like asr_cmvn_exchange_nccl. */
size_t asr_cmvn_p2p_region_bytes(int32_t world, int32_t n_cols);
int asr_cmvn_exchange_p2p(const double* msg_dev, int32_t n_cols, int32_t rank, int32_t world, void* const* peer_regions_dev,
                          uint32_t* state_dev, double* msgs_dev, void* stream);
int asr_cmvn_merge(const double* msgs_dev, int32_t world, int32_t n_cols, double* mean_dev, double* var_dev,
                   double* scale_dev, double* n_total_dev /* may be NULL */, void* stream);
int asr_cmvn_apply2(const void* x_dev, int32_t dtype, int64_t n_rows, int32_t n_cols, int64_t ld, const asr_noise* row_noise,
                    const void* workspace_dev, size_t workspace_bytes, int32_t n_slabs_pass1, int32_t n_slabs_pass2,
                    int64_t n_total_rows, double* mean_dev, double* var_dev, double* scale_dev, void* out_dev,
                    int32_t out_dtype, void* stream);

/* ---- audio ingest: the resampling step of librosa.load (VDR/extract...py:27, SR/extract...py:210) ----
 * librosa resamples every file to 22 050 Hz with resampy's `kaiser_best` table, which is not available; the
 * stand-in is scipy.signal.resample_poly semantics: Kaiser(beta)-windowed sinc of
 * 20*max(up,down)+1 taps, unit DC gain, float32 taps scaled by up, zero-padded in front so that output sample 0
 * aligns with input sample 0; out[m] = sum_k taps[k] * x_up[(m + n_pre_remove)*down - k]. */
int64_t asr_resample_out_len(int64_t n_in, int32_t up, int32_t down);   /* ceil(n_in*up/down) after reducing up/down */
/* Host-side filter design.  taps_host == NULL: only the sizes are returned.  up/down come back reduced by their gcd. */
int asr_resample_design(int32_t up, int32_t down, double kaiser_beta /* scipy default 5.0 */, float* taps_host,
                        int32_t capacity, int32_t* up_out, int32_t* down_out, int32_t* n_taps_out,
                        int32_t* n_pre_remove_out);
/* Batched resampling on the device: clip b = in[in_offsets[b] .. +in_lengths[b]) (ASR_I16: value/32768, or ASR_F32)
 * -> out[out_offsets[b] .. + asr_resample_out_len(in_lengths[b], up, down)) float32.  up/down as returned by the design. */
int asr_resample_batch(const void* in_dev, int32_t dtype, const int64_t* in_offsets_dev, const int32_t* in_lengths_dev,
                       int32_t n_clips, int32_t max_in_length, int32_t up, int32_t down, const float* taps_dev,
                       int32_t n_taps, int32_t n_pre_remove, float* out_dev, const int64_t* out_offsets_dev,
                       void* stream);

/* ---- classifier forward pass of the accuracy-vs-SNR sweep ----
 * model.predict + np.argmax of VDR/attacks.py:409-414 for the Dense stack of VDR/train_constraints.py:63-88 (SR twin),
 * BatchNormalization folded into the following Dense layer by the caller:
 *   h_0 = x ; h_{l+1} = act(h_l W_l + b_l), act = ReLU except after the last layer ; probs = softmax(h_L).
 * One fused launch (a CTA takes 32 rows through all layers, activations in shared memory, weights streamed from L2),
 * float32 FMA accumulation in ascending k like a float32 predict.  x_dev: [n_rows] rows of dims[0] floats with leading
 * dimension ld_x; weights_dev[l]: [dims[l]][dims[l+1]] row-major (the layout of a Keras Dense kernel); biases_dev[l]:
 * [dims[l+1]]; dims_host, weights_dev, biases_dev are HOST arrays (of ints / of device pointers); 1..8 layers, layer outputs
 * 1..1024 wide, the input width bounded by shared memory (2020 for the speaker network fits).  logits_dev / probs_dev ([n_rows][dims[n_layers]]) and argmax_dev ([n_rows]) may each be NULL. */
int asr_mlp_forward(const float* x_dev, int64_t n_rows, int64_t ld_x, int32_t n_layers, const int32_t* dims_host,
                    const float* const* weights_dev, const float* const* biases_dev, float* logits_dev, float* probs_dev,
                    int32_t* argmax_dev, void* stream);

/* ---- measurement probes ----
 * FP32 FMA peak of the device (non-tensor CUDA cores): n_blocks x 512 threads x iters x 64 dependent-chain FMAs, 8 chains
 * per thread.  The caller times the launch with CUDA events: flops = n_blocks * 512 * iters * 64 * 2.  sink_dev receives
 * n_blocks * 512 floats.  bench.py quotes the FP32 roofline against this measured figure. */
int asr_fp32_peak_probe(int32_t n_blocks, int32_t iters, float* sink_dev, void* stream);

/* ---- diagnostics ----
 * Breadcrumb of the tensor-core kernel's bounded waits: word 0 = wait site that timed out (0 = none), 1 = CTA,
 * 2 = thread, 3 = parity.  Lives in mapped host memory, so it can be read after a failed launch. */
int32_t asr_plan_debug_word(const asr_plan* plan, int32_t i);
/*
 * Self-test of the tcgen05 / TMEM plumbing the tensor-core path of asr_mfcc_batch relies on:
 * D[128][32] = A1[128][32] * B1[32][32]^T + A2 * B2^T with float16 operands (row-major, K contiguous) and float32
 * accumulation; d_dev receives D twice (2 x 128 x 32 floats: block loads, then strided two-column loads). */
int asr_tc_selftest(const void* a1_dev, const void* b1_dev, const void* a2_dev, const void* b2_dev, float* d_dev,
                    void* stream);

/* ---- host-buffer entry points (what a ctypes binding on the reference side calls) ----------
 * Synchronous; host<->device copies are pipelined in chunks over two streams inside. */

/* compute_mfcc_all_files (VDR/extract...py:144-150) on decoded waveforms: out_host is
 * [n_clips][rows*out_frames] of out_dtype (ASR_F64 = the reference's np.zeros float64 rows).
 * snr_mode: 0 none; 1 = add_white_noise_with_snr with target_snr_db and the seeded device
 * normal stream (seed, clip b uses indices offsets[b]..); sigma chain evaluated on the device.
 * audio_host / out_host should be page-locked (cudaHostRegister) for the copies to overlap the kernels; offsets, lengths
 * and status may be pageable (staged in pinned memory of the plan).  Chunks of <= 32 MiB of samples / 4096 clips. */
int asr_mfcc_batch_host(const asr_plan* plan, const void* audio_host, int32_t dtype, const int64_t* offsets_host,
                        const int32_t* lengths_host, int32_t n_clips, int32_t snr_mode, float target_snr_db,
                        uint64_t seed, void* out_host, int32_t out_dtype, int32_t out_frames,
                        int32_t* status_host);

#ifdef __cplusplus
}
#endif
#endif /* ASR_B200_H_ */
