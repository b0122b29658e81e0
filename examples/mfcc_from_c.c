/*
 * The C-ABI from plain C (what a cgo / JNI / ctypes binding sees): compute_mfcc_all_files of the reference
 * (Voice digit recogniton/extract_features_construct_dataset.py:144-150) on synthetic one-second clips with the reference's
 * own librosa.feature.mfcc defaults (:30), through asr_mfcc_batch_host.
 *
 *   gcc -std=c99 -Wall -Wextra -Werror -I include examples/mfcc_from_c.c -L asr-using-robust-nn_b200 -lasr_b200 \
 *       -Wl,-rpath,$PWD/asr-using-robust-nn_b200 -lm -o mfcc_from_c && ./mfcc_from_c
 *
 * Exit status 0: features computed (prints the first coefficients); 3: no usable CUDA device - the library has no CPU
 * fallback, it reports ASR_ERR_CUDA and a message; anything else: an error.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "asr_b200.h"

int main(void) {
  enum { N_CLIPS = 4, SR = 22050, OUT_FRAMES = 44 };
  asr_mfcc_params prm;
  asr_plan* plan = NULL;
  int rc, b, i;

  memset(&prm, 0, sizeof(prm));
  prm.sr = SR; prm.n_fft = 2048; prm.win_length = 0; prm.hop_length = 512;
  prm.window = ASR_WIN_HANN; prm.center = 1; prm.pad_mode = ASR_PAD_REFLECT; prm.fftfreq_mode = ASR_FFTFREQ_LINSPACE;
  prm.n_mels = 128; prm.n_mfcc = 20; prm.fmin = 0.0f; prm.fmax = 0.0f; prm.top_db = 80.0f; prm.amin = 1e-10f;
  prm.lifter = 0.0f; prm.preemph = 0.0f; prm.delta_orders = 0; prm.delta_width = 9;

  if (asr_version() != ASR_B200_VERSION) { fprintf(stderr, "header / library version mismatch\n"); return 1; }
  rc = asr_plan_create(&prm, &plan);
  if (rc == ASR_ERR_CUDA) {
    printf("no CUDA device: %s (there is no CPU fallback)\n", asr_last_error());
    return 3;
  }
  if (rc != ASR_OK) { fprintf(stderr, "asr_plan_create: %s\n", asr_last_error()); return 1; }

  {
    const int rows = asr_plan_feature_rows(plan);
    float* audio = (float*)malloc(sizeof(float) * N_CLIPS * SR);
    int64_t offsets[N_CLIPS];
    int32_t lengths[N_CLIPS], status[N_CLIPS];
    double* out = (double*)malloc(sizeof(double) * N_CLIPS * rows * OUT_FRAMES);   /* the reference's float64 rows */
    if (!audio || !out) return 1;
    for (b = 0; b < N_CLIPS; ++b) {
      offsets[b] = (int64_t)b * SR;
      lengths[b] = SR;
      for (i = 0; i < SR; ++i)
        audio[b * SR + i] = 0.3f * (float)sin(2.0 * 3.14159265358979323846 * (220.0 * (b + 1)) * i / SR);
    }
    rc = asr_mfcc_batch_host(plan, audio, ASR_F32, offsets, lengths, N_CLIPS, 0, 0.0f, 0, out, ASR_F64, OUT_FRAMES, status);
    if (rc != ASR_OK) { fprintf(stderr, "asr_mfcc_batch_host: %s\n", asr_last_error()); return 1; }
    for (b = 0; b < N_CLIPS; ++b)
      printf("clip %d (status %d, %d frames): mfcc[0][0..2] = %.3f %.3f %.3f, mfcc[1][0] = %.3f\n", b, (int)status[b],
             (int)asr_plan_num_frames(plan, lengths[b]), out[(size_t)b * rows * OUT_FRAMES], out[(size_t)b * rows * OUT_FRAMES + 1],
             out[(size_t)b * rows * OUT_FRAMES + 2], out[(size_t)b * rows * OUT_FRAMES + OUT_FRAMES]);
    free(audio);
    free(out);
  }
  asr_plan_destroy(plan);
  return 0;
}
