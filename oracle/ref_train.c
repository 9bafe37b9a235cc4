/* oracle/ref_train.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin driver around the UNMODIFIED reference compiled with -DTRAINING=1 (oracle/build_ref.py builds
 * denoise.c, pitch.c, kiss_fft.c, celt_lpc.c, rnnoise_tables.c in place): the frame loop of the
 * reference's training-data tool, src/dump_features.c:466-491, exposed as a function so the batched
 * GPU feature extractor can be checked record for record.  The reference tool keeps this loop inside
 * main() next to its file I/O and random mixing, so the three statements are restated here:
 *   rnn_frame_analysis(clean state)            dump_features.c:468
 *   rnn_compute_frame_features(noisy state)    dump_features.c:469
 *   ideal gains and masks                      dump_features.c:472-478
 * `lowpass` / `band_lp` are the globals denoise.c reads under TRAINING (denoise.c:327-330), defined by
 * dump_features.c:45-46 in the reference tool and by this file here. */
#include <math.h>
#include <stdlib.h>

#include "rnnoise.h"
#include "denoise.h"
#include "kiss_fft.h"

int lowpass = FREQ_SIZE;
int band_lp = NB_BANDS;

typedef struct {
  DenoiseState *clean, *noisy;
} RefTrain;

RefTrain *ref_train_create(void) {
  RefTrain *t = (RefTrain *)calloc(1, sizeof(*t));
  t->clean = rnnoise_create(NULL);
  t->noisy = rnnoise_create(NULL);
  return t;
}
void ref_train_destroy(RefTrain *t) {
  rnnoise_destroy(t->clean);
  rnnoise_destroy(t->noisy);
  free(t);
}

/* rec[98] = features | g | vad_target; dbg (optional) = X[962] P[962] Ex[32] Ep[32] Exp[32] Ey[32]; returns the quiet flag */
int ref_train_frame(RefTrain *t, const float *clean, const float *noisy, float vad_target, int noise_free,
                    int lp, int blp, float *rec, float *dbg) {
  kiss_fft_cpx X[FREQ_SIZE], Y[FREQ_SIZE], P[WINDOW_SIZE];
  float Ex[NB_BANDS], Ey[NB_BANDS], Ep[NB_BANDS], Exp[NB_BANDS];
  float *features = rec, *g = rec + NB_FEATURES;
  int i, quiet;
  lowpass = lp;
  band_lp = blp;
  rnn_frame_analysis(t->clean, Y, Ey, clean);
  quiet = rnn_compute_frame_features(t->noisy, X, P, Ex, Ep, Exp, features, noisy);
  for (i = 0; i < NB_BANDS; i++) {
    g[i] = sqrt((Ey[i] + 1e-3) / (Ex[i] + 1e-3));
    if (g[i] > 1) g[i] = 1;
    if (quiet || i > band_lp) g[i] = -1;
    if (Ey[i] < 5e-2 && Ex[i] < 5e-2) g[i] = -1;
    if (vad_target == 0 && noise_free) g[i] = -1;
  }
  rec[NB_FEATURES + NB_BANDS] = vad_target;
  if (dbg) {
    for (i = 0; i < FREQ_SIZE; i++) { dbg[2 * i] = X[i].r; dbg[2 * i + 1] = X[i].i; dbg[962 + 2 * i] = P[i].r; dbg[962 + 2 * i + 1] = P[i].i; }
    for (i = 0; i < NB_BANDS; i++) { dbg[1924 + i] = Ex[i]; dbg[1956 + i] = Ep[i]; dbg[1988 + i] = Exp[i]; dbg[2020 + i] = Ey[i]; }
  }
  return quiet;
}
