// dsp_tables.hpp -- host-side generation of the constant DSP tables from the reference's closed
// forms (src/dump_rnnoise_tables.c:85,92-97; src/kiss_fft.c:406-420; band edges src/denoise.c:63-65).
// Done once per engine in double precision with the C library, exactly as the reference's table
// generator does, then uploaded to global memory.
#pragma once
#include <math.h>
#include <string.h>
#include "dsp_core.cuh"

static inline void b200_fill_dsp_tables(DspTables *t) {
  static const short eband[NB_BANDS + 2] = {0, 2, 4, 6, 8, 10, 12, 15, 18, 21, 24, 28, 32, 36, 41, 47, 53,
                                            60, 68, 77, 87, 98, 110, 124, 140, 157, 176, 198, 223, 251,
                                            282, 317, 356, 400};
  const double pi = 3.14159265358979323846264338327;
  memset(t, 0, sizeof(*t));
  for (int i = 0; i < FRAME_SIZE; i++) {
    double s = sin(.5 * M_PI * (i + .5) / FRAME_SIZE);
    t->half_window[i] = (float)sin(.5 * M_PI * s * s);
  }
  for (int i = 0; i < NB_BANDS; i++)
    for (int j = 0; j < NB_BANDS; j++) {
      t->dct[i * NB_BANDS + j] = (float)cos((i + .5) * j * M_PI / NB_BANDS);
      if (j == 0) t->dct[i * NB_BANDS + j] *= (float)sqrt(.5);
    }
  for (int k = 0; k < WINDOW_SIZE; k++) {
    double phase = (-2 * pi / WINDOW_SIZE) * k;
    t->tw[k].r = (float)cos(phase);
    t->tw[k].i = (float)sin(phase);
  }
  for (int i = 0; i < WINDOW_SIZE; i++) {
    int j0 = i % 5, j1 = (i / 5) % 3, j2 = (i / 15) % 4, j3 = (i / 60) % 4, j4 = i / 240;
    t->bitrev[i] = (short)(192 * j0 + 64 * j1 + 16 * j2 + 4 * j3 + j4);
  }
  for (int b = 0; b < NB_BANDS + 2; b++) t->eband[b] = eband[b];
  for (int b = 0; b < NB_BANDS + 1; b++) {   // all 33 triangular segments (interp_bin only uses 1..31)
    int bs = eband[b + 1] - eband[b];
    for (int j = 0; j < bs; j++) {
      t->bin_band[eband[b] + j] = (unsigned char)b;
      t->bin_frac[eband[b] + j] = (float)j / bs;
      t->bin_cfrac[eband[b] + j] = 1 - t->bin_frac[eband[b] + j];
    }
  }
  t->fft_scale = 0.0010416667f;  // rnnoise_tables.c:562
}
