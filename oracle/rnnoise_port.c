/* oracle/rnnoise_port.c -- CPU restatement of the reference hot path (see rnnoise_port.h).
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
 * Compile with:  gcc -O2 -ffp-contract=off -fPIC -shared  (no FMA contraction: the reference's
 * DSP objects are SSE2 builds; every fused multiply-add the reference really executes is written
 * here as an explicit fmaf()).
 */
#include "rnnoise_port.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* same operand orientation as the reference macros (src/arch.h:72-75) so NaNs propagate alike */
#define RMAX(a, b) ((a) > (b) ? (a) : (b))
#define RMIN(a, b) ((a) < (b) ? (a) : (b))

/* ------------------------------------------------------------------------------------------ */
/* Tables: regenerated from the closed forms in reference src/dump_rnnoise_tables.c:85,92-97 and */
/* src/kiss_fft.c:406-420 (twiddles), :352-404 (factor order 5,3,4,4,4 for nfft=960).            */
/* ------------------------------------------------------------------------------------------ */
static int g_tables_ready = 0;
static float g_half_window[RP_FRAME];
static float g_dct[RP_BANDS * RP_BANDS];
static rp_cpx g_tw[RP_WINDOW];
static int g_bitrev[RP_WINDOW];
/* rnnoise_tables.c:562 stores the forward scale as this literal (1/960 printed to 8 digits). */
static const float g_fft_scale = 0.0010416667f;

/* band edges in 50 Hz bins, reference src/denoise.c:63-65 */
static const int g_eband[RP_BANDS + 2] = {0, 2, 4, 6, 8, 10, 12, 15, 18, 21, 24, 28, 32, 36, 41, 47, 53,
                                          60, 68, 77, 87, 98, 110, 124, 140, 157, 176, 198, 223, 251,
                                          282, 317, 356, 400};

void rp_tables_init(void) {
  if (g_tables_ready) return;
  const double pi = 3.14159265358979323846264338327;
  for (int i = 0; i < RP_FRAME; i++) {
    double s = sin(.5 * M_PI * (i + .5) / RP_FRAME);
    g_half_window[i] = (float)sin(.5 * M_PI * s * s);
  }
  for (int i = 0; i < RP_BANDS; i++)
    for (int j = 0; j < RP_BANDS; j++) {
      g_dct[i * RP_BANDS + j] = (float)cos((i + .5) * j * M_PI / RP_BANDS);
      if (j == 0) g_dct[i * RP_BANDS + j] *= (float)sqrt(.5);
    }
  for (int k = 0; k < RP_WINDOW; k++) {
    double phase = (-2 * pi / RP_WINDOW) * k;
    g_tw[k].r = (float)cos(phase);
    g_tw[k].i = (float)sin(phase);
  }
  /* digit reversal for radices (5,3,4,4,4): input index i = j0 + 5 j1 + 15 j2 + 60 j3 + 240 j4 lands
     at 192 j0 + 64 j1 + 16 j2 + 4 j3 + j4 (kiss_fft.c compute_bitrev_table recursion). */
  for (int i = 0; i < RP_WINDOW; i++) {
    int j0 = i % 5, j1 = (i / 5) % 3, j2 = (i / 15) % 4, j3 = (i / 60) % 4, j4 = i / 240;
    g_bitrev[i] = 192 * j0 + 64 * j1 + 16 * j2 + 4 * j3 + j4;
  }
  g_tables_ready = 1;
}
const float *rp_half_window(void) { rp_tables_init(); return g_half_window; }
const float *rp_dct_table(void) { rp_tables_init(); return g_dct; }
const rp_cpx *rp_twiddles(void) { rp_tables_init(); return g_tw; }
const int *rp_bitrev(void) { rp_tables_init(); return g_bitrev; }

/* ------------------------------------------------------------------------------------------ */
/* 960-point forward FFT, scaled by 1/960.  Same butterflies and the same order of float       */
/* operations as reference src/kiss_fft.c: rnn_fft_c:566-586, kf_bfly4:101-170, kf_bfly3:173-229, */
/* kf_bfly5:232-316 (float build macros _kiss_fft_guts.h:106-151).                                */
/* ------------------------------------------------------------------------------------------ */
static inline rp_cpx cmul(rp_cpx a, rp_cpx b) {
  rp_cpx m;
  m.r = a.r * b.r - a.i * b.i;
  m.i = a.r * b.i + a.i * b.r;
  return m;
}
static inline rp_cpx cadd(rp_cpx a, rp_cpx b) { rp_cpx m = {a.r + b.r, a.i + b.i}; return m; }
static inline rp_cpx csub(rp_cpx a, rp_cpx b) { rp_cpx m = {a.r - b.r, a.i - b.i}; return m; }

static void radix4_first(rp_cpx *F) { /* m == 1: twiddles are all 1 */
  for (int g = 0; g < 240; g++, F += 4) {
    rp_cpx s0 = csub(F[0], F[2]);
    F[0] = cadd(F[0], F[2]);
    rp_cpx s1 = cadd(F[1], F[3]);
    F[2] = csub(F[0], s1);
    F[0] = cadd(F[0], s1);
    s1 = csub(F[1], F[3]);
    F[1].r = s0.r + s1.i; F[1].i = s0.i - s1.r;
    F[3].r = s0.r - s1.i; F[3].i = s0.i + s1.r;
  }
}
static void radix4(rp_cpx *F0, int m, int groups, int gstride, int fs) {
  for (int g = 0; g < groups; g++) {
    rp_cpx *F = F0 + g * gstride;
    for (int j = 0; j < m; j++, F++) {
      rp_cpx s0 = cmul(F[m], g_tw[j * fs]);
      rp_cpx s1 = cmul(F[2 * m], g_tw[2 * j * fs]);
      rp_cpx s2 = cmul(F[3 * m], g_tw[3 * j * fs]);
      rp_cpx s5 = csub(F[0], s1);
      F[0] = cadd(F[0], s1);
      rp_cpx s3 = cadd(s0, s2);
      rp_cpx s4 = csub(s0, s2);
      F[2 * m] = csub(F[0], s3);
      F[0] = cadd(F[0], s3);
      F[m].r = s5.r + s4.i; F[m].i = s5.i - s4.r;
      F[3 * m].r = s5.r - s4.i; F[3 * m].i = s5.i + s4.r;
    }
  }
}
static void radix3(rp_cpx *F0, int m, int groups, int gstride, int fs) {
  const float epi3 = g_tw[fs * m].i;
  for (int g = 0; g < groups; g++) {
    rp_cpx *F = F0 + g * gstride;
    for (int j = 0; j < m; j++, F++) {
      rp_cpx s1 = cmul(F[m], g_tw[j * fs]);
      rp_cpx s2 = cmul(F[2 * m], g_tw[2 * j * fs]);
      rp_cpx s3 = cadd(s1, s2);
      rp_cpx s0 = csub(s1, s2);
      F[m].r = F[0].r - s3.r * .5f;
      F[m].i = F[0].i - s3.i * .5f;
      s0.r *= epi3; s0.i *= epi3;
      F[0] = cadd(F[0], s3);
      F[2 * m].r = F[m].r + s0.i; F[2 * m].i = F[m].i - s0.r;
      F[m].r = F[m].r - s0.i; F[m].i = F[m].i + s0.r;
    }
  }
}
static void radix5(rp_cpx *F, int m, int fs) {
  const rp_cpx ya = g_tw[fs * m], yb = g_tw[fs * 2 * m];
  for (int u = 0; u < m; u++) {
    rp_cpx *F0 = F + u, *F1 = F0 + m, *F2 = F0 + 2 * m, *F3 = F0 + 3 * m, *F4 = F0 + 4 * m;
    rp_cpx s0 = *F0;
    rp_cpx s1 = cmul(*F1, g_tw[u * fs]);
    rp_cpx s2 = cmul(*F2, g_tw[2 * u * fs]);
    rp_cpx s3 = cmul(*F3, g_tw[3 * u * fs]);
    rp_cpx s4 = cmul(*F4, g_tw[4 * u * fs]);
    rp_cpx s7 = cadd(s1, s4), s10 = csub(s1, s4), s8 = cadd(s2, s3), s9 = csub(s2, s3);
    F0->r = F0->r + (s7.r + s8.r);
    F0->i = F0->i + (s7.i + s8.i);
    rp_cpx s5, s6, s11, s12;
    s5.r = s0.r + (s7.r * ya.r + s8.r * yb.r);
    s5.i = s0.i + (s7.i * ya.r + s8.i * yb.r);
    s6.r = s10.i * ya.i + s9.i * yb.i;
    s6.i = -(s10.r * ya.i + s9.r * yb.i);
    *F1 = csub(s5, s6);
    *F4 = cadd(s5, s6);
    s11.r = s0.r + (s7.r * yb.r + s8.r * ya.r);
    s11.i = s0.i + (s7.i * yb.r + s8.i * ya.r);
    s12.r = s9.i * ya.i - s10.i * yb.i;
    s12.i = s10.r * yb.i - s9.r * ya.i;
    *F2 = cadd(s11, s12);
    *F3 = csub(s11, s12);
  }
}
void rp_fft960(const rp_cpx *in, rp_cpx *out) {
  rp_tables_init();
  for (int i = 0; i < RP_WINDOW; i++) {
    out[g_bitrev[i]].r = g_fft_scale * in[i].r;
    out[g_bitrev[i]].i = g_fft_scale * in[i].i;
  }
  radix4_first(out);           /* p=4 m=1   fstride 240 */
  radix4(out, 4, 60, 16, 60);  /* p=4 m=4   */
  radix4(out, 16, 15, 64, 15); /* p=4 m=16  */
  radix3(out, 64, 5, 192, 5);  /* p=3 m=64  */
  radix5(out, 192, 1);         /* p=5 m=192 */
}

/* ------------------------------------------------------------------------------------------ */
/* Frame-level DSP (reference src/denoise.c)                                                     */
/* ------------------------------------------------------------------------------------------ */

/* rnn_biquad, denoise.c:409-419 with b={-2,1}, a={-1.99599,0.996} (:469-470): the bracketed terms
   are evaluated in double and rounded to float when stored into the state. */
void rp_biquad_hp(float *y, float mem[2], const float *x, int n) {
  const float b0 = -2.f, b1 = 1.f, a0 = -1.99599f, a1 = 0.99600f;
  for (int i = 0; i < n; i++) {
    float xi = x[i];
    float yi = x[i] + mem[0];
    mem[0] = (float)((double)mem[1] + ((double)b0 * (double)xi - (double)a0 * (double)yi));
    mem[1] = (float)((double)b1 * (double)xi - (double)a1 * (double)yi);
    y[i] = yi;
  }
}

/* compute_band_energy, denoise.c:90-113 */
void rp_band_energy(float *E, const rp_cpx *X) {
  float sum[RP_BANDS + 2] = {0};
  for (int b = 0; b < RP_BANDS + 1; b++) {
    int bs = g_eband[b + 1] - g_eband[b];
    for (int j = 0; j < bs; j++) {
      float frac = (float)j / bs;
      const rp_cpx v = X[g_eband[b] + j];
      float t = v.r * v.r;
      t += v.i * v.i;
      sum[b] += (1 - frac) * t;
      sum[b + 1] += frac * t;
    }
  }
  sum[1] = (sum[0] + sum[1]) * 2 / 3;
  sum[RP_BANDS] = (sum[RP_BANDS] + sum[RP_BANDS + 1]) * 2 / 3;
  for (int b = 0; b < RP_BANDS; b++) E[b] = sum[b + 1];
}

/* compute_band_corr, denoise.c:115-138 */
void rp_band_corr(float *E, const rp_cpx *X, const rp_cpx *P) {
  float sum[RP_BANDS + 2] = {0};
  for (int b = 0; b < RP_BANDS + 1; b++) {
    int bs = g_eband[b + 1] - g_eband[b];
    for (int j = 0; j < bs; j++) {
      float frac = (float)j / bs;
      int k = g_eband[b] + j;
      float t = X[k].r * P[k].r;
      t += X[k].i * P[k].i;
      sum[b] += (1 - frac) * t;
      sum[b + 1] += frac * t;
    }
  }
  sum[1] = (sum[0] + sum[1]) * 2 / 3;
  sum[RP_BANDS] = (sum[RP_BANDS] + sum[RP_BANDS + 1]) * 2 / 3;
  for (int b = 0; b < RP_BANDS; b++) E[b] = sum[b + 1];
}

/* interp_band_gain, denoise.c:140-154.  The reference only writes bins 0..399; its callers pass
   zero-initialised arrays (rf, normf: :425,428; gf={1}: :466 whose bin 0 is overwritten), so bins
   400..480 are 0 -- made explicit here. */
void rp_interp_band_gain(float *g, const float *band) {
  for (int k = 0; k < RP_FREQ; k++) g[k] = 0;
  for (int b = 1; b < RP_BANDS; b++) {
    int bs = g_eband[b + 1] - g_eband[b];
    for (int j = 0; j < bs; j++) {
      float frac = (float)j / bs;
      g[g_eband[b] + j] = (1 - frac) * band[b - 1] + frac * band[b];
    }
  }
  for (int j = 0; j < g_eband[1]; j++) g[j] = band[0];
  for (int j = g_eband[RP_BANDS]; j < g_eband[RP_BANDS + 1]; j++) g[j] = band[RP_BANDS - 1];
}

/* dct, denoise.c:160-170: float accumulation, final scale by the double constant sqrt(2/22) */
void rp_dct(float *out, const float *in) {
  rp_tables_init();
  for (int i = 0; i < RP_BANDS; i++) {
    float sum = 0;
    for (int j = 0; j < RP_BANDS; j++) sum += in[j] * g_dct[j * RP_BANDS + i];
    out[i] = (float)(sum * sqrt(2. / 22));
  }
}

static void window_fft(rp_cpx *out481, const float *x960) { /* apply_window:219 + forward_transform:186 */
  rp_cpx a[RP_WINDOW], y[RP_WINDOW];
  rp_tables_init();
  for (int i = 0; i < RP_WINDOW; i++) {
    int wi = i < RP_FRAME ? i : RP_WINDOW - 1 - i;
    a[i].r = x960[i] * g_half_window[wi];
    a[i].i = 0;
  }
  rp_fft960(a, y);
  for (int i = 0; i < RP_FREQ; i++) out481[i] = y[i];
}

/* ------------------------------------------------------------------------------------------ */
/* Pitch analysis (reference src/pitch.c, src/celt_lpc.c; float build of src/arch.h:189-249)    */
/* ------------------------------------------------------------------------------------------ */

static float dot_seq(const float *x, const float *y, int n) { /* celt_inner_prod, pitch.h:134 */
  float s = 0;
  for (int i = 0; i < n; i++) s = s + x[i] * y[i];
  return s;
}

/* rnn_pitch_downsample, pitch.c:146-214 (C==1): [.25 .5 .25] decimation, order-4 LPC whitening */
void rp_pitch_downsample(const float *buf, float *lp) {
  const int n = RP_PITCH_BUF >> 1; /* 864 */
  for (int i = 1; i < n; i++) lp[i] = .5f * (.5f * (buf[2 * i - 1] + buf[2 * i + 1]) + buf[2 * i]);
  lp[0] = .5f * (.5f * buf[1] + buf[0]);
  /* rnn_autocorr(lag 4), celt_lpc.c:92-174: lags via the sequential dot product over the first
     n-4 samples (rnn_pitch_xcorr, pitch.c:216) plus a separately summed tail (:145-151). */
  float ac[5];
  const int fastN = n - 4;
  for (int k = 0; k < 5; k++) {
    float s = dot_seq(lp, lp + k, fastN);
    float d = 0;
    for (int i = k + fastN; i < n; i++) d = d + lp[i] * lp[i - k];
    ac[k] = s + d;
  }
  ac[0] *= 1.0001f;                                                 /* pitch.c:188 */
  for (int i = 1; i <= 4; i++) ac[i] -= ac[i] * (.008f * i) * (.008f * i); /* :197 */
  /* rnn_lpc order 4, celt_lpc.c:38-89 */
  float lpc[4] = {0, 0, 0, 0};
  float error = ac[0];
  if (ac[0] != 0) {
    for (int i = 0; i < 4; i++) {
      float rr = 0;
      for (int j = 0; j < i; j++) rr += lpc[j] * ac[i - j];
      rr += ac[i + 1];
      float r = -rr / error;
      lpc[i] = r;
      for (int j = 0; j < (i + 1) >> 1; j++) {
        float t1 = lpc[j], t2 = lpc[i - 1 - j];
        lpc[j] = t1 + r * t2;
        lpc[i - 1 - j] = t2 + r * t1;
      }
      error = error - (r * r) * error;
      if (error < .001f * ac[0]) break;
    }
  }
  float tmp = 1.f;
  for (int i = 0; i < 4; i++) {
    tmp = .9f * tmp;
    lpc[i] = lpc[i] * tmp;
  }
  const float c1 = .8f;
  float num[5];
  num[0] = lpc[0] + .8f;
  num[1] = lpc[1] + c1 * lpc[0];
  num[2] = lpc[2] + c1 * lpc[1];
  num[3] = lpc[3] + c1 * lpc[2];
  num[4] = c1 * lpc[3];
  /* celt_fir5, pitch.c:104-143, in place with zero initial memory */
  float m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0;
  for (int i = 0; i < n; i++) {
    float xi = lp[i];
    float sum = xi;
    sum = sum + num[0] * m0;
    sum = sum + num[1] * m1;
    sum = sum + num[2] * m2;
    sum = sum + num[3] * m3;
    sum = sum + num[4] * m4;
    m4 = m3; m3 = m2; m2 = m1; m1 = m0; m0 = xi;
    lp[i] = sum;
  }
}

/* find_best_pitch, pitch.c:44-102 (float build) */
static void best_two(const float *xcorr, const float *y, int len, int max_pitch, int best[2]) {
  float Syy = 1;
  float bnum0 = -1, bnum1 = -1, bden0 = 0, bden1 = 0;
  best[0] = 0; best[1] = 1;
  for (int j = 0; j < len; j++) Syy = Syy + y[j] * y[j];
  for (int i = 0; i < max_pitch; i++) {
    if (xcorr[i] > 0) {
      float x16 = xcorr[i];
      x16 *= 1e-12f;
      float num = x16 * x16;
      if (num * bden1 > bnum1 * Syy) {
        if (num * bden0 > bnum0 * Syy) {
          bnum1 = bnum0; bden1 = bden0; best[1] = best[0];
          bnum0 = num; bden0 = Syy; best[0] = i;
        } else {
          bnum1 = num; bden1 = Syy; best[1] = i;
        }
      }
    }
    Syy += y[i + len] * y[i + len] - y[i] * y[i];
    Syy = RMAX(1, Syy);
  }
}

/* rnn_pitch_search(x_lp = lp+384, y = lp, len 960, max_pitch 588), pitch.c:281-385 */
int rp_pitch_search(const float *lp) {
  const float *x_lp = lp + (RP_PITCH_MAX >> 1);
  const float *y = lp;
  const int len = RP_PITCH_FRAME, max_pitch = RP_PITCH_MAX - 3 * RP_PITCH_MIN; /* 960, 588 */
  const int lag = len + max_pitch;
  float x4[RP_PITCH_FRAME >> 2], y4[(RP_PITCH_FRAME + RP_PITCH_MAX) >> 2];
  float xcorr[RP_PITCH_MAX >> 1];
  int best[2] = {0, 0};
  for (int j = 0; j < len >> 2; j++) x4[j] = x_lp[2 * j];
  for (int j = 0; j < lag >> 2; j++) y4[j] = y[2 * j];
  /* coarse: every lag is a sequential dot product (xcorr_kernel pitch.h:51 keeps per-lag order) */
  for (int i = 0; i < max_pitch >> 2; i++) xcorr[i] = dot_seq(x4, y4 + i, len >> 2);
  best_two(xcorr, y4, len >> 2, max_pitch >> 2, best);
  /* fine, 2x decimation, only around the two coarse winners */
  for (int i = 0; i < max_pitch >> 1; i++) {
    xcorr[i] = 0;
    if (abs(i - 2 * best[0]) > 2 && abs(i - 2 * best[1]) > 2) continue;
    float sum = dot_seq(x_lp, y + i, len >> 1);
    xcorr[i] = RMAX(-1, sum);
  }
  best_two(xcorr, y, len >> 1, max_pitch >> 1, best);
  int offset = 0;
  if (best[0] > 0 && best[0] < (max_pitch >> 1) - 1) {
    float a = xcorr[best[0] - 1], b = xcorr[best[0]], c = xcorr[best[0] + 1];
    if ((c - a) > .7f * (b - a)) offset = 1;
    else if ((a - c) > .7f * (b - c)) offset = -1;
  }
  return 2 * best[0] - offset;
}

static float pitch_gain(float xy, float xx, float yy) { /* compute_pitch_gain, pitch.c:416-419 */
  return (float)(xy / sqrt(1 + xx * yy));
}

/* rnn_remove_doubling(x=lp, maxperiod 768, minperiod 60, N 960, ...), pitch.c:423-528 */
float rp_remove_doubling(const float *lp, int *T0_, int prev_period, float prev_gain) {
  static const int second_check[16] = {0, 0, 3, 2, 3, 2, 5, 2, 3, 2, 3, 2, 5, 2, 3, 2};
  const int minperiod0 = RP_PITCH_MIN;
  const int maxperiod = RP_PITCH_MAX / 2, minperiod = RP_PITCH_MIN / 2, N = RP_PITCH_FRAME / 2;
  const float *x = lp + maxperiod;
  int T0 = *T0_ / 2;
  prev_period /= 2;
  if (T0 >= maxperiod) T0 = maxperiod - 1;
  int T = T0;
  float xx = 0, xy = 0;
  for (int i = 0; i < N; i++) { /* dual_inner_prod, pitch.h:117 */
    xx = xx + x[i] * x[i];
    xy = xy + x[i] * x[i - T0];
  }
  float yy_lookup[RP_PITCH_MAX / 2 + 1];
  yy_lookup[0] = xx;
  float yy = xx;
  for (int i = 1; i <= maxperiod; i++) {
    yy = yy + x[-i] * x[-i] - x[N - i] * x[N - i];
    yy_lookup[i] = RMAX(0, yy);
  }
  yy = yy_lookup[T0];
  float best_xy = xy, best_yy = yy;
  float g0 = pitch_gain(xy, xx, yy), g = g0;
  for (int k = 2; k <= 15; k++) {
    int T1 = (2 * T0 + k) / (2 * k);
    if (T1 < minperiod) break;
    int T1b;
    if (k == 2) T1b = (T1 + T0 > maxperiod) ? T0 : T0 + T1;
    else T1b = (2 * second_check[k] * T0 + k) / (2 * k);
    float xy1 = 0, xy2 = 0;
    for (int i = 0; i < N; i++) {
      xy1 = xy1 + x[i] * x[i - T1];
      xy2 = xy2 + x[i] * x[i - T1b];
    }
    xy = .5f * (xy1 + xy2);
    yy = .5f * (yy_lookup[T1] + yy_lookup[T1b]);
    float g1 = pitch_gain(xy, xx, yy);
    float cont;
    if (abs(T1 - prev_period) <= 1) cont = prev_gain;
    else if (abs(T1 - prev_period) <= 2 && 5 * k * k < T0) cont = .5f * prev_gain;
    else cont = 0;
    float a = .7f * g0 - cont;
    float thresh = .3f > a ? .3f : a;
    if (T1 < 3 * minperiod) {
      a = .85f * g0 - cont;
      thresh = .4f > a ? .4f : a;
    } else if (T1 < 2 * minperiod) { /* unreachable, kept for fidelity (pitch.c:497) */
      a = .9f * g0 - cont;
      thresh = .5f > a ? .5f : a;
    }
    if (g1 > thresh) {
      best_xy = xy; best_yy = yy; T = T1; g = g1;
    }
  }
  best_xy = RMAX(0, best_xy);
  float pg;
  if (best_yy <= best_xy) pg = 1.f;
  else pg = best_xy / (best_yy + 1);
  float xc[3];
  for (int k = 0; k < 3; k++) xc[k] = dot_seq(x, x - (T + k - 1), N);
  int offset;
  if ((xc[2] - xc[0]) > .7f * (xc[1] - xc[0])) offset = 1;
  else if ((xc[0] - xc[2]) > .7f * (xc[1] - xc[2])) offset = -1;
  else offset = 0;
  if (pg > g) pg = g;
  *T0_ = 2 * T + offset;
  if (*T0_ < minperiod0) *T0_ = minperiod0;
  return pg;
}

/* rnn_frame_analysis + rnn_compute_frame_features, denoise.c:332-398 (TRAINING==0) */
int rp_frame_features(rp_state *st, rp_cpx *X, rp_cpx *P, float *Ex, float *Ep, float *Exp,
                      float *features, const float *in) {
  float w[RP_WINDOW];
  memcpy(w, st->analysis_mem, sizeof(float) * RP_FRAME);
  memcpy(w + RP_FRAME, in, sizeof(float) * RP_FRAME);
  memcpy(st->analysis_mem, in, sizeof(float) * RP_FRAME);
  window_fft(X, w);
  rp_band_energy(Ex, X);
  memmove(st->pitch_buf, st->pitch_buf + RP_FRAME, sizeof(float) * (RP_PITCH_BUF - RP_FRAME));
  memcpy(st->pitch_buf + RP_PITCH_BUF - RP_FRAME, in, sizeof(float) * RP_FRAME);
  float lp[RP_PITCH_BUF >> 1];
  rp_pitch_downsample(st->pitch_buf, lp);
  int pitch_index = RP_PITCH_MAX - rp_pitch_search(lp);
  float gain = rp_remove_doubling(lp, &pitch_index, st->last_period, st->last_gain);
  st->last_period = pitch_index;
  st->last_gain = gain;
  for (int i = 0; i < RP_WINDOW; i++) w[i] = st->pitch_buf[RP_PITCH_BUF - RP_WINDOW - pitch_index + i];
  window_fft(P, w);
  rp_band_energy(Ep, P);
  rp_band_corr(Exp, X, P);
  for (int i = 0; i < RP_BANDS; i++) Exp[i] = (float)(Exp[i] / sqrt(.001 + Ex[i] * Ep[i]));
  rp_dct(&features[RP_BANDS], Exp);
  features[2 * RP_BANDS] = (float)(.01 * (pitch_index - 300));
  float logMax = -2, follow = -2, E = 0, Ly[RP_BANDS];
  for (int i = 0; i < RP_BANDS; i++) {
    float ly = (float)log10(1e-2 + Ex[i]);
    double f15 = follow - 1.5;
    double m1 = f15 > ly ? f15 : ly;
    float lm7 = logMax - 7;
    Ly[i] = (float)(lm7 > m1 ? lm7 : m1);
    logMax = logMax > Ly[i] ? logMax : Ly[i];
    follow = (float)(f15 > Ly[i] ? f15 : Ly[i]);
    E += Ex[i];
  }
  if (E < 0.04) {
    memset(features, 0, sizeof(float) * RP_FEATURES);
    return 1;
  }
  rp_dct(features, Ly);
  features[0] -= 12;
  features[1] -= 4;
  return 0;
}

/* The frame loop of the reference's training-data tool (src/dump_features.c:466-491, a -DTRAINING=1 build of
 * denoise.c): rnn_frame_analysis on the clean state, rnn_compute_frame_features on the noisy state with the
 * TRAINING differences -- X and Y low-passed at `lowpass` bins (denoise.c:340-343), no silence short-circuit
 * (:389), return value E < 0.1 (:397) -- then the ideal band gains with their masks (dump_features.c:471-478).
 * rec[98] = features | g | vad_target.  Pinned against oracle/_ref/librnnoise_ref_training.so by
 * tests/test_oracle_port.py::test_port_training_frame_matches_reference_training_build. */
int rp_train_frame(rp_state *clean_st, rp_state *noisy_st, const float *clean, const float *noisy, float vad_target,
                   int noise_free, int lowpass, int band_lp, float *rec) {
  float w[RP_WINDOW], Ex[RP_BANDS], Ey[RP_BANDS], Ep[RP_BANDS], Exp[RP_BANDS];
  rp_cpx X[RP_FREQ], Y[RP_FREQ], P[RP_FREQ];
  float *features = rec, *g = rec + RP_FEATURES;
  /* clean frame: analysis only */
  memcpy(w, clean_st->analysis_mem, sizeof(float) * RP_FRAME);
  memcpy(w + RP_FRAME, clean, sizeof(float) * RP_FRAME);
  memcpy(clean_st->analysis_mem, clean, sizeof(float) * RP_FRAME);
  window_fft(Y, w);
  for (int i = lowpass; i < RP_FREQ; i++) Y[i].r = Y[i].i = 0;
  rp_band_energy(Ey, Y);
  /* noisy frame: the feature path of rp_frame_features with the TRAINING differences */
  rp_state *st = noisy_st;
  memcpy(w, st->analysis_mem, sizeof(float) * RP_FRAME);
  memcpy(w + RP_FRAME, noisy, sizeof(float) * RP_FRAME);
  memcpy(st->analysis_mem, noisy, sizeof(float) * RP_FRAME);
  window_fft(X, w);
  for (int i = lowpass; i < RP_FREQ; i++) X[i].r = X[i].i = 0;
  rp_band_energy(Ex, X);
  memmove(st->pitch_buf, st->pitch_buf + RP_FRAME, sizeof(float) * (RP_PITCH_BUF - RP_FRAME));
  memcpy(st->pitch_buf + RP_PITCH_BUF - RP_FRAME, noisy, sizeof(float) * RP_FRAME);
  float lp[RP_PITCH_BUF >> 1];
  rp_pitch_downsample(st->pitch_buf, lp);
  int pitch_index = RP_PITCH_MAX - rp_pitch_search(lp);
  float gain = rp_remove_doubling(lp, &pitch_index, st->last_period, st->last_gain);
  st->last_period = pitch_index;
  st->last_gain = gain;
  for (int i = 0; i < RP_WINDOW; i++) w[i] = st->pitch_buf[RP_PITCH_BUF - RP_WINDOW - pitch_index + i];
  window_fft(P, w);
  rp_band_energy(Ep, P);
  rp_band_corr(Exp, X, P);
  for (int i = 0; i < RP_BANDS; i++) Exp[i] = (float)(Exp[i] / sqrt(.001 + Ex[i] * Ep[i]));
  rp_dct(&features[RP_BANDS], Exp);
  features[2 * RP_BANDS] = (float)(.01 * (pitch_index - 300));
  float logMax = -2, follow = -2, E = 0, Ly[RP_BANDS];
  for (int i = 0; i < RP_BANDS; i++) {
    float ly = (float)log10(1e-2 + Ex[i]);
    double f15 = follow - 1.5;
    double m1 = f15 > ly ? f15 : ly;
    float lm7 = logMax - 7;
    Ly[i] = (float)(lm7 > m1 ? lm7 : m1);
    logMax = logMax > Ly[i] ? logMax : Ly[i];
    follow = (float)(f15 > Ly[i] ? f15 : Ly[i]);
    E += Ex[i];
  }
  rp_dct(features, Ly);
  features[0] -= 12;
  features[1] -= 4;
  const int quiet = E < 0.1;
  for (int i = 0; i < RP_BANDS; i++) {
    g[i] = (float)sqrt((Ey[i] + 1e-3) / (Ex[i] + 1e-3));
    if (g[i] > 1) g[i] = 1;
    if (quiet || i > band_lp) g[i] = -1;
    if (Ey[i] < 5e-2 && Ex[i] < 5e-2) g[i] = -1;
    if (vad_target == 0 && noise_free) g[i] = -1;
  }
  rec[RP_FEATURES + RP_BANDS] = vad_target;
  return quiet;
}

/* rnn_pitch_filter, denoise.c:421-455 */
void rp_pitch_filter(rp_cpx *X, const rp_cpx *P, const float *Ex, const float *Ep, const float *Exp,
                     const float *g) {
  float r[RP_BANDS], rf[RP_FREQ], newE[RP_BANDS], norm[RP_BANDS], normf[RP_FREQ];
  for (int i = 0; i < RP_BANDS; i++) {
    if (Exp[i] > g[i]) r[i] = 1;
    else {
      float e2 = Exp[i] * Exp[i], g2 = g[i] * g[i];
      r[i] = (float)((e2 * (1 - g2)) / (.001 + g2 * (1 - e2)));
    }
    float c = 0 > r[i] ? 0 : r[i];
    c = 1 < c ? 1 : c;
    r[i] = (float)sqrt(c);
    r[i] = (float)(r[i] * sqrt(Ex[i] / (1e-8 + Ep[i])));
  }
  rp_interp_band_gain(rf, r);
  for (int i = 0; i < RP_FREQ; i++) {
    X[i].r += rf[i] * P[i].r;
    X[i].i += rf[i] * P[i].i;
  }
  rp_band_energy(newE, X);
  for (int i = 0; i < RP_BANDS; i++) norm[i] = (float)sqrt(Ex[i] / (1e-8 + newE[i]));
  rp_interp_band_gain(normf, norm);
  for (int i = 0; i < RP_FREQ; i++) {
    X[i].r *= normf[i];
    X[i].i *= normf[i];
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Network (reference src/rnn.c:44-60, src/nnet.c:57-123, src/nnet_arch.h:79-162,                */
/* AVX2 kernels src/vec_avx.h:326-341, 398-445, 672-877)                                         */
/* ------------------------------------------------------------------------------------------ */

/* tanh8_approx, vec_avx.h:398-416, with an exact reciprocal in place of _mm256_rcp_ps */
float rp_tanh(float x) {
  const float N0 = 952.52801514f, N1 = 96.39235687f, N2 = 0.60863042f;
  const float D0 = 952.72399902f, D1 = 413.36801147f, D2 = 11.88600922f;
  float x2 = x * x;
  float num = fmaf(fmaf(N2, x2, N1), x2, N0);
  float den = fmaf(fmaf(D2, x2, D1), x2, D0);
  num = num * x;
  den = 1.f / den;
  num = num * den;
  num = num < 1.f ? num : 1.f;
  return num > -1.f ? num : -1.f;
}
/* sigmoid8_approx, vec_avx.h:426-445, same substitution */
float rp_sigmoid(float x) {
  const float N0 = 238.13200378f, N1 = 6.02452230f, N2 = 0.00950985f;
  const float D0 = 952.72399902f, D1 = 103.34200287f, D2 = 0.74287558f;
  float x2 = x * x;
  float num = fmaf(fmaf(N2, x2, N1), x2, N0);
  float den = fmaf(fmaf(D2, x2, D1), x2, D0);
  num = num * x;
  den = 1.f / den;
  num = fmaf(num, den, .5f);
  num = num < 1.f ? num : 1.f;
  return num > 0.f ? num : 0.f;
}
/* vector_ps_to_epi8, vec_avx.h:326-341: u8 = sat(rne(fma(x,127,127))) */
unsigned char rp_quant_u8(float x) {
  float f = fmaf(x, 127.f, 127.f);
  long v = lrintf(f); /* round-to-nearest-even under the default rounding mode */
  if (v < 0) v = 0;
  if (v > 255) v = 255;
  return (unsigned char)v;
}

/* compute_linear_avx2, nnet_arch.h:130-162.  acc_out (optional) receives the exact int32
   accumulators of an int8 layer. */
void rp_linear(const rp_layer *l, float *out, const float *in, int *acc_out) {
  const int M = l->nb_in, N = l->nb_out;
  if (l->is_int8) {
    unsigned char u[4096];
    for (int k = 0; k < M; k++) u[k] = rp_quant_u8(in[k]);
    for (int o = 0; o < N; o++) {
      int acc = 0;
      const signed char *w = l->w8 + (size_t)o * M;
      for (int k = 0; k < M; k++) acc += (int)u[k] * (int)w[k];
      if (acc_out) acc_out[o] = acc;
      out[o] = (float)acc * l->scale[o];   /* cvtepi32_ps, mul_ps: vec_avx.h:824-826 */
      out[o] = out[o] + l->subias[o];      /* USE_SU_BIAS, nnet_arch.h:146,149-151 */
    }
  } else {
    for (int o = 0; o < N; o++) {
      float y = 0;
      if (N >= 4) { /* sgemv vector rows: fused multiply-add, inputs in order (vec_avx.h:672-730) */
        for (int j = 0; j < M; j++) y = fmaf(l->wf[(size_t)j * N + o], in[j], y);
      } else {      /* scalar tail rows (vec_avx.h:731-735): compiled as separate mul and add */
        for (int j = 0; j < M; j++) y = y + l->wf[(size_t)j * N + o] * in[j];
      }
      out[o] = y + l->bias[o];
    }
  }
  if (l->diag) { /* nnet_arch.h:152-161; gcc contracts these to FMAs in the AVX2 object */
    for (int i = 0; i < M; i++) {
      out[i] = fmaf(l->diag[i], in[i], out[i]);
      out[i + M] = fmaf(l->diag[i + M], in[i], out[i + M]);
      out[i + 2 * M] = fmaf(l->diag[i + 2 * M], in[i], out[i + 2 * M]);
    }
  }
}

/* compute_generic_conv1d, nnet.c:113-123 */
static void conv1d(const rp_layer *l, float *out, float *mem, const float *in, int in_size, int is_tanh) {
  float tmp[4096];
  int hist = l->nb_in - in_size;
  memcpy(tmp, mem, sizeof(float) * hist);
  memcpy(tmp + hist, in, sizeof(float) * in_size);
  rp_linear(l, out, tmp, NULL);
  for (int i = 0; i < l->nb_out; i++) out[i] = is_tanh ? rp_tanh(out[i]) : out[i];
  memcpy(mem, tmp + in_size, sizeof(float) * hist);
}

/* compute_generic_gru, nnet.c:65-94 (gate order z, r, n) */
static void gru(const rp_layer *wi, const rp_layer *wr, float *state, const float *in) {
  const int N = wr->nb_in;
  float zrh[3 * RP_MAX_GRU], rec[3 * RP_MAX_GRU];
  rp_linear(wi, zrh, in, NULL);
  rp_linear(wr, rec, state, NULL);
  for (int i = 0; i < 2 * N; i++) zrh[i] += rec[i];
  for (int i = 0; i < 2 * N; i++) zrh[i] = rp_sigmoid(zrh[i]);
  float *z = zrh, *r = zrh + N, *h = zrh + 2 * N;
  for (int i = 0; i < N; i++) h[i] += rec[2 * N + i] * r[i];
  for (int i = 0; i < N; i++) h[i] = rp_tanh(h[i]);
  for (int i = 0; i < N; i++) h[i] = z[i] * state[i] + (1 - z[i]) * h[i];
  for (int i = 0; i < N; i++) state[i] = h[i];
}

/* compute_rnn, rnn.c:44-60 */
void rp_compute_rnn(const rp_model *m, rp_state *st, float *gains, float *vad, const float *features) {
  float tmp[RP_MAX_COND];
  float cat[4 * RP_MAX_GRU];
  const int G = m->gru;
  conv1d(&m->conv1, tmp, st->conv1_state, features, RP_FEATURES, 1);
  conv1d(&m->conv2, cat, st->conv2_state, tmp, m->cond, 1);
  gru(&m->gru_in[0], &m->gru_rec[0], st->gru_state[0], cat);
  gru(&m->gru_in[1], &m->gru_rec[1], st->gru_state[1], st->gru_state[0]);
  gru(&m->gru_in[2], &m->gru_rec[2], st->gru_state[2], st->gru_state[1]);
  memcpy(cat + G, st->gru_state[0], sizeof(float) * G);
  memcpy(cat + 2 * G, st->gru_state[1], sizeof(float) * G);
  memcpy(cat + 3 * G, st->gru_state[2], sizeof(float) * G);
  rp_linear(&m->dense_out, gains, cat, NULL);
  for (int i = 0; i < RP_BANDS; i++) gains[i] = rp_sigmoid(gains[i]);
  rp_linear(&m->vad_dense, vad, cat, NULL);
  *vad = rp_sigmoid(*vad);
}

/* ------------------------------------------------------------------------------------------ */
/* Model blob (format: reference src/nnet.h:43-62, src/write_weights.c:46-69; validation rules   */
/* src/parse_lpcnet_weights.c:37-52,98-176)                                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct { char head[4]; int version, type, size, block_size; char name[44]; } blob_head;
typedef struct { const char *name; int size; const void *data; } blob_arr;

static const blob_arr *find(const blob_arr *a, int n, const char *name) {
  for (int i = 0; i < n; i++) if (!strcmp(a[i].name, name)) return &a[i];
  return NULL;
}
static const void *find_sized(const blob_arr *a, int n, const char *layer, const char *suffix, int size) {
  char nm[96];
  snprintf(nm, sizeof(nm), "%s%s", layer, suffix);
  const blob_arr *e = find(a, n, nm);
  return (e && e->size == size) ? e->data : NULL;
}

static int load_float_layer(rp_layer *l, const blob_arr *a, int n, const char *name, int nb_in, int nb_out) {
  memset(l, 0, sizeof(*l));
  l->nb_in = nb_in; l->nb_out = nb_out;
  l->bias = find_sized(a, n, name, "_bias", nb_out * 4);
  l->wf = find_sized(a, n, name, "_weights_float", nb_in * nb_out * 4);
  return (l->bias && l->wf) ? 0 : 1;
}
static int load_int8_layer(rp_layer *l, const blob_arr *a, int n, const char *name, int nb_in, int nb_out,
                           int sparse, int diag) {
  char nm[96];
  memset(l, 0, sizeof(*l));
  l->nb_in = nb_in; l->nb_out = nb_out; l->is_int8 = 1;
  l->bias = find_sized(a, n, name, "_bias", nb_out * 4);
  l->subias = find_sized(a, n, name, "_subias", nb_out * 4);
  l->scale = find_sized(a, n, name, "_scale", nb_out * 4);
  if (!l->bias || !l->subias || !l->scale) return 1;
  if (diag && !(l->diag = find_sized(a, n, name, "_weights_diag", nb_out * 4))) return 1;
  snprintf(nm, sizeof(nm), "%s_weights_int8", name);
  const blob_arr *w = find(a, n, nm);
  if (!w) return 1;
  l->w8 = calloc((size_t)nb_in * nb_out, 1);
  const signed char *src = w->data;
  if (!sparse) { /* [out/8][in/4][8 out][4 in], wexchange/c_export/common.py:59-61 */
    if (w->size != nb_in * nb_out) return 1;
    for (int ob = 0; ob < nb_out / 8; ob++)
      for (int ib = 0; ib < nb_in / 4; ib++)
        for (int o = 0; o < 8; o++)
          for (int i = 0; i < 4; i++)
            l->w8[(size_t)(ob * 8 + o) * nb_in + ib * 4 + i] = *src++;
  } else { /* per 8 outputs: [nblocks, pos...] + 32-byte blocks w[4*o+i], common.py:151-165 */
    snprintf(nm, sizeof(nm), "%s_weights_idx", name);
    const blob_arr *ix = find(a, n, nm);
    if (!ix) return 1;
    const int *idx = ix->data;
    int remain = ix->size / 4, total = 0, rows = nb_out;
    for (int ob = 0; remain > 0; ob++) {
      int nb = *idx++;
      if (remain < nb + 1 || ob * 8 >= nb_out) return 1;
      for (int b = 0; b < nb; b++) {
        int pos = *idx++;
        if (pos + 3 >= nb_in || (pos & 3)) return 1;
        if ((total + 1) * 32 > w->size) return 1;
        for (int o = 0; o < 8; o++)
          for (int i = 0; i < 4; i++)
            l->w8[(size_t)(ob * 8 + o) * nb_in + pos + i] = src[(size_t)total * 32 + 4 * o + i];
        total++;
      }
      rows -= 8;
      remain -= nb + 1;
    }
    if (rows != 0 || total * 32 != w->size) return 1;
  }
  return 0;
}

rp_model *rp_model_from_buffer(const void *blob, int len) {
  blob_arr arr[256];
  int n = 0;
  const unsigned char *p = blob;
  while (len > 0) { /* parse_record, parse_lpcnet_weights.c:37-52 */
    const blob_head *h = (const blob_head *)p;
    if (len < 64 || h->block_size < h->size || h->block_size > len - 64 || h->name[43] != 0 || h->size < 0 || n >= 256)
      return NULL;
    if (h->size > 0) { arr[n].name = h->name; arr[n].size = h->size; arr[n].data = p + 64; n++; }
    else return NULL;
    p += 64 + h->block_size;
    len -= 64 + h->block_size;
  }
  /* dims from array sizes */
  const blob_arr *c1b = find(arr, n, "conv1_bias"), *g1b = find(arr, n, "gru1_recurrent_bias");
  if (!c1b || !g1b) return NULL;
  int cond = c1b->size / 4, gru = g1b->size / 12;
  if (cond <= 0 || cond > RP_MAX_COND || gru <= 0 || gru > RP_MAX_GRU || (gru & 7) || (cond & 3)) return NULL;
  rp_model *m = calloc(1, sizeof(*m));
  m->cond = cond; m->gru = gru;
  int err = 0;
  err |= load_float_layer(&m->conv1, arr, n, "conv1", 3 * RP_FEATURES, cond);
  err |= load_int8_layer(&m->conv2, arr, n, "conv2", 3 * cond, gru, 0, 0);
  for (int k = 0; k < 3 && !err; k++) {
    char nm[32];
    snprintf(nm, sizeof(nm), "gru%d_input", k + 1);
    err |= load_int8_layer(&m->gru_in[k], arr, n, nm, gru, 3 * gru, 1, 0);
    snprintf(nm, sizeof(nm), "gru%d_recurrent", k + 1);
    err |= load_int8_layer(&m->gru_rec[k], arr, n, nm, gru, 3 * gru, 1, 1);
  }
  err |= load_float_layer(&m->dense_out, arr, n, "dense_out", 4 * gru, RP_BANDS);
  err |= load_float_layer(&m->vad_dense, arr, n, "vad_dense", 4 * gru, 1);
  if (err) { rp_model_free(m); return NULL; }
  return m;
}

rp_model *rp_model_from_file(const char *path) {
  FILE *f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  long len = ftell(f);
  fseek(f, 0, SEEK_SET);
  void *buf = malloc(len);
  if (fread(buf, 1, len, f) != (size_t)len) { fclose(f); free(buf); return NULL; }
  fclose(f);
  rp_model *m = rp_model_from_buffer(buf, (int)len);
  if (!m) { free(buf); return NULL; }
  m->blob_copy = buf;
  return m;
}

void rp_model_free(rp_model *m) {
  if (!m) return;
  free(m->conv2.w8);
  for (int k = 0; k < 3; k++) { free(m->gru_in[k].w8); free(m->gru_rec[k].w8); }
  free(m->blob_copy);
  free(m);
}

rp_state *rp_state_create(void) { return calloc(1, sizeof(rp_state)); }
void rp_state_destroy(rp_state *st) { free(st); }
int rp_state_size(void) { return (int)sizeof(rp_state); }

/* rnnoise_process_frame, denoise.c:457-504 */
float rp_process_frame(const rp_model *m, rp_state *st, float *out, const float *in, rp_trace *tr) {
  rp_cpx X[RP_FREQ], P[RP_FREQ];
  float x[RP_FRAME], Ex[RP_BANDS], Ep[RP_BANDS], Exp[RP_BANDS], features[RP_FEATURES];
  float g[RP_BANDS], gf[RP_FREQ];
  float vad = 0;
  memset(g, 0, sizeof(g));
  rp_biquad_hp(x, st->mem_hp_x, in, RP_FRAME);
  int silence = rp_frame_features(st, X, P, Ex, Ep, Exp, features, x);
  if (!silence) {
    rp_compute_rnn(m, st, g, &vad, features);
    if (tr) memcpy(tr->g_raw, g, sizeof(g));
    rp_pitch_filter(st->delayed_X, st->delayed_P, st->delayed_Ex, st->delayed_Ep, st->delayed_Exp, g);
    for (int i = 0; i < RP_BANDS; i++) {
      float a = .6f * st->lastg[i];
      g[i] = g[i] > a ? g[i] : a;
      double t = g[i] * (st->delayed_Ex[i] + 1e-3) / (Ex[i] + 1e-3);
      st->lastg[i] = (float)(1.f < t ? 1.f : t);
    }
    rp_interp_band_gain(gf, g);
    for (int i = 0; i < RP_FREQ; i++) {
      st->delayed_X[i].r *= gf[i];
      st->delayed_X[i].i *= gf[i];
    }
  } else if (tr) memset(tr->g_raw, 0, sizeof(g));
  /* frame_synthesis:400-407 + inverse_transform:200-217 */
  {
    rp_cpx a[RP_WINDOW], y[RP_WINDOW];
    float t[RP_WINDOW];
    for (int i = 0; i < RP_FREQ; i++) a[i] = st->delayed_X[i];
    for (int i = RP_FREQ; i < RP_WINDOW; i++) {
      a[i].r = a[RP_WINDOW - i].r;
      a[i].i = -a[RP_WINDOW - i].i;
    }
    rp_fft960(a, y);
    t[0] = RP_WINDOW * y[0].r;
    for (int i = 1; i < RP_WINDOW; i++) t[i] = RP_WINDOW * y[RP_WINDOW - i].r;
    const float *hw = rp_half_window();
    for (int i = 0; i < RP_FRAME; i++) {
      t[i] *= hw[i];
      t[RP_WINDOW - 1 - i] *= hw[i];
    }
    for (int i = 0; i < RP_FRAME; i++) out[i] = t[i] + st->synthesis_mem[i];
    memcpy(st->synthesis_mem, t + RP_FRAME, sizeof(float) * RP_FRAME);
  }
  memcpy(st->delayed_X, X, sizeof(X));
  memcpy(st->delayed_P, P, sizeof(P));
  memcpy(st->delayed_Ex, Ex, sizeof(Ex));
  memcpy(st->delayed_Ep, Ep, sizeof(Ep));
  memcpy(st->delayed_Exp, Exp, sizeof(Exp));
  if (tr) {
    memcpy(tr->xb, x, sizeof(x));
    memcpy(tr->X, X, sizeof(X)); memcpy(tr->P, P, sizeof(P));
    memcpy(tr->Ex, Ex, sizeof(Ex)); memcpy(tr->Ep, Ep, sizeof(Ep)); memcpy(tr->Exp, Exp, sizeof(Exp));
    memcpy(tr->features, features, sizeof(features));
    tr->silence = silence; tr->pitch = st->last_period; tr->pitch_gain = st->last_gain; tr->vad = vad;
  }
  return vad;
}
