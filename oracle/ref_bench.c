/* oracle/ref_bench.c -- TEST/BENCH INFRASTRUCTURE ONLY (never on the product path).
 *
 * pthread multi-stream timing harness for the UNMODIFIED reference library built by
 * oracle/build_ref.py.  It only uses the public API of include/rnnoise.h
 * (rnnoise_model_from_filename / rnnoise_create / rnnoise_process_frame / rnnoise_destroy,
 * reference include/rnnoise.h:80-118) exactly like examples/rnnoise_demo.c:40-66 does for one
 * stream, but over S independent DenoiseStates split evenly across T worker threads.
 *
 * usage: ref_bench <model.bin> <streams> <steps> <warmup> <threads> [pcm.f32]
 *   one "step" = every stream advances by one 480-sample frame.
 *   pcm.f32 (optional): float32 [streams][steps+warmup][480] host-generated PCM (int16 units);
 *   otherwise a deterministic voiced+noise signal is synthesised per stream.
 * prints one JSON line: {"frames_per_s":..., "elapsed_s":..., "streams":..., "steps":...,
 *                        "threads":..., "checksum":...}
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include "rnnoise.h"

#define FRAME 480
#define WINDOW_FRAMES 32 /* synthetic input is a 32-frame loop per stream */

typedef struct {
  int tid, nthreads, s0, s1, steps, warmup, in_frames;
  DenoiseState **st;
  RNNModel *model;
  const float *pcm; /* [streams][in_frames][480] */
  pthread_barrier_t *bar;
  double checksum;
  struct timespec t0, t1;
} Worker;

static void synth(float *dst, int stream, int frames) {
  /* harmonic source with gliding f0, gated at 1.5 Hz, plus LCG noise; int16 units */
  uint32_t lcg = 12345u + 977u * (uint32_t)stream;
  double f0 = 90.0 + (stream % 97) * 3.0, phase = 0.0;
  double amp = 1000.0 + (stream % 13) * 500.0;
  for (int n = 0; n < frames * FRAME; n++) {
    double t = n / 48000.0;
    double f = f0 * (1.0 + 0.2 * sin(2 * M_PI * 0.5 * t));
    phase += 2 * M_PI * f / 48000.0;
    double v = 0;
    for (int k = 1; k < 20; k++) v += sin(k * phase) / k;
    double gate = sin(2 * M_PI * 1.5 * t) > 0 ? 1.0 : 0.0;
    lcg = lcg * 1664525u + 1013904223u;
    double noise = ((int32_t)(lcg >> 8) % 2001 - 1000) * 0.4;
    dst[n] = (float)floor(amp * gate * v + noise);
  }
}

static void *work(void *arg) {
  Worker *w = (Worker *)arg;
  float out[FRAME];
  double cs = 0;
  cpu_set_t set;
  CPU_ZERO(&set);
  CPU_SET(w->tid % (int)sysconf(_SC_NPROCESSORS_ONLN), &set);
  pthread_setaffinity_np(pthread_self(), sizeof(set), &set); /* best effort */
  /* each worker creates (first-touches) its own states so they are NUMA-local to it */
  for (int s = w->s0; s < w->s1; s++) {
    w->st[s] = rnnoise_create(w->model);
    if (!w->st[s]) { fprintf(stderr, "rnnoise_create failed (model dims mismatch?)\n"); exit(1); }
  }
  for (int f = 0; f < w->warmup; f++)
    for (int s = w->s0; s < w->s1; s++)
      rnnoise_process_frame(w->st[s], out, w->pcm + ((size_t)s * w->in_frames + f % w->in_frames) * FRAME);
  pthread_barrier_wait(w->bar);
  clock_gettime(CLOCK_MONOTONIC, &w->t0);
  for (int f = 0; f < w->steps; f++) {
    int fi = (w->warmup + f) % w->in_frames;
    for (int s = w->s0; s < w->s1; s++) {
      cs += rnnoise_process_frame(w->st[s], out, w->pcm + ((size_t)s * w->in_frames + fi) * FRAME);
      cs += out[17] * 1e-6;
    }
  }
  clock_gettime(CLOCK_MONOTONIC, &w->t1);
  w->checksum = cs;
  return NULL;
}

int main(int argc, char **argv) {
  if (argc < 6) {
    fprintf(stderr, "usage: %s model.bin streams steps warmup threads [pcm.f32]\n", argv[0]);
    return 2;
  }
  int S = atoi(argv[2]), steps = atoi(argv[3]), warmup = atoi(argv[4]), T = atoi(argv[5]);
  if (T < 1) T = 1;
  if (T > S) T = S;
  RNNModel *model = rnnoise_model_from_filename(argv[1]);
  if (!model) { fprintf(stderr, "cannot load model\n"); return 1; }
  DenoiseState **st = calloc(S, sizeof(*st));
  int in_frames;
  float *pcm;
  if (argc > 6) {
    in_frames = steps + warmup;
    pcm = malloc(sizeof(float) * (size_t)S * in_frames * FRAME);
    FILE *f = fopen(argv[6], "rb");
    if (!f || fread(pcm, sizeof(float) * FRAME, (size_t)S * in_frames, f) != (size_t)S * in_frames) {
      fprintf(stderr, "cannot read pcm\n"); return 1;
    }
    fclose(f);
  } else {
    in_frames = WINDOW_FRAMES;
    pcm = malloc(sizeof(float) * (size_t)S * in_frames * FRAME);
    for (int s = 0; s < S; s++) synth(pcm + (size_t)s * in_frames * FRAME, s, in_frames);
  }
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, NULL, T);
  pthread_t *th = malloc(sizeof(*th) * T);
  Worker *w = calloc(T, sizeof(*w));
  for (int t = 0; t < T; t++) {
    w[t] = (Worker){.tid = t, .nthreads = T, .s0 = (int)((long)S * t / T), .s1 = (int)((long)S * (t + 1) / T),
                    .steps = steps, .warmup = warmup, .in_frames = in_frames, .st = st, .model = model, .pcm = pcm, .bar = &bar};
    pthread_create(&th[t], NULL, work, &w[t]);
  }
  double cs = 0, tmin = 1e300, tmax = -1e300;
  for (int t = 0; t < T; t++) {
    pthread_join(th[t], NULL);
    cs += w[t].checksum;
    double a = w[t].t0.tv_sec + 1e-9 * w[t].t0.tv_nsec, b = w[t].t1.tv_sec + 1e-9 * w[t].t1.tv_nsec;
    if (a < tmin) tmin = a;
    if (b > tmax) tmax = b;
  }
  double el = tmax - tmin;
  printf("{\"frames_per_s\": %.1f, \"elapsed_s\": %.6f, \"streams\": %d, \"steps\": %d, \"threads\": %d, \"checksum\": %.6f}\n",
         (double)S * steps / el, el, S, steps, T, cs);
  for (int s = 0; s < S; s++) rnnoise_destroy(st[s]);
  rnnoise_model_free(model);
  return 0;
}
