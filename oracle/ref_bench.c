/* oracle/ref_bench.c -- TEST/BENCH INFRASTRUCTURE ONLY (never on the product path).
 *
 * pthread multi-stream timing harness for the UNMODIFIED reference library built by
 * oracle/build_ref.py.  It only uses the public API of include/rnnoise.h
 * (rnnoise_model_from_filename / rnnoise_create / rnnoise_process_frame / rnnoise_destroy,
 * reference include/rnnoise.h:80-118) exactly like examples/rnnoise_demo.c:40-66 does for one
 * stream, but over S independent DenoiseStates split evenly across T worker threads.
 *
 * usage: ref_bench <model.bin> <streams> <steps> <warmup> <threads> [repeats [pcm.f32 pool_streams pool_frames]]
 *   one "step" = every stream advances by one 480-sample frame; the timed region of `steps` steps is
 *   repeated `repeats` times back to back (default 1; the states keep running), each repeat timed on its own.
 *   pcm.f32 (optional): float32 [pool_frames][pool_streams][480] PCM pool (int16 units) -- the very layout
 *   and content bench.py's GPU arm rotates through: stream s reads pool stream s % pool_streams, step f reads
 *   pool frame f % pool_frames.  Otherwise a deterministic voiced+noise signal is synthesised per stream.
 * prints one JSON line: {"frames_per_s": median over repeats, "best_frames_per_s":..., "repeat_frames_per_s": [...],
 *                        "elapsed_s": of the median repeat, "streams":..., "steps":..., "threads":..., "checksum":...}
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

#define MAX_REPEATS 16
typedef struct {
  int tid, nthreads, s0, s1, steps, warmup, in_frames, repeats, pool_streams;
  DenoiseState **st;
  RNNModel *model;
  const float *pcm; /* [streams][in_frames][480], or the pool [in_frames][pool_streams][480] when pool_streams > 0 */
  pthread_barrier_t *bar;
  double checksum;
  struct timespec t0[MAX_REPEATS], t1[MAX_REPEATS];
} Worker;

static const float *frame_of(const Worker *w, int s, int f) {
  if (w->pool_streams > 0) return w->pcm + ((size_t)(f % w->in_frames) * w->pool_streams + s % w->pool_streams) * FRAME;
  return w->pcm + ((size_t)s * w->in_frames + f % w->in_frames) * FRAME;
}

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
      rnnoise_process_frame(w->st[s], out, frame_of(w, s, f));
  for (int r = 0; r < w->repeats; r++) {
    pthread_barrier_wait(w->bar);
    clock_gettime(CLOCK_MONOTONIC, &w->t0[r]);
    for (int f = 0; f < w->steps; f++) {
      const int fi = w->warmup + r * w->steps + f;
      for (int s = w->s0; s < w->s1; s++) {
        cs += rnnoise_process_frame(w->st[s], out, frame_of(w, s, fi));
        cs += out[17] * 1e-6;
      }
    }
    clock_gettime(CLOCK_MONOTONIC, &w->t1[r]);
  }
  w->checksum = cs;
  return NULL;
}

static int cmp_double(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return x < y ? -1 : x > y; }

int main(int argc, char **argv) {
  if (argc < 6) {
    fprintf(stderr, "usage: %s model.bin streams steps warmup threads [repeats [pcm.f32 pool_streams pool_frames]]\n", argv[0]);
    return 2;
  }
  int S = atoi(argv[2]), steps = atoi(argv[3]), warmup = atoi(argv[4]), T = atoi(argv[5]);
  int repeats = argc > 6 ? atoi(argv[6]) : 1;
  if (repeats < 1) repeats = 1;
  if (repeats > MAX_REPEATS) repeats = MAX_REPEATS;
  if (T < 1) T = 1;
  if (T > S) T = S;
  RNNModel *model = rnnoise_model_from_filename(argv[1]);
  if (!model) { fprintf(stderr, "cannot load model\n"); return 1; }
  DenoiseState **st = calloc(S, sizeof(*st));
  int in_frames, pool_streams = 0;
  float *pcm;
  if (argc > 9) {
    pool_streams = atoi(argv[8]);
    in_frames = atoi(argv[9]);
    if (pool_streams < 1 || in_frames < 1) { fprintf(stderr, "bad pool shape\n"); return 2; }
    pcm = malloc(sizeof(float) * (size_t)pool_streams * in_frames * FRAME);
    FILE *f = fopen(argv[7], "rb");
    if (!f || fread(pcm, sizeof(float) * FRAME, (size_t)pool_streams * in_frames, f) != (size_t)pool_streams * in_frames) {
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
                    .steps = steps, .warmup = warmup, .in_frames = in_frames, .repeats = repeats, .pool_streams = pool_streams,
                    .st = st, .model = model, .pcm = pcm, .bar = &bar};
    pthread_create(&th[t], NULL, work, &w[t]);
  }
  double cs = 0;
  for (int t = 0; t < T; t++) {
    pthread_join(th[t], NULL);
    cs += w[t].checksum;
  }
  double rate[MAX_REPEATS], el[MAX_REPEATS], sorted[MAX_REPEATS];
  for (int r = 0; r < repeats; r++) {
    double tmin = 1e300, tmax = -1e300;
    for (int t = 0; t < T; t++) {
      double a = w[t].t0[r].tv_sec + 1e-9 * w[t].t0[r].tv_nsec, b = w[t].t1[r].tv_sec + 1e-9 * w[t].t1[r].tv_nsec;
      if (a < tmin) tmin = a;
      if (b > tmax) tmax = b;
    }
    el[r] = tmax - tmin;
    rate[r] = sorted[r] = (double)S * steps / el[r];
  }
  qsort(sorted, repeats, sizeof(double), cmp_double);
  const double med = sorted[(repeats - 1) / 2], best = sorted[repeats - 1];   /* lower median */
  double el_med = el[0];
  for (int r = 0; r < repeats; r++) if (rate[r] == med) el_med = el[r];
  printf("{\"frames_per_s\": %.1f, \"best_frames_per_s\": %.1f, \"repeat_frames_per_s\": [", med, best);
  for (int r = 0; r < repeats; r++) printf("%s%.1f", r ? ", " : "", rate[r]);
  printf("], \"elapsed_s\": %.6f, \"streams\": %d, \"steps\": %d, \"repeats\": %d, \"threads\": %d, \"pcm\": \"%s\", \"checksum\": %.6f}\n",
         el_med, S, steps, repeats, T, pool_streams ? "pool file" : "built-in synth", cs);
  for (int s = 0; s < S; s++) rnnoise_destroy(st[s]);
  rnnoise_model_free(model);
  return 0;
}
