/* rnnoise_api.c -- the public C ABI of include/rnnoise.h, in plain C, on top of the CUDA engine.
 *
 * Mirrors the reference's API implementation (src/denoise.c:227-325, 457-504) entry point by entry
 * point; the arithmetic itself lives in the CUDA kernels (engine.cu).  There is no CPU path: every
 * creation call fails when the engine cannot be brought up on a GPU.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/rnnoise.h"
#include "engine.h"
#include "model_blob.h"

#define FRAME_SIZE 480
#define STATE_MAGIC 0x42323030 /* "B200" */

struct RNNModel {
  const void *const_blob; /* borrowed (from_buffer) */
  void *blob;             /* owned (from_file / from_filename) */
  int blob_len;
  B200HostModel host;     /* parsed + dense-expanded; valid iff parsed != 0 */
  int parsed;
};

struct RNNoiseBatch {
  B200Engine *engine;
  int nb_streams;
};

/* A single-stream state is a handle onto a private batch of one stream. */
struct DenoiseState {
  int magic;
  RNNoiseBatch *batch;
  RNNModel *owned_model; /* default model loaded on behalf of rnnoise_init(st, NULL) */
};

/* ------------------------------------------------------------------------------------------ */
static RNNModel *model_finish(RNNModel *m) {
  const void *p = m->blob ? m->blob : m->const_blob;
  if (b200_host_model_parse(&m->host, p, m->blob_len) != 0) {
    rnnoise_model_free(m);
    return NULL;
  }
  m->parsed = 1;
  return m;
}

RNNModel *rnnoise_model_from_buffer(const void *ptr, int len) {
  RNNModel *m;
  if (!ptr || len <= 0) return NULL;
  m = (RNNModel *)calloc(1, sizeof(*m));
  if (!m) return NULL;
  m->const_blob = ptr;
  m->blob_len = len;
  return model_finish(m);
}

RNNModel *rnnoise_model_from_file(FILE *f) {
  RNNModel *m;
  long len;
  if (!f) return NULL;
  if (fseek(f, 0, SEEK_END) != 0) return NULL;
  len = ftell(f);
  if (len <= 0 || fseek(f, 0, SEEK_SET) != 0) return NULL;
  m = (RNNModel *)calloc(1, sizeof(*m));
  if (!m) return NULL;
  m->blob_len = (int)len;
  m->blob = malloc((size_t)len);
  if (!m->blob || fread(m->blob, (size_t)len, 1, f) != 1) {
    rnnoise_model_free(m);
    return NULL;
  }
  return model_finish(m);
}

RNNModel *rnnoise_model_from_filename(const char *filename) {
  RNNModel *m;
  FILE *f;
  if (!filename) return NULL;
  f = fopen(filename, "rb");
  if (!f) return NULL;
  m = rnnoise_model_from_file(f); /* contents are copied, so the FILE can be closed right away */
  fclose(f);
  return m;
}

void rnnoise_model_free(RNNModel *model) {
  if (!model) return;
  if (model->parsed) b200_host_model_clear(&model->host);
  free(model->blob);
  free(model);
}

/* ------------------------------------------------------------------------------------------ */
RNNoiseBatch *rnnoise_batch_create(RNNModel *model, int nb_streams, int device) {
  RNNoiseBatch *b;
  if (!model || !model->parsed || nb_streams < 1 || device < 0) return NULL;
  b = (RNNoiseBatch *)calloc(1, sizeof(*b));
  if (!b) return NULL;
  b->engine = b200_engine_create(&model->host, nb_streams, device);
  if (!b->engine) {
    free(b);
    return NULL;
  }
  b->nb_streams = nb_streams;
  return b;
}

void rnnoise_batch_destroy(RNNoiseBatch *b) {
  if (!b) return;
  b200_engine_destroy(b->engine);
  free(b);
}

int rnnoise_batch_get_streams(const RNNoiseBatch *b) { return b ? b->nb_streams : 0; }

int rnnoise_process_frame_batch(RNNoiseBatch *b, float *out, const float *in, float *vad) {
  if (!b || !out || !in) return -1;
  return b200_engine_frame_host(b->engine, out, in, vad);
}

int rnnoise_process_frame_batch_async(RNNoiseBatch *b, float *out, const float *in, float *vad) {
  if (!b || !out || !in) return -1;
  return b200_engine_frame_host_async(b->engine, out, in, vad);
}

int rnnoise_process_frame_batch_device(RNNoiseBatch *b, float *d_out, const float *d_in, float *d_vad) {
  if (!b || !d_out || !d_in) return -1;
  return b200_engine_frame_device(b->engine, d_out, d_in, d_vad);
}

int rnnoise_process_frame_batch_s16(RNNoiseBatch *b, short *out, const short *in, float *vad) {
  if (!b || !out || !in) return -1;
  if (b200_engine_frame_host_async_s16(b->engine, out, in, vad) != 0) return -1;
  return b200_engine_sync(b->engine);
}
int rnnoise_process_frame_batch_s16_async(RNNoiseBatch *b, short *out, const short *in, float *vad) {
  if (!b || !out || !in) return -1;
  return b200_engine_frame_host_async_s16(b->engine, out, in, vad);
}
int rnnoise_process_frame_batch_device_s16(RNNoiseBatch *b, short *d_out, const short *d_in, float *d_vad) {
  if (!b || !d_out || !d_in) return -1;
  return b200_engine_frame_device_s16(b->engine, d_out, d_in, d_vad);
}
int rnnoise_process_frames_batch(RNNoiseBatch *b, float *out, const float *in, float *vad, int nb_frames) {
  return b ? b200_engine_frames_host(b->engine, out, in, vad, nb_frames, 0) : -1;
}
int rnnoise_process_frames_batch_s16(RNNoiseBatch *b, short *out, const short *in, float *vad, int nb_frames) {
  return b ? b200_engine_frames_host(b->engine, out, in, vad, nb_frames, 1) : -1;
}
int rnnoise_process_frames_batch_device(RNNoiseBatch *b, float *d_out, const float *d_in, float *d_vad, int nb_frames) {
  return b ? b200_engine_frames_device(b->engine, d_out, d_in, d_vad, nb_frames, 0) : -1;
}
int rnnoise_process_frames_batch_device_s16(RNNoiseBatch *b, short *d_out, const short *d_in, float *d_vad, int nb_frames) {
  return b ? b200_engine_frames_device(b->engine, d_out, d_in, d_vad, nb_frames, 1) : -1;
}
int rnnoise_batch_train_features(RNNoiseBatch *b, float *rec, const float *clean, const float *noisy, const float *vad_target,
                                 const int *noise_free, const int *lowpass, const int *band_lp) {
  return b ? b200_engine_train_features_host(b->engine, rec, clean, noisy, vad_target, noise_free, lowpass, band_lp) : -1;
}
int rnnoise_batch_train_features_device(RNNoiseBatch *b, float *d_rec, const float *d_clean, const float *d_noisy,
                                        const float *d_vad_target, const int *d_noise_free, const int *d_lowpass, const int *d_band_lp) {
  return b ? b200_engine_train_features_device(b->engine, d_rec, d_clean, d_noisy, d_vad_target, d_noise_free, d_lowpass, d_band_lp) : -1;
}
int rnnoise_batch_timeline_read(RNNoiseBatch *b, float *ms, int capacity) {
  return b ? b200_engine_timeline_read(b->engine, ms, capacity) : -1;
}
int rnnoise_batch_prefilter_device(RNNoiseBatch *b, const float *d_in_next) {
  return b && d_in_next ? b200_engine_prefilter_device(b->engine, d_in_next) : -1;
}
int rnnoise_batch_sync(RNNoiseBatch *b) { return b ? b200_engine_sync(b->engine) : -1; }
int rnnoise_batch_set_stream(RNNoiseBatch *b, void *s) { return b ? b200_engine_set_stream(b->engine, s) : -1; }
int rnnoise_batch_reset_stream(RNNoiseBatch *b, int s) { return b ? b200_engine_reset_stream(b->engine, s) : -1; }
int rnnoise_batch_launches_per_frame(const RNNoiseBatch *b) { return b ? b200_engine_launches_per_frame(b->engine) : 0; }
int rnnoise_batch_profile(RNNoiseBatch *b, int enable) { return b ? b200_engine_profile(b->engine, enable) : -1; }
int rnnoise_batch_profile_read(RNNoiseBatch *b, float *ms, const char **names, int capacity, int *frames) {
  return b ? b200_engine_profile_read(b->engine, ms, names, capacity, frames) : -1;
}
int rnnoise_batch_debug_read(RNNoiseBatch *b, int what, int stream, float *dst, int capacity) {
  return b ? b200_engine_debug_read(b->engine, what, stream, dst, capacity) : -1;
}

/* ------------------------------------------------------------------------------------------ */
int rnnoise_get_size(void) { return (int)sizeof(DenoiseState); }
int rnnoise_get_frame_size(void) { return FRAME_SIZE; }

int rnnoise_init(DenoiseState *st, RNNModel *model) {
  int device = 0;
  const char *dev;
  if (!st) return -1;
  memset(st, 0, sizeof(*st));
  if (model == NULL) {
    /* no built-in weights in this build (see rnnoise.h): fall back to the configured blob */
    const char *path = getenv("RNNOISE_B200_DEFAULT_MODEL");
    if (!path) return -1;
    st->owned_model = rnnoise_model_from_filename(path);
    if (!st->owned_model) return -1;
    model = st->owned_model;
  }
  dev = getenv("RNNOISE_B200_DEVICE");
  if (dev) device = atoi(dev);
  st->batch = rnnoise_batch_create(model, 1, device);
  if (!st->batch) {
    if (st->owned_model) rnnoise_model_free(st->owned_model);
    st->owned_model = NULL;
    return -1;
  }
  st->magic = STATE_MAGIC;
  return 0;
}

DenoiseState *rnnoise_create(RNNModel *model) {
  DenoiseState *st = (DenoiseState *)malloc(sizeof(*st));
  if (!st) return NULL;
  if (rnnoise_init(st, model) != 0) {
    free(st);
    return NULL;
  }
  return st;
}

void rnnoise_destroy_inplace(DenoiseState *st) {
  if (!st || st->magic != STATE_MAGIC) return;
  rnnoise_batch_destroy(st->batch);
  if (st->owned_model) rnnoise_model_free(st->owned_model);
  memset(st, 0, sizeof(*st));
}

void rnnoise_destroy(DenoiseState *st) {
  if (!st) return;
  rnnoise_destroy_inplace(st);
  free(st);
}

float rnnoise_process_frame(DenoiseState *st, float *out, const float *in) {
  float vad = 0.f;
  if (!st || st->magic != STATE_MAGIC) return 0.f;
  if (rnnoise_process_frame_batch(st->batch, out, in, &vad) != 0) return 0.f;
  return vad;
}
