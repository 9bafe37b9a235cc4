/* tests/mock/mock_engine.c -- a stand-in for the CUDA engine behind rnnoise_b200/csrc/engine.h, so that the host
 * API layer (rnnoise_api.c: lanes, pointer offsets, routing of per-stream calls) can be tested without a GPU.
 * Every "engine" tags what it writes with its creation index and the LOCAL stream index, which lets
 * tests/test_api_lanes_mock.py reconstruct which lane handled which slice of the caller's buffers.
 * Test infrastructure only: never part of librnnoise_b200.so. */
#include <stdlib.h>
#include <string.h>

#include "engine.h"

#define FRAME 480
struct B200Engine { int S, id, last_reset, profiling, device, ahead, fail_frames; long long frames; void *stream, *parent; };
static int g_next_id = 0;
void mock_reset_ids(void) { g_next_id = 0; }

B200Engine *b200_engine_create(const B200HostModel *m, int nb_streams, int device) {
  B200Engine *e;
  if (!m || nb_streams < 1 || device < 0 || device >= 8) return NULL;
  e = (B200Engine *)calloc(1, sizeof(*e));
  e->S = nb_streams; e->id = g_next_id++; e->last_reset = -1; e->device = device;
  return e;
}
B200Engine *b200_engine_create_on(const B200HostModel *m, int nb_streams, int device, int device_streams) {
  return device_streams >= nb_streams ? b200_engine_create(m, nb_streams, device) : NULL;
}
void b200_engine_destroy(B200Engine *e) { free(e); }
int b200_engine_streams(const B200Engine *e) { return e->S; }
int b200_engine_ranges(const B200Engine *e) { return e->S >= 1024 && e->S < 32768 ? 2 : 1; }

/* out = 2 * in + 1000 * engine id + local stream (+ frame index / 4); vad = 100 * id + local stream + frame / 1000 */
static void fill_f(B200Engine *e, float *out, const float *in, float *vad, int T) {
  for (int s = 0; s < e->S; s++)
    for (int t = 0; t < T; t++) {
      for (int i = 0; i < FRAME; i++) {
        size_t k = ((size_t)s * T + t) * FRAME + i;
        out[k] = 2 * in[k] + 1000 * e->id + s + 0.25f * t;
      }
      if (vad) vad[(size_t)s * T + t] = 100 * e->id + s + t / 1000.f;
    }
}
static void fill_s(B200Engine *e, short *out, const short *in, float *vad, int T) {
  for (int s = 0; s < e->S; s++)
    for (int t = 0; t < T; t++) {
      for (int i = 0; i < FRAME; i++) {
        size_t k = ((size_t)s * T + t) * FRAME + i;
        out[k] = (short)(in[k] + 1000 * e->id + s + t);
      }
      if (vad) vad[(size_t)s * T + t] = 100 * e->id + s + t / 1000.f;
    }
}
/* test hooks: engine `id` fails its next per-frame call; frames handled / device of an engine */
static int g_fail_id = -1;
void mock_fail_next(int id) { g_fail_id = id; }
static int failing(B200Engine *e) { if (e->id == g_fail_id) { g_fail_id = -1; return 1; } e->frames++; e->ahead = 0; return 0; }
int b200_engine_frame_device(B200Engine *e, float *o, const float *i, float *v) { if (failing(e)) return -1; fill_f(e, o, i, v, 1); return 0; }
int b200_engine_frame_host(B200Engine *e, float *o, const float *i, float *v) { fill_f(e, o, i, v, 1); return 0; }
int b200_engine_frame_host_async(B200Engine *e, float *o, const float *i, float *v) { if (failing(e)) return -1; fill_f(e, o, i, v, 1); return 0; }
int b200_engine_frame_device_s16(B200Engine *e, short *o, const short *i, float *v) { fill_s(e, o, i, v, 1); return 0; }
int b200_engine_frame_host_async_s16(B200Engine *e, short *o, const short *i, float *v) { fill_s(e, o, i, v, 1); return 0; }
int b200_engine_frames_device(B200Engine *e, void *o, const void *i, float *v, int T, int s16) {
  if (s16) fill_s(e, (short *)o, (const short *)i, v, T); else fill_f(e, (float *)o, (const float *)i, v, T);
  return 0;
}
int b200_engine_frames_host(B200Engine *e, void *o, const void *i, float *v, int T, int s16) { return b200_engine_frames_device(e, o, i, v, T, s16); }
int b200_engine_frames_host_enqueue(B200Engine *e, void *o, const void *i, float *v, int T, int s16, int pitch_frames) {
  return pitch_frames == T ? b200_engine_frames_device(e, o, i, v, T, s16) : -1;
}
int b200_engine_set_parent(B200Engine *e, void *p) { e->parent = p; e->stream = NULL; return 0; }
int b200_engine_set_stream(B200Engine *e, void *s) { e->stream = s; e->parent = NULL; return 0; }
/* rec[k] = clean[k] + noisy[k] for k < 96; rec[96] = 1000 * id + local stream; rec[97] = vad + 2 * noise_free + 4 * lowpass + 4096 * band_lp */
int b200_engine_train_features_device(B200Engine *e, float *rec, const float *clean, const float *noisy, const float *vt,
                                      const int *nf, const int *lp, const int *bl) {
  for (int s = 0; s < e->S; s++) {
    for (int k = 0; k < 96; k++) rec[(size_t)s * 98 + k] = clean[(size_t)s * FRAME + k] + noisy[(size_t)s * FRAME + k];
    rec[(size_t)s * 98 + 96] = 1000 * e->id + s;
    rec[(size_t)s * 98 + 97] = (vt ? vt[s] : 0) + 2 * (nf ? nf[s] : 0) + 4 * (lp ? lp[s] : 481) + 4096 * (bl ? bl[s] : 32);
  }
  return 0;
}
int b200_engine_train_features_host(B200Engine *e, float *rec, const float *c, const float *n, const float *vt, const int *nf,
                                    const int *lp, const int *bl) { return b200_engine_train_features_device(e, rec, c, n, vt, nf, lp, bl); }
/* the "hint" is recorded by writing nothing; it only has to receive the lane's slice: remember its first sample */
static float g_hint[16];
int b200_engine_prefilter_device(B200Engine *e, const float *d_in) { if (e->id < 16) g_hint[e->id] = d_in[0]; e->ahead++; return 0; }
int b200_engine_prefilter_ahead(const B200Engine *e) { return e->ahead; }
int b200_engine_debug_set_frames(B200Engine *e, long long f) { if (e->frames) return -1; e->frames = f; return 0; }
float mock_hint(int id) { return g_hint[id]; }
int b200_engine_sync(B200Engine *e) { (void)e; return 0; }
int b200_engine_reset_stream(B200Engine *e, int s) { if (s < 0 || s >= e->S) return -1; e->last_reset = s; return 0; }
int b200_engine_launches_per_frame(const B200Engine *e) { (void)e; return 10; }
int b200_engine_profile(B200Engine *e, int enable) { e->profiling = enable; return 0; }
int b200_engine_profile_read(B200Engine *e, float *ms, const char **names, int capacity, int *frames) {
  static const char *nm[2] = {"k_a", "k_b"};
  if (capacity < 2) return -1;
  ms[0] = 1.f + e->id; ms[1] = 10.f;
  if (names) { names[0] = nm[0]; names[1] = nm[1]; }
  if (frames) *frames = 7;
  return 2;
}
int b200_engine_timeline_read(B200Engine *e, float *dst, int capacity) { (void)capacity; dst[0] = (float)e->id; return 1; }
/* dst = {engine id, local stream, last reset stream of this engine, parent set?, stream set?} */
int b200_engine_debug_read(B200Engine *e, int what, int s, float *dst, int cap) {
  (void)what;
  if (s < 0 || s >= e->S || cap < 5) return -1;
  dst[0] = (float)e->id; dst[1] = (float)s; dst[2] = (float)e->last_reset; dst[3] = e->parent ? 1.f : 0.f; dst[4] = e->stream ? 1.f : 0.f;
  if (cap >= 7) { dst[5] = (float)e->device; dst[6] = (float)e->frames; }
  return 5;
}
int b200_engine_debug_read_all(B200Engine *e, int what, float *dst, int cap) {
  (void)what;
  if (cap < e->S * 2) return -1;
  for (int s = 0; s < e->S; s++) { dst[2 * s] = (float)e->id; dst[2 * s + 1] = (float)s; }
  return 2;
}
