/* engine.h -- internal C interface between the host API (rnnoise_api.c, plain C) and the CUDA
 * engine (engine.cu).  Plain pointers and ints only. */
#ifndef RNNOISE_B200_ENGINE_H
#define RNNOISE_B200_ENGINE_H

#include "model_blob.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct B200Engine B200Engine;

/* Creates device state for nb_streams streams on `device`; uploads the model. NULL on failure. */
B200Engine *b200_engine_create(const B200HostModel *m, int nb_streams, int device);
/* same for one lane of a batch that keeps device_streams streams on this device in total */
B200Engine *b200_engine_create_on(const B200HostModel *m, int nb_streams, int device, int device_streams);
void b200_engine_destroy(B200Engine *e);
int b200_engine_streams(const B200Engine *e);
/* sub-grids ("lanes") the DSP stages of a frame run as */
int b200_engine_ranges(const B200Engine *e);
/* One frame for every stream, device pointers, asynchronous on the engine's stream. */
int b200_engine_frame_device(B200Engine *e, float *d_out, const float *d_in, float *d_vad);
/* One frame, host pointers (copies in, runs, copies out, synchronises). */
int b200_engine_frame_host(B200Engine *e, float *out, const float *in, float *vad);
int b200_engine_frame_host_async(B200Engine *e, float *out, const float *in, float *vad);
int b200_engine_frame_device_s16(B200Engine *e, short *d_out, const short *d_in, float *d_vad);
int b200_engine_frame_host_async_s16(B200Engine *e, short *out, const short *in, float *vad);
/* T frames per stream in one call; buffers are [nb_streams][T * 480] (vad [nb_streams][T]). */
int b200_engine_frames_device(B200Engine *e, void *d_out, const void *d_in, float *d_vad, int nb_frames, int s16);
int b200_engine_frames_host(B200Engine *e, void *out, const void *in, float *vad, int nb_frames, int s16);
/* same without the final synchronisation; pitch_frames = frames per stream row of the host buffers */
int b200_engine_frames_host_enqueue(B200Engine *e, void *out, const void *in, float *vad, int nb_frames, int s16, int pitch_frames);
int b200_engine_set_parent(B200Engine *e, void *parent_stream);
/* Training-feature records [nb_streams][98] (dump_features.c:466-491); per-stream arrays may be NULL. */
int b200_engine_train_features_device(B200Engine *e, float *d_rec, const float *d_clean, const float *d_noisy,
                                      const float *d_vad_target, const int *d_noise_free, const int *d_lowpass, const int *d_band_lp);
int b200_engine_train_features_host(B200Engine *e, float *rec, const float *clean, const float *noisy,
                                    const float *vad_target, const int *noise_free, const int *lowpass, const int *band_lp);
int b200_engine_prefilter_device(B200Engine *e, const float *d_in);
/* frames whose high-pass prefilter has been issued ahead of processing (0, 1 or 2) */
int b200_engine_prefilter_ahead(const B200Engine *e);
/* test hook: start a fresh engine at an arbitrary frame index (counter-wrap tests) */
int b200_engine_debug_set_frames(B200Engine *e, long long frames);
int b200_engine_sync(B200Engine *e);
int b200_engine_set_stream(B200Engine *e, void *cuda_stream);
int b200_engine_reset_stream(B200Engine *e, int stream);
int b200_engine_launches_per_frame(const B200Engine *e);
int b200_engine_profile(B200Engine *e, int enable);
int b200_engine_profile_read(B200Engine *e, float *ms, const char **names, int capacity, int *frames);
int b200_engine_timeline_read(B200Engine *e, float *dst, int capacity);
int b200_engine_debug_read(B200Engine *e, int what, int stream, float *dst, int capacity);
int b200_engine_debug_read_all(B200Engine *e, int what, float *dst, int capacity);

#ifdef __cplusplus
}
#endif
#endif
