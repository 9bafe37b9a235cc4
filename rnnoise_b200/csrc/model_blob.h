/* model_blob.h -- host-side model: parsed "DNNw" weight blob, dense-expanded for the GPU upload.
 * Blob format: reference src/nnet.h:43-62 (WeightHead), src/write_weights.c:46-69; validation rules
 * follow src/parse_lpcnet_weights.c:37-52 (records), :98-121 (sparse index), :123-176 (sizes). */
#ifndef RNNOISE_B200_MODEL_BLOB_H
#define RNNOISE_B200_MODEL_BLOB_H

#ifdef __cplusplus
extern "C" {
#endif

#define B200_NB_FEATURES 65
#define B200_NB_BANDS 32

typedef struct {
  int nb_in, nb_out;
  signed char *w8;          /* owned: dense [out][in] (block-sparse layers expanded with zeros) */
  const float *wf;          /* borrowed from the blob: [in][out] */
  const float *bias, *subias, *scale, *diag; /* borrowed */
} B200Layer;

typedef struct {
  int cond, gru;            /* conv1 outputs, GRU width (inferred from array sizes) */
  B200Layer conv1, conv2, gru_in[3], gru_rec[3], dense_out, vad_dense;
} B200HostModel;

/* Parses `blob`; returns 0 and fills *m (free with b200_host_model_clear), or -1 when any array is
 * missing / mis-sized / the sparse index is inconsistent. The float arrays alias the blob. */
int b200_host_model_parse(B200HostModel *m, const void *blob, int len);
void b200_host_model_clear(B200HostModel *m);

#ifdef __cplusplus
}
#endif
#endif
