/* examples/rnnoise_batch_demo.c -- denoise several raw 16-bit 48 kHz mono files at once on one GPU.
 *
 * The batched counterpart of the reference's examples/rnnoise_demo.c (which loops rnnoise_process_frame over
 * one file, :53-64): every input file is one stream of a batch, and the whole signals go through one
 * multi-frame call (rnnoise_process_frames_batch_s16: int16 PCM in and out, the formats the reference demo
 * reads and writes).  Shorter files are zero-padded to the longest.
 *
 *   gcc -I include examples/rnnoise_batch_demo.c -o rnnoise_batch_demo \
 *       rnnoise_b200/librnnoise_b200.so -Wl,-rpath,'$ORIGIN/rnnoise_b200'
 *   ./rnnoise_batch_demo weights_blob.bin a.raw b.raw ...     -> a.raw.denoised, b.raw.denoised, ...
 *
 * Like the reference demo, the first output frame of each file is the (silent) look-ahead frame.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rnnoise.h"

#define FRAME 480

static long file_samples(FILE *f) {
  long n;
  fseek(f, 0, SEEK_END);
  n = ftell(f) / (long)sizeof(short);
  fseek(f, 0, SEEK_SET);
  return n;
}

int main(int argc, char **argv) {
  int nb, s, rc = 1;
  long frames = 0, f;
  short *in = NULL, *out = NULL;
  float *vad = NULL;
  long *len = NULL;
  RNNModel *model = NULL;
  RNNoiseBatch *batch = NULL;
  if (argc < 3) {
    fprintf(stderr, "usage: %s <weights_blob.bin> <in1.raw> [in2.raw ...]\n", argv[0]);
    return 1;
  }
  nb = argc - 2;
  len = (long *)calloc((size_t)nb, sizeof(*len));
  for (s = 0; s < nb; s++) {
    FILE *fi = fopen(argv[2 + s], "rb");
    if (!fi) { fprintf(stderr, "cannot open %s\n", argv[2 + s]); goto done; }
    len[s] = file_samples(fi);
    fclose(fi);
    if ((len[s] + FRAME - 1) / FRAME > frames) frames = (len[s] + FRAME - 1) / FRAME;
  }
  if (frames == 0) { fprintf(stderr, "empty input\n"); goto done; }
  in = (short *)calloc((size_t)nb * frames * FRAME, sizeof(short));   /* [stream][frames * 480], zero padded */
  out = (short *)malloc((size_t)nb * frames * FRAME * sizeof(short));
  vad = (float *)malloc((size_t)nb * frames * sizeof(float));
  if (!in || !out || !vad) goto done;
  for (s = 0; s < nb; s++) {
    FILE *fi = fopen(argv[2 + s], "rb");
    if (!fi || fread(in + (size_t)s * frames * FRAME, sizeof(short), (size_t)len[s], fi) != (size_t)len[s]) {
      fprintf(stderr, "cannot read %s\n", argv[2 + s]);
      if (fi) fclose(fi);
      goto done;
    }
    fclose(fi);
  }
  model = rnnoise_model_from_filename(argv[1]);
  if (!model) { fprintf(stderr, "cannot load model %s\n", argv[1]); goto done; }
  batch = rnnoise_batch_create(model, nb, 0);
  if (!batch) { fprintf(stderr, "cannot create the batch (no usable CUDA device? this library has no CPU path)\n"); goto done; }
  if (rnnoise_process_frames_batch_s16(batch, out, in, vad, (int)frames) != 0) { fprintf(stderr, "processing failed\n"); goto done; }
  for (s = 0; s < nb; s++) {
    char name[4096];
    FILE *fo;
    double speech = 0;
    snprintf(name, sizeof(name), "%s.denoised", argv[2 + s]);
    fo = fopen(name, "wb");
    if (!fo) { fprintf(stderr, "cannot write %s\n", name); goto done; }
    fwrite(out + (size_t)s * frames * FRAME, sizeof(short), (size_t)frames * FRAME, fo);
    fclose(fo);
    for (f = 0; f < frames; f++) speech += vad[(size_t)s * frames + f] > 0.5f;
    printf("%s: %ld frames, voice activity in %.1f %% of them -> %s\n", argv[2 + s], frames, 100.0 * speech / frames, name);
  }
  rc = 0;
done:
  if (batch) rnnoise_batch_destroy(batch);
  if (model) rnnoise_model_free(model);
  free(in); free(out); free(vad); free(len);
  return rc;
}
