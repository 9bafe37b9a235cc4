/* model_blob.c -- see model_blob.h. Plain C, no CUDA. */
#include "model_blob.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define BLOCK 64
#define MAX_ARRAYS 256

typedef struct { char head[4]; int version, type, size, block_size; char name[44]; } Head;
typedef struct { const char *name; int size; const void *data; } Arr;

static const Arr *lookup(const Arr *a, int n, const char *layer, const char *suffix) {
  char nm[96];
  snprintf(nm, sizeof(nm), "%s%s", layer, suffix);
  for (int i = 0; i < n; i++) if (!strcmp(a[i].name, nm)) return &a[i];
  return NULL;
}
static const void *sized(const Arr *a, int n, const char *layer, const char *suffix, int bytes) {
  const Arr *e = lookup(a, n, layer, suffix);
  return (e && e->size == bytes) ? e->data : NULL;
}

static int float_layer(B200Layer *l, const Arr *a, int n, const char *name, int nb_in, int nb_out) {
  memset(l, 0, sizeof(*l));
  l->nb_in = nb_in; l->nb_out = nb_out;
  l->bias = sized(a, n, name, "_bias", nb_out * 4);
  l->wf = sized(a, n, name, "_weights_float", nb_in * nb_out * 4);
  return (l->bias && l->wf) ? 0 : -1;
}

/* int8 layer: dense [out/8][in/4][8][4] tiles, or block-sparse: per group of 8 outputs
 * idx = {nblocks, pos...} and one 32-byte tile w[4*o+i] per kept block. */
static int int8_layer(B200Layer *l, const Arr *a, int n, const char *name, int nb_in, int nb_out,
                      int sparse, int has_diag) {
  memset(l, 0, sizeof(*l));
  l->nb_in = nb_in; l->nb_out = nb_out;
  l->bias = sized(a, n, name, "_bias", nb_out * 4);
  l->subias = sized(a, n, name, "_subias", nb_out * 4);
  l->scale = sized(a, n, name, "_scale", nb_out * 4);
  if (!l->bias || !l->subias || !l->scale) return -1;
  if (has_diag && !(l->diag = sized(a, n, name, "_weights_diag", nb_out * 4))) return -1;
  const Arr *w = lookup(a, n, name, "_weights_int8");
  if (!w || (nb_in & 3) || (nb_out & 7)) return -1;
  l->w8 = (signed char *)calloc((size_t)nb_in * nb_out, 1);
  if (!l->w8) return -1;
  const signed char *src = (const signed char *)w->data;
  if (!sparse) {
    if (w->size != nb_in * nb_out) return -1;
    for (int ob = 0; ob < nb_out / 8; ob++)
      for (int ib = 0; ib < nb_in / 4; ib++, src += 32)
        for (int o = 0; o < 8; o++)
          for (int i = 0; i < 4; i++) l->w8[(size_t)(ob * 8 + o) * nb_in + ib * 4 + i] = src[4 * o + i];
    return 0;
  }
  const Arr *ix = lookup(a, n, name, "_weights_idx");
  if (!ix) return -1;
  const int *idx = (const int *)ix->data;
  int remain = ix->size / 4, tiles = 0, rows_left = nb_out;
  for (int ob = 0; remain > 0; ob++) {
    int nb = *idx++;
    /* nb >= remain (not remain < nb + 1: nb + 1 overflows for nb == INT_MAX) */
    if (nb < 0 || nb >= remain || rows_left <= 0) return -1;
    int last_pos = -4;
    for (int b = 0; b < nb; b++) {
      int pos = *idx++;
      if (pos < 0 || pos > nb_in - 4 || (pos & 3)) return -1;
      /* positions inside one output block must increase: a repeated position would ACCUMULATE in the
         reference's sparse kernel (vec_avx.h:778-828) but overwrite in this dense expansion; the exporter
         never emits one (wexchange/c_export/common.py), so it is rejected instead of mis-read */
      if (pos <= last_pos) return -1;
      last_pos = pos;
      if ((tiles + 1) * 32 > w->size) return -1;
      const signed char *t = src + (size_t)tiles * 32;
      for (int o = 0; o < 8; o++)
        for (int i = 0; i < 4; i++) l->w8[(size_t)(ob * 8 + o) * nb_in + pos + i] = t[4 * o + i];
      tiles++;
    }
    rows_left -= 8;
    remain -= nb + 1;
  }
  if (rows_left != 0 || tiles * 32 != w->size) return -1;
  return 0;
}

int b200_host_model_parse(B200HostModel *m, const void *blob, int len) {
  Arr arr[MAX_ARRAYS];
  int n = 0;
  const unsigned char *p = (const unsigned char *)blob;
  memset(m, 0, sizeof(*m));
  if (!blob) return -1;
  while (len > 0) {
    Head h;   /* copied out: a corrupted block_size must not turn into a misaligned struct access */
    if (len < BLOCK) return -1;
    memcpy(&h, p, sizeof(h));
    if (h.block_size < h.size || h.block_size > len - BLOCK) return -1;
    /* the writer pads records to 64 bytes (write_weights.c:57); anything not a multiple of 4 would leave the
       following float / int arrays misaligned, which no valid blob does */
    if (h.block_size & 3) return -1;
    if (h.name[sizeof(h.name) - 1] != 0 || h.size <= 0 || n >= MAX_ARRAYS) return -1;
    arr[n].name = ((const Head *)p)->name; arr[n].size = h.size; arr[n].data = p + BLOCK; n++;
    p += BLOCK + h.block_size;
    len -= BLOCK + h.block_size;
  }
  const Arr *c1 = lookup(arr, n, "conv1", "_bias"), *g1 = lookup(arr, n, "gru1_recurrent", "_bias");
  if (!c1 || !g1 || (c1->size & 3) || g1->size % 12) return -1;
  int cond = c1->size / 4, gru = g1->size / 12;
  if (cond <= 0 || gru <= 0 || (cond & 3) || (gru & 7)) return -1;
  m->cond = cond; m->gru = gru;
  int err = 0;
  err |= float_layer(&m->conv1, arr, n, "conv1", 3 * B200_NB_FEATURES, cond);
  err |= int8_layer(&m->conv2, arr, n, "conv2", 3 * cond, gru, 0, 0);
  for (int k = 0; k < 3; k++) {
    char nm[32];
    snprintf(nm, sizeof(nm), "gru%d_input", k + 1);
    err |= int8_layer(&m->gru_in[k], arr, n, nm, gru, 3 * gru, 1, 0);
    snprintf(nm, sizeof(nm), "gru%d_recurrent", k + 1);
    err |= int8_layer(&m->gru_rec[k], arr, n, nm, gru, 3 * gru, 1, 1);
  }
  err |= float_layer(&m->dense_out, arr, n, "dense_out", 4 * gru, B200_NB_BANDS);
  err |= float_layer(&m->vad_dense, arr, n, "vad_dense", 4 * gru, 1);
  if (err) { b200_host_model_clear(m); return -1; }
  return 0;
}

void b200_host_model_clear(B200HostModel *m) {
  free(m->conv2.w8);
  for (int k = 0; k < 3; k++) { free(m->gru_in[k].w8); free(m->gru_rec[k].w8); }
  memset(m, 0, sizeof(*m));
}
