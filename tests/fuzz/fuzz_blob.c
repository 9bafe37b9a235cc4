/* tests/fuzz/fuzz_blob.c -- mutation fuzzer for the product blob parser (rnnoise_b200/csrc/model_blob.c), built with
 * AddressSanitizer + UBSan by tests/test_blob_fuzz.py: header bit flips, truncations, corrupted sparse indices and
 * size fields must be rejected or parsed without any out-of-bounds or misaligned access.  Test infrastructure. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "model_blob.h"   /* rnnoise_b200/csrc */
static unsigned long long s = 88172645463325252ULL;
static unsigned rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (unsigned)(s >> 11); }
int main(int argc, char **argv) {
  FILE *f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  unsigned char *blob = malloc(n); if (fread(blob, 1, n, f) != (size_t)n) return 2; fclose(f);
  long offs[128]; int nh = 0; long o = 0;
  while (o < n) { int bs; memcpy(&bs, blob + o + 16, 4); offs[nh++] = o; o += 64 + bs; }
  int ok = 0, bad = 0, iters = atoi(argv[2]);
  for (int it = 0; it < iters; it++) {
    long len = n; int kind = it % 5;
    unsigned char *b = malloc(n); memcpy(b, blob, n);   /* exact-size heap copy: ASAN sees any overrun */
    long h = offs[rnd() % nh];
    if (kind == 0) b[h + rnd() % 64] ^= 1u << (rnd() % 8);
    else if (kind == 1) len = 1 + rnd() % n;
    else if (kind == 2) { int size; memcpy(&size, b + h + 12, 4); if (strstr((char *)b + h + 20, "idx") && size >= 4) { int v = (int)(rnd() % 200000) - 1000; memcpy(b + h + 64 + 4 * (rnd() % (size / 4)), &v, 4); } else b[h + 64] ^= 0xff; }
    else if (kind == 3) { int v = (int)(rnd() % (1u << 28)) - 5; memcpy(b + h + 12 + 4 * (rnd() % 2), &v, 4); }
    else { int v = (int)rnd(); memcpy(b + h + 8 + 4 * (rnd() % 4), &v, 4); }
    unsigned char *c = malloc(len); memcpy(c, b, len); free(b);
    B200HostModel m;
    if (b200_host_model_parse(&m, c, (int)len) == 0) { ok++; b200_host_model_clear(&m); } else bad++;
    free(c);
  }
  free(blob);
  printf("accepted %d rejected %d\n", ok, bad);
  return 0;
}
