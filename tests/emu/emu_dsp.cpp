// tests/emu/emu_dsp.cpp -- HOST EMULATION of the per-stream DSP device code, for CPU-side tests only.
//
// Compiles rnnoise_b200/csrc/dsp_stream.cuh with a plain C++ compiler: every PHASE runs as a loop
// over the 128 thread ids, so indexing, work partitioning and arithmetic of the exact source the
// GPU executes can be checked against the oracle without a GPU (tests/test_dsp_emulation.py).
// This file is test infrastructure: nothing in the library links or calls it.
// build: g++ -O2 -ffp-contract=off -fPIC -shared -DPITCH_NS=4 -I rnnoise_b200/csrc tests/emu/emu_dsp.cpp
#include <stdlib.h>
#include <string.h>

#include "dsp_stream.cuh"
#include "dsp_pitch.cuh"
#include "dsp_tables.hpp"

struct EmuStream {
  float ring[PITCH_BUF_SIZE], synth_mem[FRAME_SIZE], hp[2];
  float spec[2][4 * FREQ_SIZE], band[2][96], lastg[NB_BANDS], pitch_state[2];
  float features[NB_FEATURES], xb[FRAME_SIZE];
  float clean_mem[FRAME_SIZE], rec[TRAIN_RECORD];
  int silence;
};
struct EmuState {
  DspTables T;
  EmuStream st[PITCH_NS];
  long frames;
  alignas(16) float sm[PITCH_NS * SM_PITCH_TOTAL > 8192 ? PITCH_NS * SM_PITCH_TOTAL : 8192];
};

extern "C" {
int emu_streams(void) { return PITCH_NS; }
// the FFT work buffer's address swizzle (dsp_core.cuh: fsw), for the structural test of its bank mapping
int emu_fsw(int idx) { return fsw(idx); }
void *emu_create(void) {
  EmuState *e = (EmuState *)calloc(1, sizeof(EmuState));
  b200_fill_dsp_tables(&e->T);
  return e;
}
void emu_destroy(void *p) { free(p); }

// biquad + pitch + spectrum of one frame for a group of PITCH_NS streams (in: [NS][480]); nstreams < NS
// leaves the trailing streams absent (exercises the partial-group guards)
void emu_analysis(void *p, const float *in, int nstreams) {
  EmuState *e = (EmuState *)p;
  static_assert(SM_SPEC_TOTAL <= 8192 && SS_TOTAL <= 8192, "emu scratch");
  const long f = e->frames;
  const int par = (int)(f & 1);
  PitchArgs pa[PITCH_NS];
  for (int q = 0; q < PITCH_NS; q++) {
    EmuStream &s = e->st[q];
    pa[q].ring = nullptr;
    if (q >= nstreams) continue;
    float m0 = s.hp[0], m1 = s.hp[1];
    for (int i = 0; i < FRAME_SIZE; i++) s.xb[i] = biquad_step(in[q * FRAME_SIZE + i], m0, m1);
    s.hp[0] = m0; s.hp[1] = m1;
    pa[q].xb = s.xb;
    pa[q].ring = s.ring;
    pa[q].ring_base = (int)(((f + 1) * FRAME_SIZE) % PITCH_BUF_SIZE);
    pa[q].pitch_state = s.pitch_state;
  }
  pitch_streams(e->sm, pa, &e->T);
  for (int q = 0; q < nstreams; q++) {
    EmuStream &s = e->st[q];
    SpectrumArgs a;
    a.ring = s.ring;
    a.ring_base = pa[q].ring_base;
    a.pitch_state = s.pitch_state;
    a.spec_out = s.spec[par];
    a.band_out = s.band[par];
    a.features = s.features;
    a.silence = &s.silence;
    a.lowpass = FREQ_SIZE;
    spectrum_stream<false>(e->sm, a, &e->T);
  }
}
// training-feature record of one frame (k_train_features): no prefilter, TRAINING semantics.
// clean/noisy: [NS][480]; per-stream lowpass / band_lp / vad_target / noise_free; rec: [NS][98]; quiet: [NS]
void emu_train(void *p, const float *clean, const float *noisy, int nstreams, const int *lowpass, const int *band_lp,
               const float *vad_target, const int *noise_free, float *rec, int *quiet) {
  EmuState *e = (EmuState *)p;
  const long f = e->frames;
  const int par = (int)(f & 1);
  PitchArgs pa[PITCH_NS];
  for (int q = 0; q < PITCH_NS; q++) {
    EmuStream &s = e->st[q];
    pa[q].ring = nullptr;
    if (q >= nstreams) continue;
    memcpy(s.xb, noisy + q * FRAME_SIZE, sizeof(s.xb));
    pa[q].xb = s.xb;
    pa[q].ring = s.ring;
    pa[q].ring_base = (int)(((f + 1) * FRAME_SIZE) % PITCH_BUF_SIZE);
    pa[q].pitch_state = s.pitch_state;
  }
  pitch_streams(e->sm, pa, &e->T);
  for (int q = 0; q < nstreams; q++) {
    EmuStream &s = e->st[q];
    SpectrumArgs a;
    a.ring = s.ring;
    a.ring_base = pa[q].ring_base;
    a.pitch_state = s.pitch_state;
    a.spec_out = s.spec[par];
    a.band_out = s.band[par];
    a.features = s.rec;
    a.silence = &s.silence;
    a.lowpass = lowpass[q];
    spectrum_stream<true>(e->sm, a, &e->T);
    TrainArgs t;
    t.clean = clean + q * FRAME_SIZE;
    t.clean_mem = s.clean_mem;
    t.rec = s.rec;
    t.quiet = &s.silence;
    t.lowpass = lowpass[q];
    t.band_lp = band_lp[q];
    t.vad_target = vad_target[q];
    t.noise_free = noise_free[q];
    train_targets_stream(e->sm, t, &e->T);
    memcpy(rec + q * TRAIN_RECORD, s.rec, sizeof(s.rec));
    quiet[q] = s.silence;
  }
  e->frames++;
}
// results of the last emu_analysis for stream q; returns the silence flag
int emu_get(void *p, int q, float *xb, float *features, float *X, float *P, float *bands, float *pitch) {
  EmuState *e = (EmuState *)p;
  EmuStream &s = e->st[q];
  const int par = (int)(e->frames & 1);
  memcpy(xb, s.xb, sizeof(s.xb));
  memcpy(features, s.features, sizeof(s.features));
  memcpy(X, s.spec[par], 2 * FREQ_SIZE * sizeof(float));
  memcpy(P, s.spec[par] + 2 * FREQ_SIZE, 2 * FREQ_SIZE * sizeof(float));
  memcpy(bands, s.band[par], 96 * sizeof(float));
  int period; memcpy(&period, &s.pitch_state[0], 4);
  pitch[0] = (float)period; pitch[1] = s.pitch_state[1];
  return s.silence;
}
// synthesis of the same frame for stream q given the network gains
void emu_synthesis(void *p, int q, const float *gains, float *out, float *lastg) {
  EmuState *e = (EmuState *)p;
  EmuStream &s = e->st[q];
  const int par = (int)(e->frames & 1);
  SynthesisArgs a;
  a.spec_delayed = s.spec[par ^ 1];
  a.band_delayed = s.band[par ^ 1];
  a.band_cur = s.band[par];
  a.gains = gains;
  a.silence = &s.silence;
  a.lastg = s.lastg;
  a.synthesis_mem = s.synth_mem;
  a.out = out;
  a.out_s16 = nullptr;
  synthesis_stream(e->sm, a, &e->T);
  memcpy(lastg, s.lastg, sizeof(s.lastg));
}
// ---- the default pitch kernel's body (dsp_pitch.cuh: pitch_group, PG streams per CTA) + spectrum_stream ----
struct EmuGroup {
  DspTables T;
  float xb[PG][FRAME_SIZE], ring[PG][PITCH_BUF_SIZE], pitch_state[PG][2], hp[PG][2];
  float spec[PG][4 * FREQ_SIZE], band[PG][96], features[PG][NB_FEATURES];
  int silence[PG];
  long frames;
  alignas(16) float sm[PG * P2_STRIDE > 8192 ? PG * P2_STRIDE : 8192];
};
int emu_group_streams(void) { return PG; }
void *emu_group_create(void) {
  EmuGroup *e = (EmuGroup *)calloc(1, sizeof(EmuGroup));
  b200_fill_dsp_tables(&e->T);
  return e;
}
void emu_group_destroy(void *p) { free(p); }
// biquad + pitch_group + spectrum of one frame for n <= PG streams (in: [n][480]); outputs per stream:
// features [n][65], pitch [n][2] = {period, gain}, silence [n]
void emu_group_analysis(void *p, const float *in, int n, float *features, float *pitch, int *silence) {
  EmuGroup *e = (EmuGroup *)p;
  const long f = e->frames++;
  for (int q = 0; q < n; q++) {
    float m0 = e->hp[q][0], m1 = e->hp[q][1];
    for (int i = 0; i < FRAME_SIZE; i++) e->xb[q][i] = biquad_step(in[q * FRAME_SIZE + i], m0, m1);
    e->hp[q][0] = m0; e->hp[q][1] = m1;
  }
  PitchGroup g;
  g.xb = &e->xb[0][0]; g.ring = &e->ring[0][0]; g.pitch_state = &e->pitch_state[0][0];
  g.n = n;
  g.ring_base = (int)(((f + 1) * FRAME_SIZE) % PITCH_BUF_SIZE);
  pitch_group(e->sm, g);
  for (int q = 0; q < n; q++) {
    SpectrumArgs a;
    a.ring = e->ring[q]; a.ring_base = g.ring_base; a.pitch_state = e->pitch_state[q];
    a.spec_out = e->spec[q]; a.band_out = e->band[q]; a.features = e->features[q]; a.silence = &e->silence[q];
    a.lowpass = FREQ_SIZE;
    spectrum_stream<false>(e->sm, a, &e->T);
    memcpy(features + q * NB_FEATURES, e->features[q], sizeof(e->features[q]));
    int period; memcpy(&period, &e->pitch_state[q][0], 4);
    pitch[2 * q] = (float)period; pitch[2 * q + 1] = e->pitch_state[q][1];
    silence[q] = e->silence[q];
  }
}
// rd_candidate() exposed for an exhaustive check of its division-free arithmetic
void emu_rd_candidate(int k, int T0, int *T1, int *T1b) { rd_candidate(k, T0, T1, T1b); }
void emu_advance(void *p) { ((EmuState *)p)->frames++; }
}
