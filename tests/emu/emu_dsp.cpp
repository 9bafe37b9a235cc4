// tests/emu/emu_dsp.cpp -- HOST EMULATION of the per-stream DSP device code, for CPU-side tests only.
//
// Compiles rnnoise_b200/csrc/dsp_stream.cuh with a plain C++ compiler: every PHASE runs as a loop
// over the 128 thread ids, so indexing, work partitioning and arithmetic of the exact source the
// GPU executes can be checked against the oracle without a GPU (tests/test_dsp_emulation.py).
// This file is test infrastructure: nothing in the library links or calls it.
// build: g++ -O2 -ffp-contract=off -fPIC -shared -I rnnoise_b200/csrc tests/emu/emu_dsp.cpp
#include <stdlib.h>
#include <string.h>

#include "dsp_stream.cuh"
#include "dsp_tables.hpp"

struct EmuState {
  DspTables T;
  float ring[PITCH_BUF_SIZE], synth_mem[FRAME_SIZE], hp[2];
  float spec[2][4 * FREQ_SIZE], band[2][96], lastg[NB_BANDS], pitch_state[2];
  float features[NB_FEATURES], xb[FRAME_SIZE];
  int silence;
  long frames;
  alignas(16) float sm[8192];   // >= max(SM_PITCH_TOTAL, SM_SPEC_TOTAL, SS_TOTAL)
};

extern "C" {
void *emu_create(void) {
  EmuState *e = (EmuState *)calloc(1, sizeof(EmuState));
  b200_fill_dsp_tables(&e->T);
  return e;
}
void emu_destroy(void *p) { free(p); }

// biquad + analysis of one frame; returns the silence flag
int emu_analysis(void *p, const float *in, float *xb, float *features, float *X, float *P, float *bands,
                 float *pitch) {
  EmuState *e = (EmuState *)p;
  float m0 = e->hp[0], m1 = e->hp[1];
  for (int i = 0; i < FRAME_SIZE; i++) e->xb[i] = biquad_step(in[i], m0, m1);
  e->hp[0] = m0; e->hp[1] = m1;
  const long f = e->frames;
  const int par = (int)(f & 1);
  static_assert(SM_PITCH_TOTAL <= 8192 && SM_SPEC_TOTAL <= 8192 && SS_TOTAL <= 8192, "emu scratch");
  PitchArgs pa;
  pa.xb = e->xb;
  pa.ring = e->ring;
  pa.ring_base = (int)(((f + 1) * FRAME_SIZE) % PITCH_BUF_SIZE);
  pa.pitch_state = e->pitch_state;
  pitch_stream(e->sm, pa, &e->T);
  SpectrumArgs a;
  a.ring = e->ring;
  a.ring_base = pa.ring_base;
  a.pitch_state = e->pitch_state;
  a.spec_out = e->spec[par];
  a.band_out = e->band[par];
  a.features = e->features;
  a.silence = &e->silence;
  spectrum_stream(e->sm, a, &e->T);
  memcpy(xb, e->xb, sizeof(e->xb));
  memcpy(features, e->features, sizeof(e->features));
  memcpy(X, e->spec[par], 2 * FREQ_SIZE * sizeof(float));
  memcpy(P, e->spec[par] + 2 * FREQ_SIZE, 2 * FREQ_SIZE * sizeof(float));
  memcpy(bands, e->band[par], 96 * sizeof(float));
  int period; memcpy(&period, &e->pitch_state[0], 4);
  pitch[0] = (float)period; pitch[1] = e->pitch_state[1];
  return e->silence;
}

// synthesis of the same frame given the network gains; advances the frame counter
void emu_synthesis(void *p, const float *gains, float *out, float *lastg) {
  EmuState *e = (EmuState *)p;
  const int par = (int)(e->frames & 1);
  SynthesisArgs a;
  a.spec_delayed = e->spec[par ^ 1];
  a.band_delayed = e->band[par ^ 1];
  a.band_cur = e->band[par];
  a.gains = gains;
  a.silence = &e->silence;
  a.lastg = e->lastg;
  a.synthesis_mem = e->synth_mem;
  a.out = out;
  synthesis_stream(e->sm, a, &e->T);
  memcpy(lastg, e->lastg, sizeof(e->lastg));
  e->frames++;
}
}
