/* oracle/rnnoise_port.h -- CPU restatement ("port") of the reference hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing here is linked into, imported by or called from the product
 * (rnnoise_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it, as
 * the checker.  It restates, in plain scalar C and in its own structure, what
 * rnnoise_process_frame() (reference src/denoise.c:457-504) computes, following the arithmetic of
 * the reference's canonical x86 build (`--enable-x86-rtcd` on an AVX2 host: SSE2 scalar float DSP
 * without FMA contraction, AVX2+FMA network kernels) operation for operation -- with ONE stated
 * difference: the activations' `_mm256_rcp_ps` (an implementation-defined 12-bit estimate,
 * src/vec_avx.h:413,442) is replaced by a correctly rounded reciprocal.  This is the arithmetic
 * specification the CUDA path implements bit-for-bit.
 *
 * Parity pin: oracle/_ref (the unmodified reference compiled from /root/reference) -- see
 * tests/test_oracle_port.py: every DSP quantity (X, P, Ex, Ep, Exp, features, pitch, silence) is
 * bit-identical to the reference build; network outputs differ only by the rcp envelope.
 */
#ifndef RNNOISE_PORT_H
#define RNNOISE_PORT_H

#ifdef __cplusplus
extern "C" {
#endif

#define RP_FRAME 480
#define RP_WINDOW 960
#define RP_FREQ 481
#define RP_BANDS 32
#define RP_FEATURES 65
#define RP_PITCH_MIN 60
#define RP_PITCH_MAX 768
#define RP_PITCH_FRAME 960
#define RP_PITCH_BUF 1728
#define RP_MAX_GRU 1024
#define RP_MAX_COND 512

typedef struct { float r, i; } rp_cpx;

/* One dense-expanded layer. int8 layers: w8[out][in] (zeros where the block-sparse blob has no
 * block); float layers: wf[in][out] exactly as stored (reference weights[j*N+i]). */
typedef struct {
  int nb_in, nb_out;
  int is_int8;
  signed char *w8;
  const float *wf;
  const float *bias, *subias, *scale, *diag;
} rp_layer;

typedef struct {
  int cond, gru; /* CONV1_OUT_SIZE, GRU size */
  rp_layer conv1, conv2, gru_in[3], gru_rec[3], dense_out, vad_dense;
  void *blob_copy;
} rp_model;

typedef struct {
  float analysis_mem[RP_FRAME];
  float synthesis_mem[RP_FRAME];
  float pitch_buf[RP_PITCH_BUF];
  float last_gain;
  int last_period;
  float mem_hp_x[2];
  float lastg[RP_BANDS];
  float conv1_state[2 * RP_FEATURES];
  float conv2_state[2 * RP_MAX_COND];
  float gru_state[3][RP_MAX_GRU];
  rp_cpx delayed_X[RP_FREQ], delayed_P[RP_FREQ];
  float delayed_Ex[RP_BANDS], delayed_Ep[RP_BANDS], delayed_Exp[RP_BANDS];
} rp_state;

/* Optional per-frame trace of intermediates (all may be compared against oracle/_ref). */
typedef struct {
  float xb[RP_FRAME];           /* after the high-pass biquad */
  rp_cpx X[RP_FREQ], P[RP_FREQ];
  float Ex[RP_BANDS], Ep[RP_BANDS], Exp[RP_BANDS];
  float features[RP_FEATURES];
  int silence, pitch;
  float pitch_gain;
  float g_raw[RP_BANDS];
  float vad;
} rp_trace;

/* tables */
void rp_tables_init(void);
const float *rp_half_window(void);     /* [480] */
const float *rp_dct_table(void);       /* [32*32] */
const rp_cpx *rp_twiddles(void);       /* [960] */
const int *rp_bitrev(void);            /* [960] */

/* stages (each cites the reference function it restates in rnnoise_port.c) */
void rp_fft960(const rp_cpx *in, rp_cpx *out);
void rp_biquad_hp(float *y, float mem[2], const float *x, int n);
void rp_band_energy(float *E, const rp_cpx *X);
void rp_band_corr(float *E, const rp_cpx *X, const rp_cpx *P);
void rp_interp_band_gain(float *g, const float *band); /* g[481]; bins >= 400 are 0 */
void rp_dct(float *out, const float *in);
void rp_pitch_downsample(const float *buf1728, float *lp864);
int rp_pitch_search(const float *lp864);               /* returns the lag in [0, 587) */
float rp_remove_doubling(const float *lp864, int *T0, int prev_period, float prev_gain);
int rp_frame_features(rp_state *st, rp_cpx *X, rp_cpx *P, float *Ex, float *Ep, float *Exp,
                      float *features, const float *xb);
/* dump_features.c:466-491 frame loop with -DTRAINING=1 semantics; rec[98]; returns the quiet flag */
int rp_train_frame(rp_state *clean_st, rp_state *noisy_st, const float *clean, const float *noisy, float vad_target,
                   int noise_free, int lowpass, int band_lp, float *rec);
void rp_pitch_filter(rp_cpx *X, const rp_cpx *P, const float *Ex, const float *Ep, const float *Exp,
                     const float *g);

/* network */
float rp_tanh(float x);
float rp_sigmoid(float x);
unsigned char rp_quant_u8(float x);
void rp_linear(const rp_layer *l, float *out, const float *in, int *acc_out /* nullable */);
void rp_compute_rnn(const rp_model *m, rp_state *st, float *gains, float *vad, const float *features);

/* model / state / frame */
rp_model *rp_model_from_buffer(const void *blob, int len); /* NULL on malformed blob */
rp_model *rp_model_from_file(const char *path);
void rp_model_free(rp_model *m);
rp_state *rp_state_create(void);
void rp_state_destroy(rp_state *st);
int rp_state_size(void);
float rp_process_frame(const rp_model *m, rp_state *st, float *out, const float *in, rp_trace *tr);

#ifdef __cplusplus
}
#endif
#endif
