// rnn_kernels.cuh -- compute_rnn() (reference src/rnn.c:44-60) batched over the stream dimension.
//
// Arithmetic contract (oracle/rnnoise_port.c is the executable statement of it): identical to the
// reference's AVX2 kernels (src/vec_avx.h, src/nnet_arch.h) -- u8 = sat(rne(fma(x,127,127)))
// activations, exact s32 accumulation of u8 x s8 products, (float)acc*scale + subias, FMA'd
// recurrent diagonal, float layers as a sequential FMA chain over the inputs, Pade tanh/sigmoid --
// with the single substitution of a correctly rounded reciprocal for _mm256_rcp_ps.
// The translation unit is compiled with --fmad=false: every FMA below is explicit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define RNN_TS 8          // streams per CTA in the CUDA-core kernels
#define NB_FEAT 65
#define NB_GAINS 32

// Correctly rounded reciprocal of the Pade denominators (the one place the contract departs from the
// reference's _mm256_rcp_ps, DESIGN.md "Numerics").  den >= 952.7 by construction (even polynomial with
// positive coefficients), so for den < 2^126 this is exactly the in-range path of __frcp_rn -- MUFU.RCP and
// one FMA Newton step, the instructions the library routine executes after its exponent check -- without
// the check and the call; anything else (overflowed or NaN input) takes the library routine.
__device__ __forceinline__ float rcp_rn_den(float den) {
  if (den < 8.0e37f) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
    const float e = fmaf(den, r, -1.0f);
    return fmaf(r, -e, r);
  }
  return __frcp_rn(den);
}

__device__ __forceinline__ float act_tanh(float x) {   // tanh8_approx, vec_avx.h:398-416
  const float N0 = 952.52801514f, N1 = 96.39235687f, N2 = 0.60863042f;
  const float D0 = 952.72399902f, D1 = 413.36801147f, D2 = 11.88600922f;
  float x2 = x * x;
  float num = fmaf(fmaf(N2, x2, N1), x2, N0);
  float den = fmaf(fmaf(D2, x2, D1), x2, D0);
  num = num * x;
  den = rcp_rn_den(den);
  num = num * den;
  num = num < 1.f ? num : 1.f;
  return num > -1.f ? num : -1.f;
}
__device__ __forceinline__ float act_sigmoid(float x) { // sigmoid8_approx, vec_avx.h:426-445
  const float N0 = 238.13200378f, N1 = 6.02452230f, N2 = 0.00950985f;
  const float D0 = 952.72399902f, D1 = 103.34200287f, D2 = 0.74287558f;
  float x2 = x * x;
  float num = fmaf(fmaf(N2, x2, N1), x2, N0);
  float den = fmaf(fmaf(D2, x2, D1), x2, D0);
  num = num * x;
  den = rcp_rn_den(den);
  num = fmaf(num, den, .5f);
  num = num < 1.f ? num : 1.f;
  return num > 0.f ? num : 0.f;
}
// Straight-line forms for the tensor-core epilogues.  rcp_rn_den's range test is a (potentially divergent) branch plus
// a call per activation: twelve per thread and 16-unit slice, which also keeps the compiler from interleaving the
// independent dependency chains of a thread's units.  For |x| < ACT_FAST_LIMIT the Pade denominators stay below 8e37
// (11.886 * 1e36 + ...), i.e. inside the branch-free path, whose instructions are exactly those of rcp_rn_den's in-range
// path: the epilogues test all of a slice's pre-activations once and fall back to act_tanh / act_sigmoid otherwise.
#define ACT_FAST_LIMIT 1.0e9f
__device__ __forceinline__ float rcp_den_inrange(float den) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
  const float e = fmaf(den, r, -1.0f);
  return fmaf(r, -e, r);
}
__device__ __forceinline__ float act_tanh_inrange(float x) {
  const float N0 = 952.52801514f, N1 = 96.39235687f, N2 = 0.60863042f;
  const float D0 = 952.72399902f, D1 = 413.36801147f, D2 = 11.88600922f;
  float x2 = x * x;
  float num = fmaf(fmaf(N2, x2, N1), x2, N0);
  float den = fmaf(fmaf(D2, x2, D1), x2, D0);
  num = num * x;
  den = rcp_den_inrange(den);
  num = num * den;
  num = num < 1.f ? num : 1.f;
  return num > -1.f ? num : -1.f;
}
__device__ __forceinline__ float act_sigmoid_inrange(float x) {
  const float N0 = 238.13200378f, N1 = 6.02452230f, N2 = 0.00950985f;
  const float D0 = 952.72399902f, D1 = 103.34200287f, D2 = 0.74287558f;
  float x2 = x * x;
  float num = fmaf(fmaf(N2, x2, N1), x2, N0);
  float den = fmaf(fmaf(D2, x2, D1), x2, D0);
  num = num * x;
  den = rcp_den_inrange(den);
  num = fmaf(num, den, .5f);
  num = num < 1.f ? num : 1.f;
  return num > 0.f ? num : 0.f;
}
// One GRU unit update (compute_generic_gru, nnet.c:65-94) from the six dequantised pre-activations of a unit and its
// old state, for P units of a thread at once: sums first, one range test, then the branch-free block.
template <int P>
__device__ __forceinline__ void gru_units(const float (&zi)[P], const float (&ri)[P], const float (&ni)[P], const float (&zr)[P],
                                          const float (&rr)[P], const float (&nr)[P], const float (&h)[P], float (&out)[P]) {
  float zs[P], rs[P];
  bool ok = true;
#pragma unroll
  for (int q = 0; q < P; q++) {
    zs[q] = zi[q] + zr[q];
    rs[q] = ri[q] + rr[q];
    ok = ok && fabsf(zs[q]) < ACT_FAST_LIMIT && fabsf(rs[q]) < ACT_FAST_LIMIT && fabsf(ni[q]) + fabsf(nr[q]) < ACT_FAST_LIMIT;
  }
  if (ok) {
#pragma unroll
    for (int q = 0; q < P; q++) {
      const float z = act_sigmoid_inrange(zs[q]);
      const float r = act_sigmoid_inrange(rs[q]);
      const float n = act_tanh_inrange(ni[q] + nr[q] * r);
      out[q] = z * h[q] + (1 - z) * n;
    }
  } else {
#pragma unroll
    for (int q = 0; q < P; q++) {
      const float z = act_sigmoid(zs[q]);
      const float r = act_sigmoid(rs[q]);
      const float n = act_tanh(ni[q] + nr[q] * r);
      out[q] = z * h[q] + (1 - z) * n;
    }
  }
}
__device__ __forceinline__ uint32_t quant_u8(float x) { // vector_ps_to_epi8, vec_avx.h:326-341
  int v = __float2int_rn(fmaf(x, 127.f, 127.f));
  v = v < 0 ? 0 : v;
  return (uint32_t)(v > 255 ? 255 : v);
}
__device__ __forceinline__ uint32_t quant4(float a, float b, float c, float d) {
  return quant_u8(a) | (quant_u8(b) << 8) | (quant_u8(c) << 16) | (quant_u8(d) << 24);
}
__device__ __forceinline__ int dp4a_us(uint32_t u, int w, int acc) { // 4 x (u8 * s8) + s32, exact
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(u), "r"(w), "r"(acc));
  return d;
}

__device__ __forceinline__ void cp_async16(void *dst, const void *src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = valid ? 16 : 0;   // src-size 0 -> 16 bytes of zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

__device__ __forceinline__ void cp_async4(void *dst, const void *src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = valid ? 4 : 0;   // src-size 0 -> zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

// Programmatic dependent launch (PDL): the network kernels of one frame form a chain on one stream.
// pdl_trigger() lets the next kernel's CTAs be scheduled as soon as SM resources free up (its prologue
// -- barrier init, TMEM allocation, weight prefetch -- overlaps this kernel's tail); pdl_wait() in the
// dependent blocks until the previous grid has completed and flushed, before any of its results is read.
// Both are no-ops when the kernel was launched without the attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Device-resident model (built by engine.cu from the parsed blob).
struct DevLayerF { const float *w, *bias; };                               // w[in][out]
struct DevLayerQ {
  const int *wp;                          // wp[in/4][out] packed s8x4
  const float *scale, *subias, *diag;
  // recurrent GRU matrices only: the epilogue parameters of the LAYER, one 64-byte record per unit in the order the
  // tensor-core epilogues read them -- {scale_in, subias_in, scale_rec, subias_rec} for z, r, n, then {diag_z, diag_r,
  // diag_n, 0} -- so that a CTA stages its unit slice with contiguous 16-byte asynchronous copies
  const float *packed;                    // [gru][16]
};
struct DevModel {
  int cond, gru;
  DevLayerF conv1, dense_out, vad_dense;
  DevLayerQ conv2, gru_in[3], gru_rec[3];
};

// ------------------------------------------------------------------------------------------------
// conv1: [mem(2 frames) | features] (195) -> cond, fp32, tanh   (nnet.c:113-123, sgemv vec_avx.h:672)
// grid = ceil(S / RNN_TS), block = 128.
// It also maintains conv2's operand row  c2in[s] = u8([mem2(2 x cond) | conv1_out(cond)])  (the only
// form in which conv2 ever sees its inputs, vec_avx.h:326): on a non-silent frame the row is rotated
// left by cond bytes (compute_generic_conv1d's memory update, nnet.c:122) and the new output appended.
// ------------------------------------------------------------------------------------------------
#define C1_KC 32
#define C1_XS 196   // row stride of the staged inputs (195 + 1: keeps rows 16-byte aligned)
__global__ void __launch_bounds__(128) k_conv1(int S, DevModel m, const float *__restrict__ features,
                                               float *conv1_state, const int *__restrict__ silence,
                                               uint8_t *c2in, int ldc /* row stride of c2in: 3 * cond padded to 128 */) {
  __shared__ __align__(16) float tmp[RNN_TS][C1_XS];
  __shared__ __align__(16) float wsm[2][C1_KC][128];   // weight chunks, double-buffered (cond <= 128)
  __shared__ uint32_t rot[RNN_TS][64];                  // 2*cond/4 words per stream
  const int s0 = blockIdx.x * RNN_TS, tid = threadIdx.x, W = m.cond / 4, cond = m.cond;
  pdl_trigger();
  constexpr int KIN = 3 * NB_FEAT, NCH = (KIN + C1_KC - 1) / C1_KC;
  auto stage = [&](int c, int buf) {   // lane = 16-byte column, warp = row (mod 4): no index division
    const int j0 = c * C1_KC, rows = min(C1_KC, KIN - j0), q = tid & 31;
    if (q < W)
      for (int r = tid >> 5; r < rows; r += 4)
        cp_async16(&wsm[buf][r][4 * q], &m.conv1.w[(size_t)(j0 + r) * cond + 4 * q], true);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // inputs [state(2 frames) | features] and the conv2 input words to rotate: asynchronous 4-byte copies
  // (zero fill for absent streams), all in flight together, nothing held in registers
  for (int idx = tid; idx < RNN_TS * KIN; idx += 128) {
    const int s = idx / KIN, j = idx % KIN;
    const bool live = s0 + s < S;
    const size_t row = live ? s0 + s : 0;
    cp_async4(&tmp[s][j], j < 2 * NB_FEAT ? &conv1_state[row * 2 * NB_FEAT + j] : &features[row * NB_FEAT + j - 2 * NB_FEAT], live);
  }
  for (int s = tid >> 6; s < RNN_TS; s += 2) {   // words [W, 3W) of each live row; 2W <= 64 words per row
    const int w = tid & 63;
    if (w < 2 * W) cp_async4(&rot[s][w], (const uint32_t *)(c2in + (size_t)(s0 + s < S ? s0 + s : 0) * ldc) + W + w, s0 + s < S);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  stage(0, 0);
  asm volatile("cp.async.wait_group 1;" ::: "memory");   // inputs landed (the first weight chunk may still be in flight)
  __syncthreads();
  for (int s = tid >> 6; s < RNN_TS; s += 2) {
    const int w = tid & 63;
    if (w < 2 * W && s0 + s < S && !silence[s0 + s]) ((uint32_t *)(c2in + (size_t)(s0 + s) * ldc))[w] = rot[s][w];
  }
  const int o = tid;
  float acc[RNN_TS];
#pragma unroll
  for (int s = 0; s < RNN_TS; s++) acc[s] = 0.f;
  for (int c = 0; c < NCH; c++) {
    const int buf = c & 1, j0 = c * C1_KC, rows = min(C1_KC, KIN - j0);
    if (c + 1 < NCH) {
      stage(c + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (o < cond) {
      int jj = 0;
      for (; jj + 4 <= rows; jj += 4) {   // sequential FMA chain over the inputs, 4 at a time
        const float w0 = wsm[buf][jj][o], w1 = wsm[buf][jj + 1][o], w2 = wsm[buf][jj + 2][o], w3 = wsm[buf][jj + 3][o];
#pragma unroll
        for (int s = 0; s < RNN_TS; s++) {
          const float4 x = *(const float4 *)&tmp[s][j0 + jj];
          acc[s] = fmaf(w0, x.x, acc[s]); acc[s] = fmaf(w1, x.y, acc[s]);
          acc[s] = fmaf(w2, x.z, acc[s]); acc[s] = fmaf(w3, x.w, acc[s]);
        }
      }
      for (; jj < rows; jj++) {
        const float w = wsm[buf][jj][o];
#pragma unroll
        for (int s = 0; s < RNN_TS; s++) acc[s] = fmaf(w, tmp[s][j0 + jj], acc[s]);
      }
    }
    __syncthreads();
  }
  if (o < cond) {
    const float b = m.conv1.bias[o];
#pragma unroll
    for (int s = 0; s < RNN_TS; s++)
      if (s0 + s < S && !silence[s0 + s])
        c2in[(size_t)(s0 + s) * ldc + 2 * cond + o] = (uint8_t)quant_u8(act_tanh(acc[s] + b));
  }
  // memory update: mem = tmp[65:195]; silent frames leave the state untouched (denoise.c:474)
  for (int idx = tid; idx < RNN_TS * 2 * NB_FEAT; idx += 128) {
    int s = idx / (2 * NB_FEAT), j = idx % (2 * NB_FEAT);
    if (s0 + s < S && !silence[s0 + s]) conv1_state[(size_t)(s0 + s) * 2 * NB_FEAT + j] = tmp[s][NB_FEAT + j];
  }
}

// ------------------------------------------------------------------------------------------------
// conv2 on CUDA cores (cross-check kernel for k_tc2<false>): c2in (3*cond u8) -> gru, tanh
// (cgemv8x4 vec_avx.h:829).  grid = ceil(S / RNN_TS), block = 128, dynamic smem = RNN_TS*(3*cond/4)*4
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_conv2(int S, DevModel m, const uint8_t *__restrict__ c2in, int ldc,
                                               float *__restrict__ conv2_out, uint8_t *__restrict__ conv2_out_u8, int ldo) {
  extern __shared__ uint32_t u_sm[];   // [RNN_TS][K/4]
  const int K = 3 * m.cond, K4 = K / 4, s0 = blockIdx.x * RNN_TS, tid = threadIdx.x;
  for (int idx = tid; idx < RNN_TS * K4; idx += 128) {
    int s = idx / K4;
    u_sm[idx] = s0 + s < S ? ((const uint32_t *)(c2in + (size_t)(s0 + s) * ldc))[idx % K4] : 0u;
  }
  __syncthreads();
  for (int o = tid; o < m.gru; o += 128) {
    int acc[RNN_TS];
#pragma unroll
    for (int s = 0; s < RNN_TS; s++) acc[s] = 0;
    for (int k4 = 0; k4 < K4; k4++) {
      int w = __ldg(&m.conv2.wp[(size_t)k4 * m.gru + o]);
#pragma unroll
      for (int s = 0; s < RNN_TS; s++) acc[s] = dp4a_us(u_sm[s * K4 + k4], w, acc[s]);
    }
    const float sc = m.conv2.scale[o], sb = m.conv2.subias[o];
#pragma unroll
    for (int s = 0; s < RNN_TS; s++)
      if (s0 + s < S) {
        float v = act_tanh((float)acc[s] * sc + sb);
        conv2_out[(size_t)(s0 + s) * m.gru + o] = v;
        conv2_out_u8[(size_t)(s0 + s) * ldo + o] = (uint8_t)quant_u8(v);   // operand of the GRU-1 GEMM
      }
  }
}

// ------------------------------------------------------------------------------------------------
// One GRU layer (compute_generic_gru nnet.c:65-94; sparse_cgemv8x4 vec_avx.h:778 executed dense:
// the zero blocks contribute exact zeros).  Thread = hidden unit, RNN_TS streams per CTA.
// grid = (ceil(S / RNN_TS), gru / 128), block = 128, dynamic smem = 2 * RNN_TS * (gru/4) * 4 bytes
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_gru(int S, int gru, DevLayerQ wi, DevLayerQ wr,
                                             const float *__restrict__ x, const float *__restrict__ h_old,
                                             float *__restrict__ h_new, uint8_t *__restrict__ h_new_u8, int ldo,
                                             const int *__restrict__ silence) {
  extern __shared__ uint32_t u_sm[];
  const int K4 = gru / 4, s0 = blockIdx.x * RNN_TS, tid = threadIdx.x;
  uint32_t *xu = u_sm, *hu = u_sm + RNN_TS * K4;
  for (int idx = tid; idx < RNN_TS * K4; idx += 128) {
    int s = idx / K4, k4 = idx % K4;
    uint32_t qx = 0, qh = 0;
    if (s0 + s < S) {
      float4 a = *(const float4 *)&x[(size_t)(s0 + s) * gru + 4 * k4];
      float4 b = *(const float4 *)&h_old[(size_t)(s0 + s) * gru + 4 * k4];
      qx = quant4(a.x, a.y, a.z, a.w);
      qh = quant4(b.x, b.y, b.z, b.w);
    }
    xu[idx] = qx; hu[idx] = qh;
  }
  __syncthreads();
  const int j = blockIdx.y * 128 + tid;
  if (j >= gru) return;
  int ai[3][RNN_TS], ar[3][RNN_TS];
#pragma unroll
  for (int g = 0; g < 3; g++)
#pragma unroll
    for (int s = 0; s < RNN_TS; s++) { ai[g][s] = 0; ar[g][s] = 0; }
  for (int k4 = 0; k4 < K4; k4++) {
    int wiz = __ldg(&wi.wp[(size_t)k4 * 3 * gru + j]), wir = __ldg(&wi.wp[(size_t)k4 * 3 * gru + gru + j]),
        win = __ldg(&wi.wp[(size_t)k4 * 3 * gru + 2 * gru + j]);
    int wrz = __ldg(&wr.wp[(size_t)k4 * 3 * gru + j]), wrr = __ldg(&wr.wp[(size_t)k4 * 3 * gru + gru + j]),
        wrn = __ldg(&wr.wp[(size_t)k4 * 3 * gru + 2 * gru + j]);
#pragma unroll
    for (int s = 0; s < RNN_TS; s++) {
      uint32_t xv = xu[s * K4 + k4], hv = hu[s * K4 + k4];
      ai[0][s] = dp4a_us(xv, wiz, ai[0][s]);
      ai[1][s] = dp4a_us(xv, wir, ai[1][s]);
      ai[2][s] = dp4a_us(xv, win, ai[2][s]);
      ar[0][s] = dp4a_us(hv, wrz, ar[0][s]);
      ar[1][s] = dp4a_us(hv, wrr, ar[1][s]);
      ar[2][s] = dp4a_us(hv, wrn, ar[2][s]);
    }
  }
  float sci[3], sbi[3], scr[3], sbr[3], dg[3];
#pragma unroll
  for (int g = 0; g < 3; g++) {
    sci[g] = wi.scale[g * gru + j]; sbi[g] = wi.subias[g * gru + j];
    scr[g] = wr.scale[g * gru + j]; sbr[g] = wr.subias[g * gru + j];
    dg[g] = wr.diag[g * gru + j];
  }
#pragma unroll
  for (int s = 0; s < RNN_TS; s++) {
    if (s0 + s >= S) continue;
    const float h = h_old[(size_t)(s0 + s) * gru + j];
    float out = h;
    if (!silence[s0 + s]) {
      float zi = (float)ai[0][s] * sci[0] + sbi[0];
      float ri = (float)ai[1][s] * sci[1] + sbi[1];
      float ni = (float)ai[2][s] * sci[2] + sbi[2];
      float zr = fmaf(dg[0], h, (float)ar[0][s] * scr[0] + sbr[0]);
      float rr = fmaf(dg[1], h, (float)ar[1][s] * scr[1] + sbr[1]);
      float nr = fmaf(dg[2], h, (float)ar[2][s] * scr[2] + sbr[2]);
      float z = act_sigmoid(zi + zr);
      float r = act_sigmoid(ri + rr);
      float n = act_tanh(ni + nr * r);
      out = z * h + (1 - z) * n;
    }
    h_new[(size_t)(s0 + s) * gru + j] = out;
    h_new_u8[(size_t)(s0 + s) * ldo + j] = (uint8_t)quant_u8(out);
  }
}

// ------------------------------------------------------------------------------------------------
// Output heads on cat = [conv2_out | gru1 | gru2 | gru3] (rnn.c:53-57): dense_out (32, sigmoid;
// sequential FMA chain over the 4*gru inputs) and vad_dense (1, sigmoid; the reference's scalar
// tail multiplies and adds separately, vec_avx.h:731-735).
// grid = ceil(S / 16), block = 160: warps 0..3 own 4 streams each (lane = output, 4 accumulator
// chains per thread), warp 4 runs the 16 VAD chains (lane = stream).  Activations and weights stream
// through a double-buffered cp.async pipeline in chunks of 64 inputs, so the FMA chains only ever
// wait on shared memory while the next chunk is in flight.
// ------------------------------------------------------------------------------------------------
#define HEAD_TS 16
#define HEAD_KC 64
#define HEAD_XS (HEAD_KC + 4)
__global__ void __launch_bounds__(160) k_heads(int S, DevModel m, const float *__restrict__ c2,
                                               const float *__restrict__ g1, const float *__restrict__ g2,
                                               const float *__restrict__ g3, const int *__restrict__ silence,
                                               float *__restrict__ gains, float *__restrict__ vad,
                                               float *__restrict__ vad_user, int vad_stride) {
  __shared__ __align__(16) float xs[2][HEAD_TS][HEAD_XS];
  __shared__ __align__(16) float ws[2][HEAD_KC][NB_GAINS];
  __shared__ __align__(16) float wv[2][HEAD_KC];
  const int s0 = blockIdx.x * HEAD_TS, tid = threadIdx.x, gru = m.gru, K = 4 * gru, nchunk = K / HEAD_KC;
  const int o = tid & 31, sg = tid >> 5;
  pdl_trigger();
  pdl_wait();   // GRU-3 state of this frame
  auto stage = [&](int c, int buf) {   // gru % 64 == 0: a chunk never straddles two source arrays
    const int c0 = c * HEAD_KC, src = c0 / gru, off = c0 % gru;
    const float *p = src == 0 ? c2 : src == 1 ? g1 : src == 2 ? g2 : g3;
    for (int idx = tid; idx < HEAD_TS * HEAD_KC / 4; idx += 160) {
      int s = idx / (HEAD_KC / 4), k4 = idx % (HEAD_KC / 4);
      const bool live = s0 + s < S;
      cp_async16(&xs[buf][s][4 * k4], &p[(size_t)(live ? s0 + s : 0) * gru + off + 4 * k4], live);
    }
    for (int idx = tid; idx < HEAD_KC * NB_GAINS / 4; idx += 160)
      cp_async16(&ws[buf][0][0] + 4 * idx, &m.dense_out.w[(size_t)c0 * NB_GAINS + 4 * idx], true);
    if (tid < HEAD_KC / 4) cp_async16(&wv[buf][4 * tid], &m.vad_dense.w[c0 + 4 * tid], true);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float y = 0.f;
  stage(0, 0);
  for (int c = 0; c < nchunk; c++) {
    const int buf = c & 1;
    if (c + 1 < nchunk) {
      stage(c + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (sg < 4) {
#pragma unroll 4
      for (int kk = 0; kk < HEAD_KC; kk += 4) {
        float4 x0 = *(const float4 *)&xs[buf][sg * 4 + 0][kk], x1 = *(const float4 *)&xs[buf][sg * 4 + 1][kk];
        float4 x2 = *(const float4 *)&xs[buf][sg * 4 + 2][kk], x3 = *(const float4 *)&xs[buf][sg * 4 + 3][kk];
        float w0 = ws[buf][kk][o], w1 = ws[buf][kk + 1][o], w2 = ws[buf][kk + 2][o], w3 = ws[buf][kk + 3][o];
        acc[0] = fmaf(w0, x0.x, acc[0]); acc[1] = fmaf(w0, x1.x, acc[1]); acc[2] = fmaf(w0, x2.x, acc[2]); acc[3] = fmaf(w0, x3.x, acc[3]);
        acc[0] = fmaf(w1, x0.y, acc[0]); acc[1] = fmaf(w1, x1.y, acc[1]); acc[2] = fmaf(w1, x2.y, acc[2]); acc[3] = fmaf(w1, x3.y, acc[3]);
        acc[0] = fmaf(w2, x0.z, acc[0]); acc[1] = fmaf(w2, x1.z, acc[1]); acc[2] = fmaf(w2, x2.z, acc[2]); acc[3] = fmaf(w2, x3.z, acc[3]);
        acc[0] = fmaf(w3, x0.w, acc[0]); acc[1] = fmaf(w3, x1.w, acc[1]); acc[2] = fmaf(w3, x2.w, acc[2]); acc[3] = fmaf(w3, x3.w, acc[3]);
      }
    } else if (o < HEAD_TS) {
#pragma unroll 8
      for (int kk = 0; kk < HEAD_KC; kk++) y = y + wv[buf][kk] * xs[buf][o][kk];
    }
    __syncthreads();   // everyone is done with `buf` before chunk c+2 is staged into it
  }
  if (sg < 4) {
    const float b = m.dense_out.bias[o];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      int s = s0 + sg * 4 + q;
      if (s < S) gains[(size_t)s * NB_GAINS + o] = act_sigmoid(acc[q] + b);
    }
  } else if (o < HEAD_TS) {
    int s = s0 + o;
    if (s < S) {
      float v = silence[s] ? 0.f : act_sigmoid(y + m.vad_dense.bias[0]);
      vad[s] = v;
      if (vad_user) vad_user[(size_t)s * vad_stride] = v;
    }
  }
}
