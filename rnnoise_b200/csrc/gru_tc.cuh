// gru_tc.cuh -- one GRU layer for 128 streams x 32 hidden units per CTA on the 5th-gen tensor cores.
//
//   D_in [128 x 96] = Xu8[128 x K] . Wi_slice^T      (u8 x s8 -> s32, tcgen05.mma kind::i8)
//   D_rec[128 x 96] = Hu8[128 x K] . Wr_slice^T      96 = {z, r, n} x 32 units, K = gru (384)
//
// Operands arrive by TMA (cp.async.bulk.tensor, SWIZZLE_128B, K-major) straight from the u8 mirrors
// of the activations / the pre-permuted s8 weights; both accumulators live in TMEM (192 of 256
// allocated columns); one elected thread issues the 2 x (K/32) MMAs and commits to an mbarrier;
// the four epilogue warps read their TMEM lane quarter with tcgen05.ld and apply, in registers,
// exactly the arithmetic of the reference (compute_linear + compute_generic_gru, src/nnet_arch.h:
// 130-162, src/nnet.c:65-94): (float)acc*scale + subias, fma(diag,h,.), sigmoid/sigmoid/tanh,
// h' = z*h + (1-z)*n, then store h' as fp32 AND as the u8 operand of the next consumer.
// The integer accumulators are exact, so this kernel is bit-identical to the dp4a kernel k_gru.
//
// grid = (ceil(S/128), gru/32), block = 160 (warps 0..3 epilogue, warp 4 = TMA + MMA issuer),
// dynamic smem = GRU_TC_SMEM bytes, 1 CTA / SM.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "rnn_kernels.cuh"

#define TC_M 128          // streams per CTA (UMMA M)
#define TC_UNITS 32       // hidden units per CTA
#define TC_N (3 * TC_UNITS)  // UMMA N = 96
#define TC_KATOM 128      // bytes of K per 128B-swizzle atom
#define TC_TMEM_COLS 256  // power of two >= 2 * TC_N

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a lost transaction (bad tensor map / byte count) becomes a trap -> CUDA error on the
// host instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 | LBO(=1)<<16 | SBO(1024>>4)<<32 | version 1 <<46 | SW128(2)<<61)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// cute::UMMA::InstrDescriptor for kind::i8: D = S32, A = U8, B = S8, both K-major
__device__ __forceinline__ uint32_t umma_idesc_i8(int M, int N) {
  return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

struct GruTcMaps {
  CUtensorMap x, h, wi, wr;   // x,h: u8 [S][K]; wi,wr: s8 [(K/32 slices) * 96][K] permuted
};

// smem: A tiles (X, H): 2 x (K/128) x 16 KB ; B tiles (Wi, Wr): 2 x (K/128) x 12 KB ; params ; barriers
#define TC_A_ATOM_BYTES (TC_M * TC_KATOM)     // 16384
#define TC_B_ATOM_BYTES (TC_N * TC_KATOM)     // 12288
__host__ __device__ constexpr int gru_tc_smem_bytes(int gru) {
  return 1024 /*align slack*/ + 2 * (gru / TC_KATOM) * (TC_A_ATOM_BYTES + TC_B_ATOM_BYTES) + 15 * TC_UNITS * 4 + 64;
}

__global__ void __launch_bounds__(160, 1)
k_gru_tc(int S, int gru, const __grid_constant__ GruTcMaps maps, DevLayerQ wi, DevLayerQ wr,
         const float *__restrict__ h_old, float *__restrict__ h_new, uint8_t *__restrict__ h_new_u8,
         const int *__restrict__ silence) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * TC_M, j0 = blockIdx.y * TC_UNITS, natoms = gru / TC_KATOM;
  // 1024-byte aligned operand area (SWIZZLE_128B requirement); plain pointer arithmetic on the shared
  // array keeps the address space known to the compiler (LDS/STS instead of generic loads)
  uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t *sAx = base, *sAh = sAx + natoms * TC_A_ATOM_BYTES;
  uint8_t *sBi = sAh + natoms * TC_A_ATOM_BYTES, *sBr = sBi + natoms * TC_B_ATOM_BYTES;
  float *prm = (float *)(sBr + natoms * TC_B_ATOM_BYTES);       // [15][32]
  uint64_t *bars = (uint64_t *)(prm + 15 * TC_UNITS);           // [0] operands landed, [1] MMAs done
  uint32_t *tmem_slot = (uint32_t *)(bars + 2);
  const uint32_t bar_full = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);

  if (tid == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {   // TMEM allocation is warp-wide; the same warp frees it at the end
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // epilogue parameters of this unit slice: [g] scale_i, subias_i, scale_r, subias_r, diag
  for (int i = tid; i < 15 * TC_UNITS; i += blockDim.x) {
    int which = i / (3 * TC_UNITS), g = (i / TC_UNITS) % 3, u = i % TC_UNITS;
    const float *src = which == 0 ? wi.scale : which == 1 ? wi.subias : which == 2 ? wr.scale : which == 3 ? wr.subias : wr.diag;
    prm[i] = src[g * gru + j0 + u];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      // ---- TMA producer: everything this CTA needs, one transaction barrier ----
      mbar_expect_tx(bar_full, (uint32_t)(2 * natoms * (TC_A_ATOM_BYTES + TC_B_ATOM_BYTES)));
      for (int a = 0; a < natoms; a++) {
        tma_load_2d(smem_u32(sAx + a * TC_A_ATOM_BYTES), &maps.x, bar_full, a * TC_KATOM, m0);
        tma_load_2d(smem_u32(sBi + a * TC_B_ATOM_BYTES), &maps.wi, bar_full, a * TC_KATOM, blockIdx.y * TC_N);
      }
      for (int a = 0; a < natoms; a++) {
        tma_load_2d(smem_u32(sAh + a * TC_A_ATOM_BYTES), &maps.h, bar_full, a * TC_KATOM, m0);
        tma_load_2d(smem_u32(sBr + a * TC_B_ATOM_BYTES), &maps.wr, bar_full, a * TC_KATOM, blockIdx.y * TC_N);
      }
      // ---- MMA issuer ----
      mbar_wait(bar_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t idesc = umma_idesc_i8(TC_M, TC_N);
      for (int g = 0; g < 2; g++) {
        const uint8_t *A = g ? sAh : sAx, *B = g ? sBr : sBi;
        for (int a = 0; a < natoms; a++) {
          const uint64_t ad = umma_desc_sw128(smem_u32(A + a * TC_A_ATOM_BYTES));
          const uint64_t bd = umma_desc_sw128(smem_u32(B + a * TC_B_ATOM_BYTES));
#pragma unroll
          for (int k = 0; k < TC_KATOM / 32; k++)   // UMMA K = 32 bytes; advance inside the swizzle atom
            umma_i8(tmem + g * TC_N, ad + (uint64_t)(k * 32 >> 4), bd + (uint64_t)(k * 32 >> 4), idesc, (a | k) ? 1u : 0u);
        }
      }
      umma_commit(bar_mma);   // implies tcgen05.fence::before_thread_sync
    }
  } else {
    // ---- epilogue warps: thread = stream row (TMEM lane 32*warp + lane) ----
    const int s = m0 + warp * 32 + lane;
    const bool live = s < S;
    const bool silent = live ? silence[s] != 0 : true;
    float hrow[TC_UNITS];
    if (live) {
#pragma unroll
      for (int q = 0; q < TC_UNITS / 4; q++) {
        float4 v = *(const float4 *)&h_old[(size_t)s * gru + j0 + 4 * q];
        hrow[4 * q] = v.x; hrow[4 * q + 1] = v.y; hrow[4 * q + 2] = v.z; hrow[4 * q + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int q = 0; q < TC_UNITS; q++) hrow[q] = 0.f;
    }
    mbar_wait(bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int half = 0; half < 2; half++) {
      int az[16], ar[16], an[16], bz[16], br[16], bn[16];
      const int c = half * 16;
      tmem_ld16(trow + 0 * TC_UNITS + c, az); tmem_ld16(trow + 1 * TC_UNITS + c, ar); tmem_ld16(trow + 2 * TC_UNITS + c, an);
      tmem_ld16(trow + TC_N + 0 * TC_UNITS + c, bz); tmem_ld16(trow + TC_N + 1 * TC_UNITS + c, br); tmem_ld16(trow + TC_N + 2 * TC_UNITS + c, bn);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float outv[16];
#pragma unroll
      for (int q = 0; q < 16; q++) {
        const int u = c + q;
        const float h = hrow[u];
        float out = h;
        if (!silent) {
          float zi = (float)az[q] * prm[(0 * 3 + 0) * TC_UNITS + u] + prm[(1 * 3 + 0) * TC_UNITS + u];
          float ri = (float)ar[q] * prm[(0 * 3 + 1) * TC_UNITS + u] + prm[(1 * 3 + 1) * TC_UNITS + u];
          float ni = (float)an[q] * prm[(0 * 3 + 2) * TC_UNITS + u] + prm[(1 * 3 + 2) * TC_UNITS + u];
          float zr = fmaf(prm[(4 * 3 + 0) * TC_UNITS + u], h, (float)bz[q] * prm[(2 * 3 + 0) * TC_UNITS + u] + prm[(3 * 3 + 0) * TC_UNITS + u]);
          float rr = fmaf(prm[(4 * 3 + 1) * TC_UNITS + u], h, (float)br[q] * prm[(2 * 3 + 1) * TC_UNITS + u] + prm[(3 * 3 + 1) * TC_UNITS + u]);
          float nr = fmaf(prm[(4 * 3 + 2) * TC_UNITS + u], h, (float)bn[q] * prm[(2 * 3 + 2) * TC_UNITS + u] + prm[(3 * 3 + 2) * TC_UNITS + u]);
          float z = act_sigmoid(zi + zr);
          float r = act_sigmoid(ri + rr);
          float n = act_tanh(ni + nr * r);
          out = z * h + (1 - z) * n;
        }
        outv[q] = out;
      }
      if (live) {
        float *dst = &h_new[(size_t)s * gru + j0 + c];
#pragma unroll
        for (int q = 0; q < 4; q++) *(float4 *)&dst[4 * q] = make_float4(outv[4 * q], outv[4 * q + 1], outv[4 * q + 2], outv[4 * q + 3]);
        uint4 pk;
        pk.x = quant4(outv[0], outv[1], outv[2], outv[3]);
        pk.y = quant4(outv[4], outv[5], outv[6], outv[7]);
        pk.z = quant4(outv[8], outv[9], outv[10], outv[11]);
        pk.w = quant4(outv[12], outv[13], outv[14], outv[15]);
        *(uint4 *)&h_new_u8[(size_t)s * gru + j0 + c] = pk;
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS) : "memory");
  }
}

// ================================================================================================
// k_tc2<kGru> -- persistent, warp-specialised tensor-core kernel (the default int8 path).
//
//   kGru = true : one GRU layer (k_gru_tc2).   kGru = false: conv2 (k_conv2_tc), a single GEMM + tanh.
//
// CTA = 128 streams x (N/4) output units, processed as slices of 16 units through a two-deep TMA ring
// (TC2_STAGES) for the weight slices and a two-deep TMEM ring for the accumulators:
//     warp 16 (one elected thread): TMA producer + tcgen05.mma issuer
//     warps 0..15 (512 threads)   : epilogue -- warp w reads TMEM lane quarter (w & 3), units 4*(w>>2)..+4
// so that  TMA / MMA(slice s+1) || epilogue(slice s).  The u8 activation tiles (128 x K;
// GRU: Xu8 and Hu8) are loaded once per CTA and stay resident.  Per slice and matrix one
// tcgen05.mma.kind::i8 chain of K/32 instructions, M128 x N48 (GRU: z|r|n of 16 units) or N16 (conv2);
// accumulators: GRU [in z|r|n (48) | rec z|r|n (48)] = 96 TMEM columns per stage, conv2 16.
// Same arithmetic as the dp4a kernels k_gru / k_conv2: bit-identical results.
// grid = (ceil(S/128), 4), block = 544, 1 CTA / SM.
// ================================================================================================
#define P_SLICE 16
#ifndef P_STAGES
#define P_STAGES 3                        // weight ring of the fused network kernel (net_kernel.cuh)
#endif
// Weight ring of the per-layer kernels: TWO stages.  The epilogue of a slice (2.7 us) is far longer than a slice's TMA +
// MMAs, so the third stage never ran ahead usefully, while its 37 KB are what lets three CTAs of the DSP kernels share
// an SM with a GRU CTA (same-box A/B at 4096 streams, profiles/r2p: 0.2804 -> 0.2775 ms per step, twice).
#ifndef TC2_STAGES
#define TC2_STAGES 2
#endif
// How many slices ahead the epilogue fetches the old state (1 or 2).  Two slices ahead take a layer from 22.3 to 21.0 us
// at 4096 streams but cost 8 registers per thread (94 instead of 86), which leaves room for one CTA less of the DSP
// kernels beside a GRU CTA: the pipelined step is slower with it (0.2790 vs 0.2770 ms, profiles/r2q_ab_4096.txt).
#ifndef TC2_HAHEAD
#define TC2_HAHEAD 1
#endif
// L2 prefetch of the CTA's old-state rows at kernel start (no registers held): the state was written a whole frame ago
// and has left the L2 at large batch sizes, so the per-slice gathers of the epilogue otherwise wait on HBM.  Same-box
// A/B at 4096 streams (profiles/r2t_ab_4096.txt): 0.2570 -> 0.2564 ms per step -- inside the noise, kept on because it
// costs nothing; it is the build the r2t evidence was measured with.
#ifndef TC2_H_L2PF
#define TC2_H_L2PF 1
#endif
#define P_TMEM_COLS 256                   // >= 2 stages x 96 columns, power of two

template <bool kGru> struct TcCfg {
  static constexpr int kMats = kGru ? 2 : 1;                 // GEMMs per slice (input, recurrent)
  static constexpr int kN = kGru ? 3 * P_SLICE : P_SLICE;    // UMMA N: 48 / 16
  static constexpr int kBAtom = kN * TC_KATOM;               // bytes of one weight atom: 6144 / 2048
  static constexpr int kPrm = kGru ? 16 : 2;                 // epilogue parameters per unit (GRU: 4 float4, see below)
  static constexpr int kCols = kMats * kN;                   // TMEM columns per stage: 96 / 16
};
template <bool kGru>
__host__ __device__ constexpr int tc2_smem_bytes(int K, int N) {
  return 1024 + TcCfg<kGru>::kMats * (K / TC_KATOM) * TC_A_ATOM_BYTES +
         TC2_STAGES * TcCfg<kGru>::kMats * (K / TC_KATOM) * TcCfg<kGru>::kBAtom + TcCfg<kGru>::kPrm * (N / 4) * 4 + 16 * 8 + 64;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
#define P_EPI_WARPS 16                    // epilogue warps: 4 per TMEM lane quarter
#define P_UPT (P_SLICE / (P_EPI_WARPS / 4))   // units per epilogue thread and slice: 4
#define P_THREADS (32 * (P_EPI_WARPS + 1))
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, int (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}

// K = contraction length = bytes of one (zero-weight padded) activation row, a multiple of 128; N = number of output
// units (gru, a multiple of 64); ldo = row stride of the u8 output mirror (the padded K of its consumer).
// GRU : maps.x/h = Xu8/Hu8 [S][K]; maps.wi/wr = s8 [(N/16) x 48][K];  out = h_new (+u8), aux = h_old
// conv: maps.x = conv2 input u8 [S][K]; maps.wi = s8 [N][K] (unit-major); out = conv2_out (+u8)
template <bool kGru>
__global__ void __launch_bounds__(P_THREADS, 1)
k_tc2(int S, int K, int N, int ldo, const __grid_constant__ GruTcMaps maps, DevLayerQ wi, DevLayerQ wr,
      const float *__restrict__ h_old, float *__restrict__ out_f32, uint8_t *__restrict__ out_u8,
      const int *__restrict__ silence) {
  using C = TcCfg<kGru>;
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int natoms = K / TC_KATOM, upc = N / 4, nslice = upc / P_SLICE;   // units / slices per CTA
  const int m0 = blockIdx.x * TC_M, jq = blockIdx.y * upc;
  // 1024-byte aligned operand area; pointer arithmetic on the shared array keeps the address space
  // known to the compiler (LDS instead of generic loads for the epilogue parameters)
  uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t *sAx = base, *sAh = sAx + natoms * TC_A_ATOM_BYTES;
  uint8_t *sB = sAx + C::kMats * natoms * TC_A_ATOM_BYTES;
  const int stage_bytes = C::kMats * natoms * C::kBAtom;
  float *prm = (float *)(sB + TC2_STAGES * stage_bytes);            // [kPrm][upc]
  uint64_t *bars = (uint64_t *)(prm + C::kPrm * upc);
  uint32_t *tmem_slot = (uint32_t *)(bars + 16);
  const uint32_t bar_a = smem_u32(&bars[0]);
  auto bar_bfull = [&](int i) { return smem_u32(&bars[1 + i]); };
  auto bar_bempty = [&](int i) { return smem_u32(&bars[4 + i]); };
  auto bar_tfull = [&](int i) { return smem_u32(&bars[7 + i]); };
  auto bar_tempty = [&](int i) { return smem_u32(&bars[9 + i]); };

  pdl_trigger();
  // The producer thread initialises the barriers ITSELF and starts the weight and operand-tile loads right away:
  // they overlap the TMEM allocation and the (scattered) parameter staging below instead of following them (the
  // other threads touch the barriers only after the __syncthreads that ends the prologue).
  auto load_B0 = [&](int s) {
    const int st = s % TC2_STAGES;
    uint8_t *dst = sB + st * stage_bytes;
    const int row = (blockIdx.y * nslice + s) * C::kN;
    mbar_expect_tx(bar_bfull(st), (uint32_t)stage_bytes);
    for (int a = 0; a < natoms; a++) {
      tma_load_2d(smem_u32(dst + a * C::kBAtom), &maps.wi, bar_bfull(st), a * TC_KATOM, row);
      if (kGru) tma_load_2d(smem_u32(dst + (natoms + a) * C::kBAtom), &maps.wr, bar_bfull(st), a * TC_KATOM, row);
    }
  };
  if (warp == P_EPI_WARPS && lane == 0) {
    mbar_init(bar_a, 1);
    for (int i = 0; i < TC2_STAGES; i++) { mbar_init(bar_bfull(i), 1); mbar_init(bar_bempty(i), 1); }
    for (int i = 0; i < 2; i++) { mbar_init(bar_tfull(i), 1); mbar_init(bar_tempty(i), P_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // first what the first slice needs (its weights, then the operand tiles), then the second weight stage: every CTA
    // of the grid starts here at the same time and the burst is bound by L2 bandwidth (17 MB at 4096 streams)
    load_B0(0);                                                    // weights: independent of the previous kernel
    pdl_wait();                                                    // activations of this frame are complete
    mbar_expect_tx(bar_a, (uint32_t)(C::kMats * natoms * TC_A_ATOM_BYTES));
    for (int a = 0; a < natoms; a++) {
      tma_load_2d(smem_u32(sAx + a * TC_A_ATOM_BYTES), &maps.x, bar_a, a * TC_KATOM, m0);
      if (kGru) tma_load_2d(smem_u32(sAh + a * TC_A_ATOM_BYTES), &maps.h, bar_a, a * TC_KATOM, m0);
    }
    for (int s = 1; s < TC2_STAGES && s < nslice; s++) load_B0(s);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(P_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (kGru) {
    // per unit u: {sc_i, sb_i, sc_r, sb_r} for z, r, n, then {diag_z, diag_r, diag_n, 0} (four LDS.128 in the epilogue):
    // this CTA's slice of the layer's packed records (DevLayerQ::packed) is contiguous -- 16-byte asynchronous copies
    // issued by the epilogue warps, which wait for them only right before their first slice (the three dependent
    // scattered loads per thread of the earlier staging loop were 8-14 % of this kernel's warp time, profiles/r2p)
    const float *src = wr.packed + (size_t)jq * 16;
    if (warp < P_EPI_WARPS)
      for (int c = tid; c < 4 * upc; c += 32 * P_EPI_WARPS) cp_async16(prm + 4 * c, src + 4 * c, true);
    asm volatile("cp.async.commit_group;" ::: "memory");
  } else {
    for (int i = tid; i < C::kPrm * upc; i += blockDim.x) prm[i] = (i < upc ? wi.scale : wi.subias)[jq + i % upc];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == P_EPI_WARPS) {
    if (lane == 0) {
      auto load_B = [&](int s) {
        const int st = s % TC2_STAGES;
        uint8_t *dst = sB + st * stage_bytes;
        const int row = (blockIdx.y * nslice + s) * C::kN;
        mbar_expect_tx(bar_bfull(st), (uint32_t)stage_bytes);
        for (int a = 0; a < natoms; a++) {
          tma_load_2d(smem_u32(dst + a * C::kBAtom), &maps.wi, bar_bfull(st), a * TC_KATOM, row);
          if (kGru) tma_load_2d(smem_u32(dst + (natoms + a) * C::kBAtom), &maps.wr, bar_bfull(st), a * TC_KATOM, row);
        }
      };
      mbar_wait(bar_a, 0);   // (weights and operand tiles were requested in the prologue)
      const uint32_t idesc = umma_idesc_i8(TC_M, C::kN);
      for (int s = 0; s < nslice; s++) {
        const int st = s % TC2_STAGES, ts = s & 1;
        mbar_wait(bar_bfull(st), (uint32_t)((s / TC2_STAGES) & 1));
        mbar_wait(bar_tempty(ts), (uint32_t)(((s >> 1) & 1) ^ 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint8_t *Bs = sB + st * stage_bytes;
        for (int g = 0; g < C::kMats; g++) {
          const uint8_t *A = g ? sAh : sAx;
          for (int a = 0; a < natoms; a++) {
            const uint64_t ad = umma_desc_sw128(smem_u32(A + a * TC_A_ATOM_BYTES));
            const uint64_t bd = umma_desc_sw128(smem_u32(Bs + (g * natoms + a) * C::kBAtom));
#pragma unroll
            for (int k = 0; k < TC_KATOM / 32; k++)
              umma_i8(tmem + ts * C::kCols + g * C::kN, ad + (uint64_t)(k * 32 >> 4), bd + (uint64_t)(k * 32 >> 4), idesc, (a | k) ? 1u : 0u);
          }
        }
        umma_commit(bar_bempty(st));   // weight stage reusable once these MMAs retire
        umma_commit(bar_tfull(ts));    // accumulators of slice s ready for the epilogue
        if (s >= 1 && s - 1 + TC2_STAGES < nslice) {   // refill the stage slice s-1 used
          mbar_wait(bar_bempty((s - 1) % TC2_STAGES), (uint32_t)(((s - 1) / TC2_STAGES) & 1));
          load_B(s - 1 + TC2_STAGES);
        }
      }
    }
  } else {
    const int lq = warp & 3, ch = warp >> 2;
    const int srow = m0 + lq * 32 + lane;
    const bool live = srow < S;
    const bool silent = live ? silence[srow] != 0 : true;
    const uint32_t trow = tmem + ((uint32_t)(lq * 32) << 16);
    float hcur[P_UPT], hnext[P_UPT], hnext2[P_UPT];   // old state: fetched two slices ahead (a strided 16-byte gather from HBM)
    auto load_h = [&](int s, float (&dst)[P_UPT]) {
      if (kGru && live && s < nslice) {
        float4 a = __ldg((const float4 *)&h_old[(size_t)srow * N + jq + s * P_SLICE + ch * P_UPT]);
        dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w;
      } else {
#pragma unroll
        for (int q = 0; q < P_UPT; q++) dst[q] = 0.f;
      }
    };
    if (TC2_H_L2PF && kGru && live)   // the row's upc floats = upc / 32 lines of 128 bytes, one per column group of the warp quartet
      for (int l = ch; l * 32 < upc; l += P_EPI_WARPS / 4)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(&h_old[(size_t)srow * N + jq + l * 32]) : "memory");
    load_h(0, hcur);
    if (TC2_HAHEAD == 2) load_h(1, hnext);
    if (kGru) {   // the parameter records requested in the prologue: own copies landed, then visible to all epilogue warps
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("bar.sync 1, %0;" ::"n"(32 * P_EPI_WARPS) : "memory");
    }
    for (int s = 0; s < nslice; s++) {
      const int ts = s & 1;
      if (TC2_HAHEAD == 2) load_h(s + 2, hnext2); else load_h(s + 1, hnext);
      mbar_wait(bar_tfull(ts), (uint32_t)((s >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t t0 = trow + ts * C::kCols + ch * P_UPT;
      const int ub = s * P_SLICE + ch * P_UPT;     // unit index inside this CTA's quarter
      float outv[P_UPT];
      if (kGru) {
        int az[P_UPT], ar[P_UPT], an[P_UPT], bz[P_UPT], br[P_UPT], bn[P_UPT];
        tmem_ld4(t0 + 0 * P_SLICE, az); tmem_ld4(t0 + 1 * P_SLICE, ar); tmem_ld4(t0 + 2 * P_SLICE, an);
        tmem_ld4(t0 + C::kN + 0 * P_SLICE, bz); tmem_ld4(t0 + C::kN + 1 * P_SLICE, br); tmem_ld4(t0 + C::kN + 2 * P_SLICE, bn);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // accumulators are in registers: hand the TMEM stage back to the MMA issuer
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty(ts));
        if (silent) {
#pragma unroll
          for (int q = 0; q < P_UPT; q++) outv[q] = hcur[q];
        } else {
          float zi[P_UPT], ri[P_UPT], ni[P_UPT], zr[P_UPT], rr[P_UPT], nr[P_UPT];
#pragma unroll
          for (int q = 0; q < P_UPT; q++) {
            const int u = ub + q;
            const float h = hcur[q];
            const float4 pz = *(const float4 *)&prm[16 * u], pr = *(const float4 *)&prm[16 * u + 4];
            const float4 pn = *(const float4 *)&prm[16 * u + 8], pd = *(const float4 *)&prm[16 * u + 12];
            zi[q] = (float)az[q] * pz.x + pz.y;
            ri[q] = (float)ar[q] * pr.x + pr.y;
            ni[q] = (float)an[q] * pn.x + pn.y;
            zr[q] = fmaf(pd.x, h, (float)bz[q] * pz.z + pz.w);
            rr[q] = fmaf(pd.y, h, (float)br[q] * pr.z + pr.w);
            nr[q] = fmaf(pd.z, h, (float)bn[q] * pn.z + pn.w);
          }
          gru_units<P_UPT>(zi, ri, ni, zr, rr, nr, hcur, outv);
        }
      } else {
        int acc[P_UPT];
        tmem_ld4(t0, acc);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty(ts));
#pragma unroll
        for (int q = 0; q < P_UPT; q++) outv[q] = (float)acc[q] * prm[ub + q] + prm[upc + ub + q];
          if (fabsf(outv[0]) < ACT_FAST_LIMIT && fabsf(outv[1]) < ACT_FAST_LIMIT && fabsf(outv[2]) < ACT_FAST_LIMIT && fabsf(outv[3]) < ACT_FAST_LIMIT) {
#pragma unroll
            for (int q = 0; q < P_UPT; q++) outv[q] = act_tanh_inrange(outv[q]);
          } else {
#pragma unroll
            for (int q = 0; q < P_UPT; q++) outv[q] = act_tanh(outv[q]);
          }
      }
      if (live) {
        *(float4 *)&out_f32[(size_t)srow * N + jq + ub] = make_float4(outv[0], outv[1], outv[2], outv[3]);
        *(uint32_t *)&out_u8[(size_t)srow * ldo + jq + ub] = quant4(outv[0], outv[1], outv[2], outv[3]);
      }
#pragma unroll
      for (int q = 0; q < P_UPT; q++) { hcur[q] = hnext[q]; if (TC2_HAHEAD == 2) hnext[q] = hnext2[q]; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(P_TMEM_COLS) : "memory");
  }
}
