// net_kernel.cuh -- conv1 -> conv2 -> GRU1 -> GRU2 -> GRU3 of compute_rnn() (reference src/rnn.c:44-60) as ONE persistent
// tcgen05 kernel over a 4-CTA thread-block cluster per 128-stream tile (the default network path; the per-layer
// kernels k_tc2<> of gru_tc.cuh run the same arithmetic one layer per launch and are kept as cross-checks).
//
//   cluster = 4 CTAs = one tile of 128 streams; CTA r of the cluster owns output units [r*N/4, (r+1)*N/4) of
//   EVERY layer.  Per layer a CTA needs the whole u8 activation rows of the previous layer (K = N bytes per
//   stream): each CTA writes its unit quarter (fp32 state + u8 operand mirror) to global memory, the cluster
//   synchronises (barrier.cluster release/acquire + proxy fence), and every CTA TMA-loads the full 128 x K tile
//   back from L2.  The recurrent operand Hu8 of the NEXT layer does not depend on this frame, so its TMA load is
//   issued as soon as the current layer's MMAs have retired and hides behind the epilogue tail; the weight-slice
//   ring (3 stages) and the TMEM accumulator ring (2 stages) simply run on across layer boundaries, so the weights
//   of the next layer's first slices are already in shared memory when its activations arrive.
//
//   conv1 (fp32, 195 -> cond, one sequential FMA chain per output: nnet.c:113-123, sgemv vec_avx.h:672) runs as a
//   prologue on the epilogue warps while the producer prefetches weights: CTA r computes it for streams
//   [32 r, 32 r + 32) of the tile (thread = output x 8 streams, inputs transposed in shared memory), updates the
//   conv1 memory and conv2's u8 operand rows in global memory, and the cluster barrier that starts conv2 publishes
//   them.  (k_conv1 of rnn_kernels.cuh is the stand-alone cross-check.)
//
//   warp 16 (one elected thread): TMA producer + tcgen05.mma issuer      (as in k_tc2)
//   warps 0..15                 : epilogue, warp w -> TMEM lane quarter w & 3, units 4 * (w >> 2) .. + 4 of a slice
//
// Against one launch per layer this removes four kernel boundaries (drain, launch latency, barrier/TMEM/parameter
// prologue, cold TMA pipeline) per frame and lane.  Arithmetic: identical to k_tc2 / the dp4a kernels, bit for bit
// (exact s32 accumulators; (float)acc*scale + subias; fma(diag,h,.); Pade sigmoid/tanh; h' = z*h + (1-z)*n).
// grid = (ceil(S/128), R), cluster (1,R,1) with R = 4 or 8, block = 544, dynamic smem = net_smem_bytes(), 1 CTA / SM.
#pragma once
#include "gru_tc.cuh"

#define NET_LAYERS 4   // conv2 + 3 GRU

struct NetMaps {       // TMA descriptors of one frame parity
  CUtensorMap x[NET_LAYERS], h[NET_LAYERS], wi[NET_LAYERS], wr[NET_LAYERS];   // h / wr unused for layer 0 (conv2)
};
struct NetPtrs {
  const float *scale_i[NET_LAYERS], *subias_i[NET_LAYERS];   // input matrix (conv2: the only matrix)
  const float *scale_r[NET_LAYERS], *subias_r[NET_LAYERS], *diag[NET_LAYERS];
  const float *packed[NET_LAYERS];    // GRU layers: DevLayerQ::packed, [N][16] epilogue parameter records
  const float *h_old[NET_LAYERS];     // fp32 state of the previous frame (GRU layers)
  float *out_f32[NET_LAYERS];         // conv2_out / new fp32 state
  uint8_t *out_u8[NET_LAYERS];        // their u8 operand mirrors
  // conv1 prologue (conv1_w == nullptr: conv2's operand rows were prepared by k_conv1)
  const float *conv1_w, *conv1_b;     // [195][cond], [cond]
  const float *features;              // [S][65] of this frame
  float *conv1_state;                 // [S][130]
  uint8_t *c2in;                      // [S][Kc] conv2 operand rows: [memory (2 x cond) | newest (cond) | pad]
  int cond;
};
#define NET_C1_IN (3 * NB_FEAT)       // 195

__host__ __device__ constexpr int net_stage_bytes(int K) { return 2 * (K / TC_KATOM) * (3 * P_SLICE * TC_KATOM); }
__host__ __device__ constexpr int net_prm_floats(int N) { return (2 + 3 * 16) * (N / 4); }   // sized for the smallest cluster (4)
__host__ __device__ constexpr int net_smem_bytes(int Kc, int Kn, int N) {
  // A tiles: X (max(Kc, Kn) bytes per row) + H (Kn); B ring; parameters of all layers; barriers
  return 1024 + ((Kc > Kn ? Kc : Kn) / TC_KATOM + Kn / TC_KATOM) * TC_A_ATOM_BYTES + P_STAGES * net_stage_bytes(Kn) +
         net_prm_floats(N) * 4 + 24 * 8 + 64;
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Kc = conv2's contraction length (3 * cond) and Kn = the GRU layers' (gru), both padded up to multiples of 128 with
// zero weights; N = gru (a multiple of 64).  Kn is also the row stride of every u8 activation mirror.
// The cluster size R = gridDim.y (4 or 8, set by the launch attribute) splits every layer's units R ways: R = 8 halves
// each CTA's share (and the kernel's latency) at the price of twice the SMs per tile.
__global__ void __launch_bounds__(P_THREADS, 1)
k_net(int S, int Kc, int Kn, int N, const __grid_constant__ NetMaps maps, const __grid_constant__ NetPtrs p, const int *__restrict__ silence) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = (int)gridDim.y, upc = N / R, nslice = upc / P_SLICE;   // units / slices per CTA
  const int atoms_c = Kc / TC_KATOM, atoms_n = Kn / TC_KATOM, atoms_x = atoms_c > atoms_n ? atoms_c : atoms_n;
  const int m0 = blockIdx.x * TC_M, jq = blockIdx.y * upc;
  uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t *sAx = base, *sAh = sAx + atoms_x * TC_A_ATOM_BYTES;
  uint8_t *sB = sAh + atoms_n * TC_A_ATOM_BYTES;
  const int stage_bytes = net_stage_bytes(Kn);
  float *prm = (float *)(sB + P_STAGES * stage_bytes);   // conv: [2][upc]; then per GRU layer [upc][16]
  uint64_t *bars = (uint64_t *)(prm + net_prm_floats(N));
  uint32_t *tmem_slot = (uint32_t *)(bars + 24);
  const uint32_t bar_x = smem_u32(&bars[0]), bar_h = smem_u32(&bars[1]), bar_adone = smem_u32(&bars[2]);
  auto bar_bfull = [&](int i) { return smem_u32(&bars[4 + i]); };
  auto bar_bempty = [&](int i) { return smem_u32(&bars[8 + i]); };
  auto bar_tfull = [&](int i) { return smem_u32(&bars[12 + i]); };
  auto bar_tempty = [&](int i) { return smem_u32(&bars[14 + i]); };

  if (tid == 0) {
    mbar_init(bar_x, 1); mbar_init(bar_h, 1); mbar_init(bar_adone, 1);
    for (int i = 0; i < P_STAGES; i++) { mbar_init(bar_bfull(i), 1); mbar_init(bar_bempty(i), 1); }
    for (int i = 0; i < 2; i++) { mbar_init(bar_tfull(i), 1); mbar_init(bar_tempty(i), P_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(P_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // epilogue parameters of this CTA's unit quarter, all layers
  for (int i = tid; i < 2 * upc; i += blockDim.x) prm[i] = (i < upc ? p.scale_i[0] : p.subias_i[0])[jq + i % upc];
  // GRU layers: this CTA's unit slice of the packed parameter records, contiguous in memory and in shared memory:
  // 16-byte asynchronous copies, all in flight together (six dependent scattered loads per thread before)
  for (int L = 1; L < NET_LAYERS; L++) {
    float *pl = prm + 2 * upc + (L - 1) * 16 * upc;
    const float *src = p.packed[L] + (size_t)jq * 16;
    for (int c = tid; c < 4 * upc; c += blockDim.x) cp_async16(pl + 4 * c, src + 4 * c, true);
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (p.conv1_w && warp < P_EPI_WARPS) {
    // ---- conv1 prologue for streams [m0 + 32 r, + 32), r = this CTA's rank; the X tile region is still free ----
    float *tmpT = (float *)sAx;                      // [195][32]: input j of the 32 streams (conflict-free, LDS.128 broadcast)
    const int cond = p.cond, W = cond / 4, nrow = TC_M / R, r0 = m0 + nrow * (int)blockIdx.y;   // this CTA's 32 (R = 4) or 16 streams
    {
      // 32 x 195 inputs, 13 per thread (idx = tid + 512 u): all global loads in flight together, then the stores
      float v[13];
#pragma unroll
      for (int u = 0; u < 13; u++) {
        const int idx = tid + u * 32 * P_EPI_WARPS, sl = idx & 31, j = idx >> 5, row = r0 + sl;
        v[u] = 0.f;
        if (idx < 32 * NET_C1_IN && sl < nrow && row < S)
          v[u] = j < 2 * NB_FEAT ? p.conv1_state[(size_t)row * 2 * NB_FEAT + j] : p.features[(size_t)row * NB_FEAT + j - 2 * NB_FEAT];
      }
#pragma unroll
      for (int u = 0; u < 13; u++) {
        const int idx = tid + u * 32 * P_EPI_WARPS;
        if (idx < 32 * NET_C1_IN) tmpT[idx] = v[u];
      }
    }
    // the words of the operand rows that the memory update moves down (read everything before anything is written)
    uint32_t rot[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int idx = tid + k * 32 * P_EPI_WARPS, sl = idx / (2 * W), w = idx - sl * 2 * W, row = r0 + sl;
      rot[k] = (sl < nrow && row < S) ? ((const uint32_t *)(p.c2in + (size_t)row * Kc))[W + w] : 0u;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * P_EPI_WARPS) : "memory");
    const int o = tid & 127, sg = tid >> 7;          // output, group of 8 streams
    if (o < cond && sg * 8 < nrow) {
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; i++) acc[i] = 0.f;
      const float *wp = p.conv1_w + o;
      // sequential FMA chain over the 195 inputs (sgemv order); the weights come from L2 / L1 in blocks of 13
      // independent loads, the next block requested before the current one is consumed (195 = 15 x 13)
      float wn[13];
#pragma unroll
      for (int u = 0; u < 13; u++) wn[u] = __ldg(wp + (size_t)u * cond);
      for (int j0 = 0; j0 < NET_C1_IN; j0 += 13) {
        float wc[13];
#pragma unroll
        for (int u = 0; u < 13; u++) wc[u] = wn[u];
        if (j0 + 13 < NET_C1_IN) {
#pragma unroll
          for (int u = 0; u < 13; u++) wn[u] = __ldg(wp + (size_t)(j0 + 13 + u) * cond);
        }
#pragma unroll
        for (int u = 0; u < 13; u++) {
          const int j = j0 + u;
          const float w = wc[u];
          const float4 a = *(const float4 *)&tmpT[j * 32 + sg * 8], b = *(const float4 *)&tmpT[j * 32 + sg * 8 + 4];
          acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]); acc[2] = fmaf(w, a.z, acc[2]); acc[3] = fmaf(w, a.w, acc[3]);
          acc[4] = fmaf(w, b.x, acc[4]); acc[5] = fmaf(w, b.y, acc[5]); acc[6] = fmaf(w, b.z, acc[6]); acc[7] = fmaf(w, b.w, acc[7]);
        }
      }
      const float bias = p.conv1_b[o];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int row = r0 + sg * 8 + i;
        if (row < S && !silence[row]) p.c2in[(size_t)row * Kc + 2 * cond + o] = (uint8_t)quant_u8(act_tanh(acc[i] + bias));
      }
    }
    // memory updates; silent frames leave both memories untouched (denoise.c:474)
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int idx = tid + k * 32 * P_EPI_WARPS, sl = idx / (2 * W), w = idx - sl * 2 * W, row = r0 + sl;
      if (sl < nrow && row < S && !silence[row]) ((uint32_t *)(p.c2in + (size_t)row * Kc))[w] = rot[k];
    }
    for (int idx = tid; idx < 32 * 2 * NB_FEAT; idx += 32 * P_EPI_WARPS) {
      const int sl = idx / (2 * NB_FEAT), j = idx - sl * 2 * NB_FEAT, row = r0 + sl;
      if (sl < nrow && row < S && !silence[row]) p.conv1_state[(size_t)row * 2 * NB_FEAT + j] = tmpT[(NB_FEAT + j) * 32 + sl];
    }
    fence_proxy_async();   // the operand rows are read back by TMA after the cluster barrier
  }

  if (warp == P_EPI_WARPS) {
    // ------------------------------------------------------------------------------------------------
    // producer / MMA issuer: lane 0 works, the whole warp takes part in the cluster barriers
    // ------------------------------------------------------------------------------------------------
    int job = 0;   // global slice index over all layers: weight stage job % P_STAGES, TMEM stage job & 1
    auto load_B = [&](int j) {
      const int L = j / nslice, s = j - L * nslice, st = j % P_STAGES;
      uint8_t *dst = sB + st * stage_bytes;
      if (L == 0) {
        const int batom = P_SLICE * TC_KATOM;
        mbar_expect_tx(bar_bfull(st), (uint32_t)(atoms_c * batom));
        for (int a = 0; a < atoms_c; a++)
          tma_load_2d(smem_u32(dst + a * batom), &maps.wi[0], bar_bfull(st), a * TC_KATOM, jq + s * P_SLICE);
      } else {
        const int batom = 3 * P_SLICE * TC_KATOM, row = (blockIdx.y * nslice + s) * 3 * P_SLICE;
        mbar_expect_tx(bar_bfull(st), (uint32_t)(2 * atoms_n * batom));
        for (int a = 0; a < atoms_n; a++) {
          tma_load_2d(smem_u32(dst + a * batom), &maps.wi[L], bar_bfull(st), a * TC_KATOM, row);
          tma_load_2d(smem_u32(dst + (atoms_n + a) * batom), &maps.wr[L], bar_bfull(st), a * TC_KATOM, row);
        }
      }
    };
    const int njobs = NET_LAYERS * nslice;
    if (lane == 0) {
      // weights and GRU1's recurrent operand (previous frame's state) do not depend on conv1: fetched beside it
      for (int j = 0; j < P_STAGES && j < njobs; j++) load_B(j);
      mbar_expect_tx(bar_h, (uint32_t)(atoms_n * TC_A_ATOM_BYTES));
      for (int a = 0; a < atoms_n; a++) tma_load_2d(smem_u32(sAh + a * TC_A_ATOM_BYTES), &maps.h[1], bar_h, a * TC_KATOM, m0);
    }
    __syncwarp();
    cluster_sync_all();   // every CTA is up (barriers initialised) and the conv1 prologues of the whole tile are in global memory
    if (lane == 0) {
      fence_proxy_async();
      mbar_expect_tx(bar_x, (uint32_t)(atoms_c * TC_A_ATOM_BYTES));
      for (int a = 0; a < atoms_c; a++) tma_load_2d(smem_u32(sAx + a * TC_A_ATOM_BYTES), &maps.x[0], bar_x, a * TC_KATOM, m0);
    }
    for (int L = 0; L < NET_LAYERS; L++) {
      if (lane == 0) {
        const int natoms = L == 0 ? atoms_c : atoms_n, nmat = L == 0 ? 1 : 2;
        const int kN = L == 0 ? P_SLICE : 3 * P_SLICE, batom = kN * TC_KATOM;
        const uint32_t idesc = umma_idesc_i8(TC_M, kN);
        mbar_wait(bar_x, (uint32_t)(L & 1));
        if (L >= 1) mbar_wait(bar_h, (uint32_t)((L - 1) & 1));
        for (int s = 0; s < nslice; s++, job++) {
          const int st = job % P_STAGES, ts = job & 1;
          mbar_wait(bar_bfull(st), (uint32_t)((job / P_STAGES) & 1));
          mbar_wait(bar_tempty(ts), (uint32_t)(((job >> 1) & 1) ^ 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint8_t *Bs = sB + st * stage_bytes;
          for (int g = 0; g < nmat; g++) {
            const uint8_t *A = g ? sAh : sAx;
            for (int a = 0; a < natoms; a++) {
              const uint64_t ad = umma_desc_sw128(smem_u32(A + a * TC_A_ATOM_BYTES));
              const uint64_t bd = umma_desc_sw128(smem_u32(Bs + (g * natoms + a) * batom));
#pragma unroll
              for (int k = 0; k < TC_KATOM / 32; k++)
                umma_i8(tmem + ts * (6 * P_SLICE) + g * kN, ad + (uint64_t)(k * 32 >> 4), bd + (uint64_t)(k * 32 >> 4), idesc, (a | k) ? 1u : 0u);
            }
          }
          umma_commit(bar_bempty(st));
          umma_commit(bar_tfull(ts));
          if (job >= 1 && job - 1 + P_STAGES < njobs) {   // refill the stage the previous job used
            mbar_wait(bar_bempty((job - 1) % P_STAGES), (uint32_t)(((job - 1) / P_STAGES) & 1));
            load_B(job - 1 + P_STAGES);
          }
        }
        if (L + 1 < NET_LAYERS) {
          // the activation tiles are reusable once this layer's MMAs have retired: fetch the next layer's recurrent
          // operand right away (it only depends on the previous frame)
          umma_commit(bar_adone);
          mbar_wait(bar_adone, (uint32_t)(L & 1));
          if (L + 1 >= 2) {   // (GRU1's was loaded in the prologue: conv2 does not use the tile)
            mbar_expect_tx(bar_h, (uint32_t)(atoms_n * TC_A_ATOM_BYTES));
            for (int a = 0; a < atoms_n; a++) tma_load_2d(smem_u32(sAh + a * TC_A_ATOM_BYTES), &maps.h[L + 1], bar_h, a * TC_KATOM, m0);
          }
        }
      }
      if (L + 1 < NET_LAYERS) {
        __syncwarp();
        cluster_sync_all();   // all four unit quarters of layer L are in global memory
        if (lane == 0) {
          fence_proxy_async();
          mbar_expect_tx(bar_x, (uint32_t)(atoms_n * TC_A_ATOM_BYTES));
          for (int a = 0; a < atoms_n; a++) tma_load_2d(smem_u32(sAx + a * TC_A_ATOM_BYTES), &maps.x[L + 1], bar_x, a * TC_KATOM, m0);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------------
    // epilogue warps
    // ------------------------------------------------------------------------------------------------
    cluster_sync_all();   // (pairs with the producer warp's: conv1 of the whole tile is published)
    const int lq = warp & 3, ch = warp >> 2;
    const int srow = m0 + lq * 32 + lane;
    const bool live = srow < S;
    const bool silent = live ? silence[srow] != 0 : true;
    const uint32_t trow = tmem + ((uint32_t)(lq * 32) << 16);
    int job = 0;
    for (int L = 0; L < NET_LAYERS; L++) {
      const float *h_old = p.h_old[L];
      float *out_f32 = p.out_f32[L];
      uint8_t *out_u8 = p.out_u8[L];
      const float *pl = prm + (L == 0 ? 0 : 2 * upc + (L - 1) * 16 * upc);
      float hcur[P_UPT], hnext[P_UPT];
      auto load_h = [&](int s, float (&dst)[P_UPT]) {
        if (L > 0 && live && s < nslice) {
          float4 a = __ldg((const float4 *)&h_old[(size_t)srow * N + jq + s * P_SLICE + ch * P_UPT]);
          dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w;
        } else {
#pragma unroll
          for (int q = 0; q < P_UPT; q++) dst[q] = 0.f;
        }
      };
      load_h(0, hcur);
      for (int s = 0; s < nslice; s++, job++) {
        const int ts = job & 1;
        load_h(s + 1, hnext);
        mbar_wait(bar_tfull(ts), (uint32_t)((job >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t t0 = trow + ts * (6 * P_SLICE) + ch * P_UPT;
        const int ub = s * P_SLICE + ch * P_UPT;     // unit index inside this CTA's quarter
        float outv[P_UPT];
        if (L > 0) {
          int az[P_UPT], ar[P_UPT], an[P_UPT], bz[P_UPT], br[P_UPT], bn[P_UPT];
          tmem_ld4(t0 + 0 * P_SLICE, az); tmem_ld4(t0 + 1 * P_SLICE, ar); tmem_ld4(t0 + 2 * P_SLICE, an);
          tmem_ld4(t0 + 3 * P_SLICE, bz); tmem_ld4(t0 + 4 * P_SLICE, br); tmem_ld4(t0 + 5 * P_SLICE, bn);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
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
              const float4 pz = *(const float4 *)&pl[16 * u], pr = *(const float4 *)&pl[16 * u + 4];
              const float4 pn = *(const float4 *)&pl[16 * u + 8], pd = *(const float4 *)&pl[16 * u + 12];
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
          for (int q = 0; q < P_UPT; q++) outv[q] = (float)acc[q] * pl[ub + q] + pl[upc + ub + q];
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
          *(uint32_t *)&out_u8[(size_t)srow * Kn + jq + ub] = quant4(outv[0], outv[1], outv[2], outv[3]);
        }
#pragma unroll
        for (int q = 0; q < P_UPT; q++) hcur[q] = hnext[q];
      }
      if (L + 1 < NET_LAYERS) {
        // this thread's part of layer L is written: make it visible to the TMA (async proxy) reads of the whole
        // cluster, then wait until every CTA has done the same
        fence_proxy_async();
        cluster_sync_all();
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(P_TMEM_COLS) : "memory");
  }
  cluster_sync_all();   // no CTA of the cluster exits while a peer could still be arriving on the cluster barrier
}
