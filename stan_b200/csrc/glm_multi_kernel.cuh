// glm_multi_kernel.cuh -- a FEW chains (four per pass: Stan's default chain count; batches of 5-8 lanes take two passes)
// in one pass over X, on the FMA path.
//
// The single-chain gradient is HBM-bound with the fp64 pipe at 15 %: the same panel stream can feed the
// arithmetic of four chains.  The DMMA path of glm_batched_kernel.cuh is the wrong tool here -- its chain blocks are 16
// wide, so 4 lanes pay for 16, and at cfg2's shape (N = 10M, K = 100) one batched leapfrog of 4 lanes took 4.95 ms against
// 1.1 ms of HBM time.  This kernel is the narrow single-chain kernel's structure (glm_kernels.cuh: persistent grid, TMA
// ring of the same swizzled 32-row panels, a stage owned by one consumer warp for the whole launch, private accumulators
// reduced once at the end) with NCH chains per row in registers:
//   phase 1  lane = row:   eta[c] = sum_k X[r,k] beta[k,c]   -- X[r,k] read from shared memory ONCE, NCH FMAs
//   link     lane = row:   link<FAMILY>() once per (row, chain)
//   phase 2  lane = (column group, row group):  acc[s][c] += X[r,k] r[r,c]   -- again one read of X for NCH FMAs
// so the shared-memory traffic per panel is that of ONE chain and the fp64 work that of NCH: ~60 % of the fp64 pipe at
// NCH = 4, still under the HBM time of the panel.  It plugs into the batched path (batched_begin -> THIS -> batched_reduce
// -> batched_finish) by writing the slice partials in the batched layout ([slice][K + 2][64], columns [0, NCH) used), so
// the per-chain epilogue, the leapfrog tail and every caller (b200glm_leapfrog_batched, the device-side NUTS rounds)
// are unchanged.  Per-chain results do not depend on which lane a chain is evaluated in; they differ from the DMMA
// kernel's in summation order only.
// Eight warps, no producer warp: a ring stage belongs to ONE warp for the whole launch, so the warp that has just consumed
// a stage is the one that knows it is free -- its lane 0 requests the stage's next panel itself (no empty-barriers).  Eight
// warps are two per SM sub-partition, which leaves each thread the 255-register budget (a ninth warp caps it at 168:
// 16 K registers per sub-partition / 3 warps; the first version of this kernel spilled there and ran latency-bound at
// 2.8 TB/s -- ncu: 14 % warps active, short-scoreboard / wait stalls, profiles/r2_glm_multi_*).
#pragma once

#include "glm_batched_kernel.cuh"

namespace b200glm {

constexpr int MULTI_THREADS = NUM_CONSUMER_WARPS * 32;
constexpr int MULTI_MAX_LANES = 8;   // batches of up to 8 lanes: one or two passes of four chains
constexpr int MULTI_MAX_K = 128;   // CPL <= 16 column slots x 4 chains = 64 fp64 accumulators per lane

// columns per ring stage: the panel's C columns, padded so that phase 2 may read all 8 CPL column slots of every lane
// without a guard (what lies beyond column C is never stored)
__host__ __device__ inline int multi_stage_cols(int C, int cpl) { return C > 8 * cpl ? C : 8 * cpl; }

__host__ __device__ inline size_t multi_smem_bytes(int K, int C, int S, int NCH, int cpl) {
  const int Kpad = (K + 7) & ~7;
  // the warps' partial rows (8 x (Kpad + 2) x NCH doubles) reuse the ring once every panel has been consumed
  return ((size_t)S * multi_stage_cols(C, cpl) * 32 + (size_t)Kpad * NCH + 2 * NCH + (size_t)NUM_CONSUMER_WARPS * 32 * NCH) * 8
         + (size_t)S * 8;
}

template <int FAMILY, int CPL, int NCH>
__global__ void __launch_bounds__(MULTI_THREADS, 1) glm_multi_kernel(const BatchedParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int K = p.K, C = p.C, S = p.n_stages, P = p.P;
  const int Kpad = (K + 7) & ~7;
  const int tile_doubles = C * 32;                                        // what a bulk copy brings
  const int stage_doubles = multi_stage_cols(C, CPL) * 32;                // ring stage stride
  double* tiles = reinterpret_cast<double*>(smem_raw);                    // S * stage_doubles
  double* sbeta = tiles + (size_t)S * stage_doubles;                      // [Kpad][NCH]
  double* salpha = sbeta + (size_t)Kpad * NCH;                            // [NCH]
  double* sisig = salpha + NCH;                                           // [NCH]
  double* sr = sisig + NCH;                                               // [8 warps][32 rows][NCH]
  double* red = tiles;                                                    // [8 warps][Kpad + 2][NCH], after the last panel
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sr + (size_t)NUM_CONSUMER_WARPS * 32 * NCH);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grid = gridDim.x;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full_bar[s], 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  // the chains' coefficients, chain-minor: one broadcast 16-byte load per chain pair in phase 1
  const int lane0 = p.lane0;   // this pass evaluates lanes [lane0, lane0 + NCH) of the batch
  for (int j = tid; j < Kpad * NCH; j += MULTI_THREADS) {
    const int k = j / NCH, c = j - k * NCH;
    sbeta[j] = (k < K && lane0 + c < p.n_lanes) ? p.theta_c[(size_t)(p.off_beta + k) * p.ldc + lane0 + c] : 0.0;
  }
  if (tid < NCH) {
    const bool act = lane0 + tid < p.n_lanes;
    salpha[tid] = act ? p.theta_c[lane0 + tid] : 0.0;
    sisig[tid] = (FAMILY == FAM_NORMAL_ID && act) ? 1.0 / exp(p.theta_c[(size_t)(P - 1) * p.ldc + lane0 + tid]) : 1.0;
  }
  __syncthreads();

  const long long n_panels = p.n_panels;
  const long long p_count = n_panels > blockIdx.x ? (n_panels - blockIdx.x + grid - 1) / grid : 0;
  const uint64_t pol = policy_evict_first();
  const uint32_t tile_bytes = (uint32_t)tile_doubles * 8u;
  // panel n of the CTA's sequence lives in stage n mod S, and stage s belongs to warp s mod 8
  auto request = [&](long long n, int s) {
    const long long pi = blockIdx.x + n * grid;
    mbar_arrive_expect_tx(&full_bar[s], tile_bytes);
    tma_load_1d(tiles + (size_t)s * stage_doubles, p.panels + (size_t)pi * tile_doubles, tile_bytes, &full_bar[s], pol);
  };

  double acc[CPL][NCH];
#pragma unroll
  for (int s2 = 0; s2 < CPL; ++s2)
#pragma unroll
    for (int c = 0; c < NCH; ++c) acc[s2][c] = 0.0;
  double lp_acc[NCH], r_acc[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) lp_acc[c] = r_acc[c] = 0.0;

  if (warp < S) {
    if (lane == 0)
      for (int s = warp; s < S; s += NUM_CONSUMER_WARPS)
        if (s < p_count) request(s, s);                    // first fill of this warp's stages
    const int rg = lane & 3, cg = lane >> 2, cgl = cg & 3;
    const int o1 = lane ^ 4, o2 = lane ^ 8, o3 = lane ^ 12;
    const int ycol = K * 32 + (lane ^ ((K & 3) << 2));
    const int tcol = (K + 1) * 32 + (lane ^ (((K + 1) & 3) << 2));   // binomial_logit: population sizes
    int off[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) off[m] = rg + 4 * (m ^ cgl);
    double* my_sr = sr + (size_t)warp * 32 * NCH;
    uint32_t parity = 0;
    for (long long n0 = 0; n0 < p_count; n0 += S, parity ^= 1u)
      for (int s = warp; s < S; s += NUM_CONSUMER_WARPS) {
        const long long n = n0 + s;
        if (n >= p_count) break;
        const long long pi = blockIdx.x + n * grid;
        mbar_wait(&full_bar[s], parity);
        const double* tile = tiles + (size_t)s * stage_doubles;

        // ---- phase 1: eta of row `lane` for every chain; even / odd feature pairs in separate accumulators:
        //      2 NCH independent FMA chains per lane ----
        double ea[NCH], eb[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) ea[c] = eb[c] = 0.0;
        int k = 0;
#pragma unroll 2
        for (; k + 4 <= K; k += 4) {
          const double x0 = tile[(k + 0) * 32 + lane], x1 = tile[(k + 1) * 32 + o1];
          const double x2 = tile[(k + 2) * 32 + o2], x3 = tile[(k + 3) * 32 + o3];
#pragma unroll
          for (int c = 0; c < NCH; c += 2) {
            const double2 b0 = *reinterpret_cast<const double2*>(sbeta + (size_t)(k + 0) * NCH + c);
            const double2 b1 = *reinterpret_cast<const double2*>(sbeta + (size_t)(k + 1) * NCH + c);
            const double2 b2 = *reinterpret_cast<const double2*>(sbeta + (size_t)(k + 2) * NCH + c);
            const double2 b3 = *reinterpret_cast<const double2*>(sbeta + (size_t)(k + 3) * NCH + c);
            ea[c] = fma(x0, b0.x, ea[c]);
            ea[c + 1] = fma(x0, b0.y, ea[c + 1]);
            eb[c] = fma(x1, b1.x, eb[c]);
            eb[c + 1] = fma(x1, b1.y, eb[c + 1]);
            ea[c] = fma(x2, b2.x, ea[c]);
            ea[c + 1] = fma(x2, b2.y, ea[c + 1]);
            eb[c] = fma(x3, b3.x, eb[c]);
            eb[c + 1] = fma(x3, b3.y, eb[c + 1]);
          }
        }
        for (; k < K; ++k) {
          const double x = tile[k * 32 + (lane ^ ((k & 3) << 2))];
#pragma unroll
          for (int c = 0; c < NCH; ++c) ea[c] = fma(x, sbeta[(size_t)k * NCH + c], ea[c]);
        }
        const double y = tile[ycol];
        const double aux = FAMILY == FAM_BINOMIAL_LOGIT ? tile[tcol] : 0.0;
        const bool valid = (pi * PANEL_ROWS + lane) < p.n_rows;
        double rres[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          double lp_i, r_i;
          link_bf<FAMILY>((ea[c] + eb[c]) + salpha[c], y, sisig[c], lp_i, r_i, aux);   // branch-free: the chains interleave
          if (!valid) {
            lp_i = 0.0;
            r_i = 0.0;
          }
          lp_acc[c] += lp_i;
          r_acc[c] += r_i;
          rres[c] = r_i;
        }
        // ---- phase 2: X^T r from the same tile ----
#pragma unroll
        for (int c = 0; c < NCH; c += 2)
          *reinterpret_cast<double2*>(my_sr + (size_t)lane * NCH + c) = make_double2(rres[c], rres[c + 1]);
        __syncwarp();
        const double* base = tile + cg * 32;
#pragma unroll
        for (int mh = 0; mh < 8; mh += 4) {          // rows in two halves: 4 x NCH residuals in registers at a time
          double rr[4][NCH];
#pragma unroll
          for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int c = 0; c < NCH; c += 2) {
              const double2 v = *reinterpret_cast<const double2*>(my_sr + (size_t)(rg + 4 * (mh + m)) * NCH + c);
              rr[m][c] = v.x;
              rr[m][c + 1] = v.y;
            }
          // row outermost: successive FMAs into one accumulator are 4 CPL instructions apart (no dependency stalls),
          // no guard on the column (the stage is padded to 8 CPL columns; sums of columns >= K are never stored)
#pragma unroll
          for (int m = 0; m < 4; ++m) {
#pragma unroll
            for (int s2 = 0; s2 < CPL; ++s2) {
              const double x = base[s2 * 256 + off[mh + m]];
#pragma unroll
              for (int c = 0; c < NCH; ++c) acc[s2][c] = fma(x, rr[m][c], acc[s2][c]);
            }
          }
        }
        __syncwarp();                                      // every lane has read the tile and my_sr
        if (lane == 0 && n + S < p_count) request(n + S, s);   // the stage's next fill
      }
  }
  // every warp has consumed its last panel (so every bulk copy has landed): the ring becomes `red`
  __syncthreads();
  // ---- per-warp reduction of the private accumulators (idle warps contribute zeros) ----
  {
    double* my_red = red + (size_t)warp * (Kpad + 2) * NCH;
    const int rg = lane & 3, cg = lane >> 2;
#pragma unroll
    for (int s2 = 0; s2 < CPL; ++s2)
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        double v = acc[s2][c];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (rg == 0 && cg + 8 * s2 < Kpad) my_red[(size_t)(cg + 8 * s2) * NCH + c] = v;
      }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const double a = warp_sum(lp_acc[c]), b = warp_sum(r_acc[c]);
      if (lane == 0) {
        my_red[(size_t)Kpad * NCH + c] = a;
        my_red[(size_t)(Kpad + 1) * NCH + c] = b;
      }
    }
  }
  __syncthreads();

  // ---- CTA partial in the batched layout: [slice = CTA][row in [0, K + 2)][64 chain columns]; this pass owns columns
  //      [lane0, lane0 + NCH), warps folded in fixed order (the other columns belong to other passes or to no lane:
  //      the epilogue only reads the columns of live lanes) ----
  double* part = p.partials + (size_t)blockIdx.x * (K + 2) * BATCH_CB + lane0;
  for (int j = tid; j < (K + 2) * NCH; j += MULTI_THREADS) {
    const int row = j / NCH, c = j - row * NCH;
    const int src = row < K ? row : Kpad + (row - K);
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < NUM_CONSUMER_WARPS; ++w) v += red[((size_t)w * (Kpad + 2) + src) * NCH + c];
    part[(size_t)row * BATCH_CB + c] = v;
  }
}

}  // namespace b200glm
