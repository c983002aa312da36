// glm_wide_kernel.cuh -- single-pass GLM log-density + gradient for WIDE design matrices
// (K > 256, BASELINE configs[4]: K = 1000), where one 32-row panel no longer fits a warp's
// registers / a CTA's shared memory.  Same arithmetic and same tail as glm_fused_kernel
// (glm_kernels.cuh); what changes is how a row panel is spread over the CTA.
//
// Data layout in HBM ("wide row-panel format", relayout_kernel with PR = WR, no swizzle):
//   rows are cut into panels of WR = 16, 8 or 4 rows (wide_rows_for: the panel is kept near 64 KB);
//   panel n is one contiguous block of Cpad columns x WR doubles (column-major inside the panel,
//   Cpad = C rounded up to the columns a warp covers per step, padding columns are zero).
//   A panel is streamed as J <= 16 sub-panels of KC columns (one cp.async.bulk of 4-8 KB each).
//   Why ~64 KB panels: with 16 rows at K = 1000 (128 KB resident, 80 KB look-ahead) loads could only be
//   issued in bursts while P2 freed slots and HBM idled one latency per panel (ncu: 57 % DRAM, 71 % of
//   the copy roof); with the ring holding >= 2 panels the stream is continuous (89 %).
//
// CTA (persistent, one per SM) = 8 consumer warps + 1 TMA producer warp + 1 link warp.
//   The shared-memory ring has T >= 2J sub-panel slots: the panel between its eta pass and its X^T r
//   pass stays resident, the rest holds the following panel(s).
//   consumer warp w owns sub-panels j = w, w+8, ... of every row panel (so it owns those columns'
//   gradient accumulators outright: no cross-warp reduction of the gradient):
//     P1(n, j): partial eta for the panel's rows over the sub-panel's columns
//     -> 8 partials/row meet in smem, ETA barrier -> link warp adds them in fixed order, applies
//        the link function (lp_i, r_i), R barrier ->
//     P2(n, j): acc[col] += X[row][col] * r[row] from the SAME smem bytes, then the slot is released.
//   Software pipeline: a consumer publishes eta(n+1) BEFORE it waits for r(n), so the link warp's
//   fp64 exp/log1p dependency chain for panel n+1 overlaps P2(n); eta / r buffers and the two named
//   barriers are double-buffered by panel parity.
//   Round 2 measured four variants of the consumer / link pipeline -- two link warps (even / odd panels), 4-row
//   panels, the panel kept in registers between the passes with the ring slot released after P1, two link warps with a
//   look-ahead of two panels -- and K = 1000 stayed within 1 % of 1.347 ms in all of them
//   (profiles/r2_wide_variants.txt): the limiter was the single-lane TMA issue loop (see the producer below), not
//   the pipeline; the consumer / link structure is round 1's.
//   Lane mapping (both passes): lane = (rq, cq) handles column cq + CPS*t and four rows
//   {2rq, 2rq+1, 2(rq+LPC), 2(rq+LPC)+1} with two LDS.128; half of the columns swap the order of the
//   two loads, which makes every quarter-warp cover all 32 banks (conflict free without a swizzle).
#pragma once

#include "glm_kernels.cuh"

namespace b200glm {

constexpr int WIDE_CONSUMER_WARPS = 8;
constexpr int WIDE_THREADS = (WIDE_CONSUMER_WARPS + 2) * 32;
constexpr int WIDE_MAX_SLOTS = 96;
// named barriers: ETA / R, double-buffered by panel parity
enum { WIDE_BAR_ETA = 1, WIDE_BAR_R = 3 };
constexpr int WIDE_BAR_COUNT = (WIDE_CONSUMER_WARPS + 1) * 32;  // consumers + link warp

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// rows per panel: the panel (WR x Cpad doubles) should be about 64 KB so that the ring holds the
// resident panel plus at least one more in flight, and long enough (>= 1 us of HBM time) for the link
// warp's fp64 exp/log1p chain to stay off the critical path
__host__ __device__ inline int wide_rows_for(int C) { return C <= 512 ? 16 : (C <= 1024 ? 8 : 4); }
__host__ __device__ inline int wide_cps(int WR) { return 32 / (WR / 4); }

// doubles of dynamic shared memory besides the ring slots and their barriers
__host__ __device__ inline size_t wide_fixed_doubles(int WR, int J, int KC, int G, int stage_a, int P_state = 0) {
  return (size_t)J * KC + 2 * WIDE_CONSUMER_WARPS * WR + 2 * WR + (stage_a ? ((G + 1) & ~1) : 0)
         + (P_state ? state_smem_doubles(P_state) : 0);
}

template <int FAMILY, int WR, int SPW, int SPC>
__global__ void __launch_bounds__(WIDE_THREADS, 1) glm_wide_kernel(const __grid_constant__ KernelParams p) {
  constexpr int LPC = WR / 4, CPS = 32 / LPC;                  // lanes per column, columns per warp step
  constexpr int KC = SPC * CPS;                                // sub-panel width (columns)
  constexpr int SLOT = KC * WR;                                // doubles per ring slot
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int K = p.K, G = p.G, P = p.P, J = p.J, T = p.n_stages, Cpad = p.Cpad;

  double* ring = reinterpret_cast<double*>(smem_raw);            // T * SLOT
  double* sbeta = ring + (size_t)T * SLOT;                       // J * KC (zero beyond K)
  double* eta_part = sbeta + J * KC;                             // 2 parities x 8 warps x WR
  double* r_sh = eta_part + 2 * WIDE_CONSUMER_WARPS * WR;        // 2 parities x WR
  double* sa = r_sh + 2 * WR;                                    // G (optional)
  double* st_base = sa + (p.stage_a_in_smem ? ((G + 1) & ~1) : 0);   // chain state (optional)
  const StateSmem st = carve_state_smem(st_base, P, p.state_in_smem);
  double* after_a = st_base + (p.state_in_smem ? state_smem_doubles(P) : 0);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(after_a);     // T
  uint64_t* empty_bar = full_bar + T;                            // T
  __shared__ double sh_scratch[64];
  __shared__ int sh_is_last;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grid = gridDim.x;

  pdl_launch_dependents();
  if (tid == 0) {
    for (int s = 0; s < T; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }

  // ---- theta for this launch (leapfrog: begin_update_p + update_q, expl_leapfrog.hpp:16-26) ----
  auto theta_at = [&](int i) -> double {
    if (p.mode == MODE_LEAPFROG) {
      const double ph = p.st_in[P + i] - (0.5 * p.eps) * p.st_in[2 * P + i];
      return p.st_in[i] + p.eps * (p.inv_metric[i] * ph);
    }
    return p.theta_inline_n ? p.theta_inline[i] : p.theta_in[i];
  };
  pdl_grid_dependency_wait();   // theta / the leapfrog state come from the previous launch on this stream
  stage_theta(p, tid, WIDE_THREADS, sbeta, J * KC, p.stage_a_in_smem ? sa : nullptr, st);
  __syncthreads();

  const long long n_panels = p.n_panels;
  const long long n_my = (long long)blockIdx.x < n_panels ? (n_panels - blockIdx.x + grid - 1) / grid : 0;
  double* my_part = p.partials + (size_t)blockIdx.x * p.pstride;

  if (warp < WIDE_CONSUMER_WARPS) {
    // =============================== consumers ===============================
    const int rq = lane & (LPC - 1), cq = lane / LPC;
    const int swap = (cq / (4 / LPC)) & 1;  // which half of the 128-byte bank window the first load takes
    const int chunkA = swap ? rq + LPC : rq, chunkB = swap ? rq : rq + LPC;
    const int offA = 2 * chunkA, offB = 2 * chunkB;

    int slot[SPW];      // ring slot / phase parity of this warp's sub-panels of the panel P2 works on
    uint32_t par[SPW];
    int steps[SPW];     // 0 = this warp has no such sub-panel
#pragma unroll
    for (int i = 0; i < SPW; ++i) {
      const int j = warp + WIDE_CONSUMER_WARPS * i;
      slot[i] = j;
      par[i] = 0;
      const int cols = j < J ? min(KC, Cpad - j * KC) : 0;
      steps[i] = cols / CPS;
    }
    double acc[SPW * SPC];
#pragma unroll
    for (int a = 0; a < SPW * SPC; ++a) acc[a] = 0.0;

    // P1 of one whole panel (this warp's sub-panels, at ring position `ahead` panels past slot[]),
    // then publish the warp's partial eta of the panel's rows into parity buffer `buf`
    auto pass1_publish = [&](int ahead, int buf) {
      double eA0 = 0.0, eA1 = 0.0, eB0 = 0.0, eB1 = 0.0;
#pragma unroll
      for (int i = 0; i < SPW; ++i) {
        if (steps[i] > 0) {
          int sl = slot[i];
          uint32_t pr = par[i];
          if (ahead) {
            sl += J;
            if (sl >= T) {
              sl -= T;
              pr ^= 1u;
            }
          }
          mbar_wait(&full_bar[sl], pr);
          const double* tile = ring + (size_t)sl * SLOT + cq * WR;
          const double* bj = sbeta + (warp + WIDE_CONSUMER_WARPS * i) * KC + cq;
#pragma unroll
          for (int t = 0; t < SPC; ++t) {
            if (t < steps[i]) {
              const double b = bj[CPS * t];
              const double2 xa = *reinterpret_cast<const double2*>(tile + CPS * WR * t + offA);
              const double2 xb = *reinterpret_cast<const double2*>(tile + CPS * WR * t + offB);
              eA0 = fma(xa.x, b, eA0);
              eA1 = fma(xa.y, b, eA1);
              eB0 = fma(xb.x, b, eB0);
              eB1 = fma(xb.y, b, eB1);
            }
          }
        }
      }
      double lo0 = swap ? eB0 : eA0, lo1 = swap ? eB1 : eA1;  // rows 2rq, 2rq+1
      double hi0 = swap ? eA0 : eB0, hi1 = swap ? eA1 : eB1;  // rows 2(rq+LPC), 2(rq+LPC)+1
#pragma unroll
      for (int o = LPC; o < 32; o <<= 1) {
        lo0 += __shfl_xor_sync(0xffffffffu, lo0, o);
        lo1 += __shfl_xor_sync(0xffffffffu, lo1, o);
        hi0 += __shfl_xor_sync(0xffffffffu, hi0, o);
        hi1 += __shfl_xor_sync(0xffffffffu, hi1, o);
      }
      if (cq == 0) {
        double* ep = eta_part + (buf * WIDE_CONSUMER_WARPS + warp) * WR;
        *reinterpret_cast<double2*>(ep + 2 * rq) = make_double2(lo0, lo1);
        *reinterpret_cast<double2*>(ep + 2 * (rq + LPC)) = make_double2(hi0, hi1);
      }
      __threadfence_block();
      named_bar_arrive(WIDE_BAR_ETA + buf, WIDE_BAR_COUNT);
    };

    if (n_my > 0) pass1_publish(0, 0);
    for (long long n = 0; n < n_my; ++n) {
      const int buf = (int)(n & 1);
      // eta of panel n+1 goes to the link warp BEFORE this warp waits for r of panel n: the link
      // function of n+1 then overlaps P2(n) (the ring holds at least two panels: T >= 2J)
      if (n + 1 < n_my) pass1_publish(1, buf ^ 1);

      // ---- P2: X^T r from the same resident sub-panels ----
      named_bar_sync(WIDE_BAR_R + buf, WIDE_BAR_COUNT);
      const double2 rA = *reinterpret_cast<const double2*>(r_sh + buf * WR + offA);
      const double2 rB = *reinterpret_cast<const double2*>(r_sh + buf * WR + offB);
#pragma unroll
      for (int i = 0; i < SPW; ++i) {
        if (steps[i] > 0) {
          const double* tile = ring + (size_t)slot[i] * SLOT + cq * WR;
#pragma unroll
          for (int t = 0; t < SPC; ++t) {
            if (t < steps[i]) {
              const double2 xa = *reinterpret_cast<const double2*>(tile + CPS * WR * t + offA);
              const double2 xb = *reinterpret_cast<const double2*>(tile + CPS * WR * t + offB);
              double a = acc[i * SPC + t];
              a = fma(xa.x, rA.x, a);
              a = fma(xa.y, rA.y, a);
              a = fma(xb.x, rB.x, a);
              a = fma(xb.y, rB.y, a);
              acc[i * SPC + t] = a;
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[slot[i]]);
        }
        slot[i] += J;
        if (slot[i] >= T) {
          slot[i] -= T;
          par[i] ^= 1u;
        }
      }
    }

    // ---- this warp's columns of the CTA partial (sum over the row-lanes of each column) ----
#pragma unroll
    for (int i = 0; i < SPW; ++i) {
#pragma unroll
      for (int t = 0; t < SPC; ++t) {
        double v = acc[i * SPC + t];
#pragma unroll
        for (int o = 1; o < LPC; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const int c = (warp + WIDE_CONSUMER_WARPS * i) * KC + cq + CPS * t;
        if (rq == 0 && c < K) my_part[c] = v;
      }
    }
  } else if (warp == WIDE_CONSUMER_WARPS) {
    // =============================== TMA producer ===============================
    // Lane j of the producer warp streams sub-panel j of EVERY row panel (J <= 16 lanes busy): the wait on the slot's
    // empty-barrier, the expect_tx and the bulk copy of a panel's J sub-panels are then ONE pass of warp instructions
    // instead of J iterations of a single-lane loop.  That loop was what held this kernel at ~6.0 TB/s in rounds 1-2
    // whatever the consumers did (five pipeline variants within 1 % of 1.347 ms at K = 1000,
    // profiles/r2_wide_variants.txt): ~390 cycles per 8 KB copy = 21 B/cycle/SM = 6.1 TB/s over 148 SMs.
    if (p.wide_producer == 0) {
      // round 1's producer: one lane walks the ring slot by slot (B200GLM_WIDE_PRODUCER=single; the default only when the
      // ring cannot hold three row panels)
      if (lane == 0) {
        const uint64_t pol = policy_evict_first();
        int sl = 0;
        uint32_t round = 0;
        for (long long n = 0; n < n_my; ++n) {
          const double* src = p.panels + (size_t)(blockIdx.x + n * grid) * Cpad * WR;
          for (int j = 0; j < J; ++j) {
            if (round > 0) mbar_wait(&empty_bar[sl], (round - 1) & 1);
            const uint32_t bytes = (uint32_t)min(KC, Cpad - j * KC) * WR * 8u;
            mbar_arrive_expect_tx(&full_bar[sl], bytes);
            tma_load_1d(ring + (size_t)sl * SLOT, src + (size_t)j * KC * WR, bytes, &full_bar[sl], pol);
            if (++sl == T) {
              sl = 0;
              ++round;
            }
          }
        }
      }
    } else if (lane < J) {
      const uint64_t pol = policy_evict_first();
      const uint32_t bytes = (uint32_t)min(KC, Cpad - lane * KC) * WR * 8u;
      int sl = lane;                          // ring slot of this lane's next sub-panel (n * J + lane) mod T
      uint32_t round = 0;                     // ... and how often the ring has wrapped for it
      // The lanes wait for their slots together: a panel's J copies go out when the last of its J slots is free, so the
      // ring must hold a whole panel beyond the two resident ones (T >= 3 J; the host drops the on-chip chain state for
      // that).  With T = 23, J = 8 (K = 1000 with the state on chip) this form ran at 1.53 ms against 1.35 ms for the
      // single-lane loop; with T = 27 at 1.14 ms = 7.0 TB/s (profiles/r2_wide_variants.txt).  A form in which every lane
      // polls its own slot (mbarrier.test_wait) measured the same as this one in every shape and was dropped.
      for (long long n = 0; n < n_my; ++n) {
        const double* src = p.panels + (size_t)(blockIdx.x + n * grid) * Cpad * WR;
        if (round > 0) mbar_wait(&empty_bar[sl], (round - 1) & 1);
        mbar_arrive_expect_tx(&full_bar[sl], bytes);
        tma_load_1d(ring + (size_t)sl * SLOT, src + (size_t)lane * KC * WR, bytes, &full_bar[sl], pol);
        sl += J;                            // T >= 2 J: at most one wrap
        if (sl >= T) {
          sl -= T;
          ++round;
        }
      }
    }
  } else {
    // =============================== link warp ===============================
    const int r = lane & (WR - 1);
    const int jy = K / KC, coly = K - jy * KC;            // y lives in column K
    const int jt = (K + 1) / KC, colt = K + 1 - jt * KC;  // binomial population sizes in column K+1
    const int Kg = fam_group_col(FAMILY, K);
    const int jg = Kg / KC, colg = Kg - jg * KC;          // group id after the y (and trials) columns (G > 0)
    int slot_y = jy, slot_g = jg, slot_t = jt;
    uint32_t par_y = 0, par_g = 0, par_t = 0;
    const double alpha = G > 0 ? 0.0 : theta_at(0);
    LinkConst lc;
    lc.inv_sigma = 1.0;
    lc.phi = 1.0;
    lc.log_phi = lc.dg_phi = lc.lg_phi = 0.0;
    lc.inc_phi_terms = (!p.mc.propto || !p.mc.lik_only || p.mc.sigma_is_var) ? 1 : 0;
    lc.inc_ytheta = (!p.mc.propto || !p.mc.lik_only || p.mc.sigma_is_var != 2) ? 1 : 0;
    if (FAMILY == FAM_NORMAL_ID) lc.inv_sigma = 1.0 / exp(theta_at(P - 1));  // normal_id_glm_lpdf.hpp:117
    if (FAMILY == FAM_NEG_BINOMIAL_2_LOG) {
      lc.phi = exp(theta_at(P - 1));
      lc.log_phi = log(lc.phi);
      lc.dg_phi = digamma_pos(lc.phi);
      lc.lg_phi = lgamma(lc.phi);
    }
    double lp_acc = 0.0, r_acc = 0.0, x_acc = 0.0;
    for (long long n = 0; n < n_my; ++n) {
      const int buf = (int)(n & 1);
      const long long pi = blockIdx.x + n * grid;
      mbar_wait(&full_bar[slot_y], par_y);
      const double y = ring[(size_t)slot_y * SLOT + coly * WR + r];
      double trials = 0.0;
      if (FAMILY == FAM_BINOMIAL_LOGIT) {
        mbar_wait(&full_bar[slot_t], par_t);
        trials = ring[(size_t)slot_t * SLOT + colt * WR + r];
      }
      double off = alpha;
      if (G > 0) {
        mbar_wait(&full_bar[slot_g], par_g);
        const int gi = (int)ring[(size_t)slot_g * SLOT + colg * WR + r] - 1;
        const bool gok = gi >= 0 && gi < G;
        off = p.stage_a_in_smem ? sa[gok ? gi : 0] : theta_at(2 + (gok ? gi : 0));
      }
      named_bar_sync(WIDE_BAR_ETA + buf, WIDE_BAR_COUNT);
      double eta = 0.0;
#pragma unroll
      for (int w = 0; w < WIDE_CONSUMER_WARPS; ++w) eta += eta_part[(buf * WIDE_CONSUMER_WARPS + w) * WR + r];
      eta += off;
      double lp_i, r_i, x_i;
      link_ext<FAMILY>(eta, y, trials, lc, lp_i, r_i, x_i);
      if (pi * WR + r >= p.n_rows) {
        lp_i = 0.0;
        r_i = 0.0;
        x_i = 0.0;
      }
      if (lane < WR) {
        r_sh[buf * WR + lane] = r_i;
        if (G > 0) p.r_out[pi * WR + lane] = r_i;
        lp_acc += lp_i;
        r_acc += r_i;
        x_acc += x_i;
      }
      __threadfence_block();
      named_bar_arrive(WIDE_BAR_R + buf, WIDE_BAR_COUNT);
      slot_y += J;
      if (slot_y >= T) {
        slot_y -= T;
        par_y ^= 1u;
      }
      slot_g += J;
      if (slot_g >= T) {
        slot_g -= T;
        par_g ^= 1u;
      }
      slot_t += J;
      if (slot_t >= T) {
        slot_t -= T;
        par_t ^= 1u;
      }
    }
    lp_acc = warp_sum(lp_acc);
    r_acc = warp_sum(r_acc);
    x_acc = warp_sum(x_acc);
    if (lane == 0) {
      my_part[K] = lp_acc;
      my_part[K + 1] = r_acc;
      my_part[K + 2] = x_acc;   // neg_binomial_2_log: sum of the per-row d/dphi terms
    }
  }

  cross_cta_reduce_and_finish<FAMILY>(p, sh_scratch, &sh_is_last, st, ring);
}

}  // namespace b200glm
