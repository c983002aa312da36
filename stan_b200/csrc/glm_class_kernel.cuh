// glm_class_kernel.cuh -- the reference's two class-outcome GLMs in ONE pass over X, same launch contract as
// glm_fused_kernel (glm_kernels.cuh): persistent grid, TMA ring of 32-row swizzled panels (the same HBM layout: aux
// column K holds the observed class y in 1..C), private accumulators reduced once, last CTA (ticket) folds the grid's
// partial rows in fixed order and runs the model epilogue (finish_class_model, glm_model.cuh) incl. the leapfrog tail.
//
//   stan::math::ordered_logistic_glm_lpmf   (SM/prim/prob/ordered_logistic_glm_lpmf.hpp:49, row arithmetic :113-207)
//     theta = [beta (K), unconstrained cut-points (C - 1)];  per row: loc = x . beta -> lp, w = d lp / d loc, and the
//     row's two cut-point partials.  beta gradient = X^T w (phase 2 exactly as the narrow kernel); the cut-point
//     partials are kept per LANE in shared memory (one column of 32 lanes per cut-point: a lane adds into its own
//     word whatever class its row has -- no conflicts, no atomics, fixed order) and folded at the end.
//   stan::math::categorical_logit_glm_lpmf  (SM/prim/prob/categorical_logit_glm_lpmf.hpp:47, :88-186)
//     theta = [alpha (C), beta (K x C, column-major)];  per row: lin_c = x . beta[:, c] + alpha_c for all classes in
//     registers (lane = row), softmax -> C weights per row; beta gradient = X^T W with CPL x CMAX accumulators per lane
//     (lane = (column group, row group) as in the narrow kernel, classes innermost).  fp64 FMA work is 4 K C flops per
//     row against 8 K bytes: HBM-bound up to C ~ 4, bound by the fp64 / shared-memory pipes beyond.
//
// Partial row of a CTA / likelihood sums (aligned with theta, as finish_class_model expects):
//   [0, P) d lp / d (constrained parameter), [P] lp-sum.     P = K + C - 1 (ordered) | C (1 + K) (categorical)
#pragma once

#include "glm_kernels.cuh"

namespace b200glm {

constexpr int CLASS_MAX_CLASSES = 16;

__host__ __device__ constexpr int class_partial_stride(int P) { return (P + 1 + 15) & ~15; }
// doubles of dynamic shared memory besides the ring and its barriers
__host__ __device__ inline size_t class_fixed_doubles(int ordered, int K, int NC, int P) {
  const int Pp = (P + 2 + 1) & ~1;
  size_t d = Pp;                                        // theta
  d += NUM_CONSUMER_WARPS * 32;                         // sr: the row weights of a warp's panel (ordered)
  d += (size_t)NUM_CONSUMER_WARPS * Pp;                 // red: per-warp partial rows
  d += ordered ? 16 + (size_t)NUM_CONSUMER_WARPS * (NC > 1 ? NC - 1 : 1) * 32      // cut-points; per-lane cut partials
               : (size_t)NUM_CONSUMER_WARPS * NC * 32 + (size_t)K * CLASS_MAX_CLASSES;   // class weights of a panel per
                                                                                        // warp; beta as [k][class]
  return d;
}

template <bool ORDERED, int CPL, int CMAX>
__global__ void __launch_bounds__(NUM_THREADS, 1) glm_class_kernel(const __grid_constant__ KernelParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int K = p.K, C = p.C, P = p.P, NC = p.n_classes, S = p.n_stages;
  const int tile_doubles = C * PANEL_ROWS;
  const int Pp = (P + 2 + 1) & ~1;
  const int ncut = NC > 1 ? NC - 1 : 1;

  double* tiles = reinterpret_cast<double*>(smem_raw);                  // S * tile_doubles
  double* sth = tiles + (size_t)S * tile_doubles;                       // theta (P), later the likelihood sums (P + 2)
  double* sr = sth + Pp;                                                // 8 x 32
  double* red = sr + NUM_CONSUMER_WARPS * 32;                           // 8 x Pp
  double* scut = red + (size_t)NUM_CONSUMER_WARPS * Pp;                 // ordered: 16 cut-points
  double* cacc = scut + 16;                                             // ordered: 8 x ncut x 32 per-lane cut partials
  double* swt = red + (size_t)NUM_CONSUMER_WARPS * Pp;                  // categorical: 8 x NC x 32 class weights
  double* sbt = swt + (size_t)NUM_CONSUMER_WARPS * NC * 32;             // categorical: beta as [k][CMAX] (zero padded)
  double* after = ORDERED ? cacc + (size_t)NUM_CONSUMER_WARPS * ncut * 32 : sbt + (size_t)K * CLASS_MAX_CLASSES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(after);
  uint64_t* empty_bar = full_bar + S;
  __shared__ double sh_scratch[64];
  __shared__ int sh_is_last;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grid = gridDim.x;

  pdl_launch_dependents();
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();
  constexpr int NC_THREADS = NUM_CONSUMER_WARPS * 32;
  if (warp != NUM_CONSUMER_WARPS) {
    pdl_grid_dependency_wait();
    // theta for this launch (leapfrog: begin_update_p + update_q, expl_leapfrog.hpp:16-26)
    const bool lf = p.mode == MODE_LEAPFROG;
    const double he = 0.5 * p.eps;
    for (int i = tid; i < P; i += NC_THREADS) {
      double q;
      if (lf) {
        const double ph = p.st_in[P + i] - he * p.st_in[2 * P + i];
        q = p.st_in[i] + p.eps * (p.inv_metric[i] * ph);
      } else {
        q = p.theta_inline_n ? p.theta_inline[i] : p.theta_in[i];
      }
      sth[i] = q;
      if (blockIdx.x == 0) p.theta_used[i] = q;
    }
    named_barrier_sync(1, NC_THREADS);
    if (ORDERED) {
      if (tid == 0) {   // ordered_constrain.hpp:34-37
        double c = 0.0;
        for (int k = 0; k < NC - 1; ++k) {
          c = k == 0 ? sth[K] : c + exp(sth[K + k]);
          scut[k] = c;
        }
      }
      for (int j = lane; j < ncut * 32; j += 32) cacc[(size_t)warp * ncut * 32 + j] = 0.0;
      named_barrier_sync(1, NC_THREADS);
    } else {
      for (int idx = tid; idx < K * CMAX; idx += NC_THREADS) {
        const int k = idx / CMAX, c = idx % CMAX;
        sbt[idx] = c < NC ? sth[NC + k + K * c] : 0.0;
      }
      named_barrier_sync(1, NC_THREADS);
    }
  }

  const long long n_panels = p.n_panels;
  const long long p_count = n_panels > blockIdx.x ? (n_panels - blockIdx.x + grid - 1) / grid : 0;

  double acc[CPL][CMAX];
#pragma unroll
  for (int s = 0; s < CPL; ++s)
#pragma unroll
    for (int c = 0; c < CMAX; ++c) acc[s][c] = 0.0;
  double lp_acc = 0.0;
  double wa[CMAX];      // categorical: per-lane sums of the class weights (the alpha gradient)
#pragma unroll
  for (int c = 0; c < CMAX; ++c) wa[c] = 0.0;

  if (warp == NUM_CONSUMER_WARPS) {
    // ===================== TMA producer (one elected lane) =====================
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      const uint32_t bytes = (uint32_t)tile_doubles * 8u;
      int s = 0;
      uint32_t round = 0;
      bool dep_waited = false;
      for (long long n = 0; n < p_count; ++n) {
        const long long pi = blockIdx.x + n * grid;
        if (n == p.pdl_prefetch && !dep_waited) {
          pdl_grid_dependency_wait();
          dep_waited = true;
        }
        if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        tma_load_1d(tiles + (size_t)s * tile_doubles, p.panels + (size_t)pi * tile_doubles, bytes, &full_bar[s], pol);
        if (++s == S) {
          s = 0;
          ++round;
        }
      }
    }
    pdl_grid_dependency_wait();
  } else if (warp < (S < NUM_CONSUMER_WARPS ? S : NUM_CONSUMER_WARPS)) {
    // ===================== consumers: a ring stage belongs to one warp for the whole launch =====================
    const int rg = lane & 3, cg = lane >> 2, cgl = cg & 3;
    const int ycol = K * 32 + (lane ^ ((K & 3) << 2));
    int off[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) off[m] = rg + 4 * (m ^ cgl);
    double* my_sr = sr + warp * 32;
    double* my_cacc = cacc + (size_t)warp * ncut * 32;
    double* my_swt = swt + (size_t)warp * NC * 32;
    const double* sbeta = sth;                           // ordered: beta = theta[0, K)

    uint32_t parity = 0;
    for (long long n0 = 0; n0 < p_count; n0 += S, parity ^= 1u)
    for (int s = warp; s < S; s += NUM_CONSUMER_WARPS) {
      const long long n = n0 + s;
      if (n >= p_count) break;
      const long long pi = blockIdx.x + n * grid;
      mbar_wait(&full_bar[s], parity);
      const double* tile = tiles + (size_t)s * tile_doubles;
      const bool valid = (pi * PANEL_ROWS + lane) < p.n_rows;
      const int y = valid ? (int)tile[ycol] : 1;

      if (ORDERED) {
        // ---- phase 1: location of row `lane`, link, cut-point partials ----
        double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;      // four independent FMA chains, as in the narrow kernel
        const int o1 = lane ^ 4, o2 = lane ^ 8, o3 = lane ^ 12;
        int c = 0;
#pragma unroll 4
        for (; c + 4 <= K; c += 4) {
          e0 = fma(tile[(c + 0) * 32 + lane], sbeta[c], e0);
          e1 = fma(tile[(c + 1) * 32 + o1], sbeta[c + 1], e1);
          e2 = fma(tile[(c + 2) * 32 + o2], sbeta[c + 2], e2);
          e3 = fma(tile[(c + 3) * 32 + o3], sbeta[c + 3], e3);
        }
        for (; c < K; ++c) e0 = fma(tile[c * 32 + (lane ^ ((c & 3) << 2))], sbeta[c], e0);
        double lp_i, w_i, d1, d2;
        ordered_logistic_row((e0 + e1) + (e2 + e3), y, NC, scut, lp_i, w_i, d1, d2);
        if (!valid) {
          lp_i = 0.0;
          w_i = 0.0;
        } else {
          if (y != NC) my_cacc[(y - 1) * 32 + lane] += d2;      // ordered_logistic_glm_lpmf.hpp:202-204
          if (y != 1) my_cacc[(y - 2) * 32 + lane] -= d1;       // :205-207
        }
        lp_acc += lp_i;
        // ---- phase 2: X^T w from the same smem tile ----
        my_sr[lane] = w_i;
        __syncwarp();
        double rr[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) rr[m] = my_sr[rg + 4 * m];
        const double* base = tile + cg * 32;
#pragma unroll
        for (int s2 = 0; s2 < CPL; ++s2)
          if (cg + 8 * s2 < K) {
#pragma unroll
            for (int m = 0; m < 8; ++m) acc[s2][0] = fma(base[s2 * 256 + off[m]], rr[m], acc[s2][0]);
          }
      } else {
        // ---- phase 1: lin_c for row `lane`, all classes in registers ----
        // Plain loop, CMAX independent FMA chains per lane.  Measured alternatives (N = 10M, K = 100, 4 classes):
        // this form 3.51 ms; `#pragma unroll 4` 3.74 ms; two accumulator sets (even / odd columns) 3.96 ms -- the
        // extra live registers cost more than the longer chains (the kernel sits at the 168-register cap).
        double lin[CMAX];
#pragma unroll
        for (int c = 0; c < CMAX; ++c) lin[c] = 0.0;
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
          const double x = tile[k * 32 + (lane ^ ((k & 3) << 2))];
#pragma unroll
          for (int c = 0; c < CMAX; c += 2) {     // one broadcast 16-byte load per class pair
            const double2 b = *reinterpret_cast<const double2*>(sbt + k * CMAX + c);
            lin[c] = fma(x, b.x, lin[c]);
            lin[c + 1] = fma(x, b.y, lin[c + 1]);
          }
        }
        // categorical_logit_glm_lpmf.hpp:88-110 (value), :155-165 (weights: -softmax + [c == y])
        double mx = -CUDART_INF, lin_y = 0.0;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < NC) {
            lin[c] += sth[c];
            mx = lin[c] > mx ? lin[c] : mx;
            if (c == y - 1) lin_y = lin[c];
          }
        double se = 0.0;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < NC) {
            lin[c] = exp(lin[c] - mx);
            se += lin[c];
          }
        const double inv = 1.0 / se;
        lp_acc += valid ? log(inv) - mx + lin_y : 0.0;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < NC) {
            const double w = valid ? -lin[c] * inv + (c == y - 1 ? 1.0 : 0.0) : 0.0;
            wa[c] += w;
            my_swt[c * 32 + lane] = w;
          }
        __syncwarp();
        // ---- phase 2: X^T W, classes innermost ----
        const double* base = tile + cg * 32;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          double w[CMAX];
#pragma unroll
          for (int c = 0; c < CMAX; ++c) w[c] = c < NC ? my_swt[c * 32 + rg + 4 * m] : 0.0;
#pragma unroll
          for (int s2 = 0; s2 < CPL; ++s2)
            if (cg + 8 * s2 < K) {
              const double x = base[s2 * 256 + off[m]];
#pragma unroll
              for (int c = 0; c < CMAX; ++c) acc[s2][c] = fma(x, w[c], acc[s2][c]);
            }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

    // ---- this warp's partial row (theta-aligned) ----
    double* my_red = red + (size_t)warp * Pp;
#pragma unroll
    for (int s2 = 0; s2 < CPL; ++s2)
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        if (ORDERED && c > 0) continue;
        double v = acc[s2][c];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        const int k = cg + 8 * s2;
        if (rg == 0 && k < K && c < (ORDERED ? 1 : NC)) my_red[ORDERED ? k : NC + k + K * c] = v;
      }
    lp_acc = warp_sum(lp_acc);
    if (lane == 0) my_red[P] = lp_acc;
    if (ORDERED) {
      __syncwarp();
      for (int k = 0; k < NC - 1; ++k) {
        const double v = warp_sum(my_cacc[k * 32 + lane]);
        if (lane == 0) my_red[K + k] = v;
      }
    } else {
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        const double v = warp_sum(wa[c]);
        if (lane == 0 && c < NC) my_red[c] = v;
      }
    }
  } else {
    for (int j = lane; j < Pp; j += 32) red[(size_t)warp * Pp + j] = 0.0;   // idle consumer warp
  }
  __syncthreads();

  // ---- CTA partial -> global, ticket ----
  double* my_part = p.partials + (size_t)blockIdx.x * p.pstride;
  for (int j = tid; j < P + 1; j += NUM_THREADS) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < NUM_CONSUMER_WARPS; ++w) v += red[(size_t)w * Pp + j];
    my_part[j] = v;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int t = atomicAdd(p.ticket, 1u);
    sh_is_last = (t == (unsigned int)(grid - 1));
  }
  __syncthreads();
  if (!sh_is_last) return;
  __threadfence();

  // ---- last CTA: the grid's partial rows in fixed order (warp w takes rows w, w + 9, ...; lanes read consecutive
  //      words of a row), folded across warps through the idle ring; then the model epilogue ----
  double* lik = sth;                                   // theta itself is read from p.theta_used by the epilogue
  {
    const int nw = NUM_THREADS / 32, n_sums = P + 1;
    const int n_chunks = (n_sums + 31) >> 5;
    double* scratch = tiles;
    for (int c0 = 0; c0 < n_chunks; c0 += 4) {
      double a4[4] = {0.0, 0.0, 0.0, 0.0};
      for (int r = warp; r < grid; r += nw)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if ((c0 + q) * 32 + lane < n_sums) a4[q] += __ldcg(p.partials + (size_t)r * p.pstride + (c0 + q) * 32 + lane);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (c0 + q < n_chunks) scratch[(size_t)warp * n_chunks * 32 + (c0 + q) * 32 + lane] = a4[q];
    }
    __syncthreads();
    for (int j = tid; j < n_sums; j += NUM_THREADS) {
      double v = 0.0;
      for (int w = 0; w < nw; ++w) v += scratch[(size_t)w * n_chunks * 32 + j];
      lik[j] = v;
    }
    if (tid == 0) {
      lik[P + 1] = 0.0;
      *p.ticket = 0u;
    }
    __syncthreads();
  }
  if (p.peer_in_main && !peer_allreduce_lik(p, &sh_is_last, lik)) {
    peer_timeout_result(p);
    return;
  }
  for (int j = tid; j < P + 2; j += NUM_THREADS) p.lik[j] = lik[j];
  __threadfence();
  __syncthreads();
  class_epilogue(p, sh_scratch);
}

}  // namespace b200glm
