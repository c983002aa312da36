// glm_batched_kernel.cuh -- many chains at once: the per-chain GEMV pair (X beta, X^T r) becomes a
// pair of fp64 GEMMs on the DMMA tensor path (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4, the only fp64
// tensor instruction on sm_100a), fused flash-attention style so that neither E = X*B nor R = f(E)
// (both N x C) ever exists in HBM  (BASELINE configs[2]: normal_id N=1M K=200, 1024 chains).
//
//   E[i,c] = sum_k X[i,k] B[k,c]        GEMM1: m = rows,     n = chains, k = features
//   R[i,c] = link'(E[i,c] + alpha_c)    registers (per-chain lp-sum and r-sum accumulate here too)
//   G[k,c] = sum_i X[i,k] R[i,c]        GEMM2: m = features, n = chains, k = rows
//
// Same arithmetic per chain as the reference's {bernoulli_logit,poisson_log,normal_id}_glm_lp*f
// (link<> in glm_kernels.cuh); same HBM data as the single-chain kernel: the 32-row swizzled panel
// format serves BOTH DMMA A-fragment access patterns (row-major for GEMM1, transposed for GEMM2)
// without bank conflicts, so a handle needs no second copy of X.
//
// Grid = NCB chain blocks (CB = 64 chains) x NS row slices, one CTA per SM.  A CTA keeps its
// K x 64 block of G in registers (8 warps x 13 m-blocks x 2 n-blocks) over its whole row slice, its
// 64 chains' beta in shared memory, and streams its panels through a TMA ring (all NCB CTAs of a
// slice read the same panels at about the same time, so X comes from HBM once and from L2 NCB-1 times).
// Warp w = (mp = w & 1, nq = w >> 1): the pair {2nq, 2nq+1} owns chains [16nq, 16nq+16) end to end --
// GEMM1 splits the 32 rows between the two warps, GEMM2 splits the features -- so R only ever has to
// cross between two warps (a 4 KB smem patch and a 64-thread named barrier), never the whole CTA.
// Slice partials are combined in fixed order by batched_reduce_kernel (deterministic).
//
// Row-split mode (template RS, batches of <= 16 lanes -- most batches of a lock-step NUTS run are the last few
// straggler chains): one warp pair is all the DMMA work 16 chains have per panel, and a single pair per SM is
// latency-bound (1.1 ms per batch at N = 1M, K = 200 against 0.25 ms of HBM time).  So the 16 chains' beta is
// staged ONCE (K x 16) and every pair works on its OWN row panels for the same 16 chains: pair q takes the CTA's
// panels q, q + PA, q + 2 PA, ... through a private ring of SP stages with its own barriers and its own TMA
// issuer, keeps its own K x 16 accumulator block and writes it to columns [16q, 16q + 16) of the CTA's partial
// row -- the slot pair q's chains have in the normal mode -- where batched_reduce_kernel folds the PA blocks.
#pragma once

#include "glm_kernels.cuh"

namespace b200glm {

constexpr int BATCH_CB = 64;                  // chains per CTA
constexpr int BATCH_THREADS = 8 * 32;         // 8 DMMA warps; lane 0 of warp 0 also issues the TMA loads
constexpr int BATCH_MAX_K = 208;              // 2 x 13 m-blocks of 8 features

struct BatchedParams {
  const double* panels;
  long long n_rows, n_panels;
  int K, C, P, off_beta, family;
  int n_stages;
  int NCB, NS;             // chain blocks x row slices = grid
  int ldc;                 // padded number of chain lanes (NCB * 64)
  int n_lanes;             // active lanes; warp pairs whose 16 chains are all padding do no work
  int pairs;               // row-split mode: active warp pairs PA (n_stages is then the stages PER PAIR)
  int lane0;               // glm_multi_kernel: first lane of this pass (lanes [lane0, lane0 + 4))
  const double* theta_c;   // [P][ldc] feature-major: the point each lane is evaluated at
  double* partials;        // [NS][NCB][K + 2][64]: rows [0,K) = X^T r, row K = lp-sum, row K+1 = r-sum
};

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void tma_load_1d_nohint(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                                   uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void pair_bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// S = ring stages in all (row-split mode: pairs x stages per pair), cbw = chains staged in smem (64, or 16)
__host__ __device__ inline size_t batched_smem_bytes(int K, int C, int S, int cbw = BATCH_CB) {
  const int K4 = (K + 3) & ~3;
  return ((size_t)S * C * 32 + (size_t)K4 * cbw + 4 * 32 * 16 + 2 * BATCH_CB) * 8 + (size_t)2 * S * 8;
}

template <int FAMILY, int MBH, bool RS>
__global__ void __launch_bounds__(BATCH_THREADS, 1) glm_batched_kernel(const BatchedParams p) {
  constexpr int CB = BATCH_CB;            // width of a partial row (chain slots of a CTA)
  constexpr int CBW = RS ? 16 : CB;       // chains whose beta is staged in shared memory
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int K = p.K, C = p.C, S = p.n_stages, NS = p.NS, NCB = p.NCB;
  const int PA = RS ? p.pairs : 4;
  const int S_all = RS ? PA * S : S;
  const int K4 = (K + 3) & ~3, Kfull = K & ~3;
  const int tile_doubles = C * 32;
  double* tiles = reinterpret_cast<double*>(smem_raw);            // S_all * C * 32
  double* sB = tiles + (size_t)S_all * tile_doubles;              // K4 x CBW, column index XOR-swizzled
  double* sR = sB + (size_t)K4 * CBW;                             // 4 pairs x (32 rows x 16 chains)
  double* sAl = sR + 4 * 32 * 16;                                 // alpha of the 64 chains
  double* sIs = sAl + CB;                                         // 1 / sigma of the 64 chains
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sIs + CB);
  uint64_t* empty_bar = full_bar + S_all;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cb = blockIdx.x % NCB, sl = blockIdx.x / NCB;
  const long long n_panels = p.n_panels;

  // pairs of this chain block that hold at least one real chain (the last block of a small batch is
  // mostly padding: the straggler batches of a lock-step NUTS run have a handful of lanes)
  const int act_pairs = RS ? PA : min(4, (p.n_lanes - cb * CB + 15) / 16);
  if (tid == 0) {
    for (int s = 0; s < S_all; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], RS ? 2 : 2 * act_pairs);   // row-split: a stage belongs to one pair
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  // this CTA's chains' beta: sB[k][c ^ ((k & 3) << 2)]
  for (int idx = tid; idx < K4 * CBW; idx += BATCH_THREADS) {
    const int k = idx / CBW, c = idx % CBW;
    sB[k * CBW + (c ^ ((k & 3) << 2))] = k < K ? p.theta_c[(size_t)(p.off_beta + k) * p.ldc + cb * CB + c] : 0.0;
  }
  for (int c = tid; c < CB; c += BATCH_THREADS) {
    sAl[c] = p.theta_c[(size_t)cb * CB + c];
    sIs[c] = FAMILY == FAM_NORMAL_ID ? 1.0 / exp(p.theta_c[(size_t)(p.P - 1) * p.ldc + cb * CB + c]) : 1.0;
  }
  __syncthreads();

  // ===================== TMA issue (lane 0 of warp 0) =====================
  // Tile m goes to stage m % S; it is requested at the top of iteration m - (S - 1), once all 8 warps
  // have released that stage (they finished GEMM2 of tile m - S one iteration ago).  No separate
  // producer warp: 8 warps = 2 per SM sub-partition leaves each thread the full 255-register budget.
  // Row-split mode: the same scheme per pair -- pair nq's m-th tile is the CTA's tile nq + m * PA, its ring is
  // stages [nq * S, (nq + 1) * S), its issuer lane 0 of its first warp.
  const uint32_t tile_bytes = (uint32_t)tile_doubles * 8u;
  const int mp = warp & 1, nq = warp >> 1;
  const int st0 = RS ? nq * S : 0;                     // first stage of the ring this warp works on
  const long long tile0 = RS ? nq : 0, tile_step = RS ? PA : 1;
  const bool issuer = RS ? (mp == 0 && lane == 0) : (tid == 0);
  auto request_tile = [&](long long m) {
    const long long pm = sl + (tile0 + m * tile_step) * NS;
    if (pm >= n_panels) return;
    const int sm = st0 + (int)(m % S);
    if (m >= S) mbar_wait(&empty_bar[sm], (uint32_t)(((m / S) - 1) & 1));
    mbar_arrive_expect_tx(&full_bar[sm], tile_bytes);
    tma_load_1d_nohint(tiles + (size_t)sm * tile_doubles, p.panels + (size_t)pm * tile_doubles, tile_bytes,
                       &full_bar[sm]);
  };
  if (nq >= act_pairs) return;   // nothing but padding lanes (never warp 0: pair 0 always has lane 0)
  if (issuer)
    for (int m = 0; m < S - 1; ++m) request_tile(m);

  // ===================== DMMA warps =====================
  const int l4 = lane & 3, lq = lane >> 2;
  const int sw = l4 << 2;
  double* sRp = sR + nq * (32 * 16);
  const int MB = (K + 7) >> 3, MB0 = (MB + 1) >> 1;
  const int mb_base = mp ? MB0 : 0, mb_cnt = mp ? MB - MB0 : MB0;
  const bool dense = (MB - MB0) >= MBH - 1;          // both warps of a pair fill (all but at most one of) their slots
  const int swr = ((l4 & 1) << 3) | ((l4 & 2) << 1); // residual-patch swizzle of row kr + l4 (see the link step)

  double g[MBH][2][2];
#pragma unroll
  for (int m = 0; m < MBH; ++m) g[m][0][0] = g[m][0][1] = g[m][1][0] = g[m][1][1] = 0.0;
  double lp_acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, r_acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};

  // fragment addressing (see header): A1 = X[row][k] for GEMM1, B1 = beta[k][chain]
  const int ra0 = (16 * mp + lq) ^ sw, ra1 = (16 * mp + 8 + lq) ^ sw;
  const int coff = RS ? 0 : 16 * nq;     // this pair's chains within the staged beta block
  const int cb0 = (coff + lq) ^ sw, cb1 = (coff + 8 + lq) ^ sw;
  const double* tb = sB + l4 * CBW;
  const int fsw = (lq & 3) << 2;
  const int ycol = K * 32, ysw = (K & 3) << 2;

  long long n = 0;
  for (long long pi = sl + tile0 * NS; pi < n_panels; pi += tile_step * NS, ++n) {
    const int s = st0 + (int)(n % S);
    if (issuer) request_tile(n + S - 1);
    __syncwarp();
    mbar_wait(&full_bar[s], (uint32_t)((n / S) & 1));
    const double* tile = tiles + (size_t)s * tile_doubles;

    // ---- GEMM1: E (16 rows of this warp x 16 chains of this pair) ----
    // two accumulator sets (even / odd k4 steps): 8 independent DMMA chains per warp instead of 4
    double e[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
    double f[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
    const double* ta = tile + l4 * 32;
    int k0 = 0;
#pragma unroll 2
    for (; k0 + 8 <= Kfull; k0 += 8) {
      const double a0 = ta[k0 * 32 + ra0], a1 = ta[k0 * 32 + ra1];
      const double b0 = tb[k0 * CBW + cb0], b1 = tb[k0 * CBW + cb1];
      const double c0 = ta[(k0 + 4) * 32 + ra0], c1 = ta[(k0 + 4) * 32 + ra1];
      const double d0 = tb[(k0 + 4) * CBW + cb0], d1 = tb[(k0 + 4) * CBW + cb1];
      dmma884(e[0][0], a0, b0);
      dmma884(e[0][1], a0, b1);
      dmma884(e[1][0], a1, b0);
      dmma884(e[1][1], a1, b1);
      dmma884(f[0][0], c0, d0);
      dmma884(f[0][1], c0, d1);
      dmma884(f[1][0], c1, d0);
      dmma884(f[1][1], c1, d1);
    }
    for (; k0 < K4; k0 += 4) {  // at most one full step and one partial step (K % 4 != 0)
      double a0 = ta[k0 * 32 + ra0], a1 = ta[k0 * 32 + ra1];
      if (k0 + l4 >= K) {  // the y column must not leak into eta
        a0 = 0.0;
        a1 = 0.0;
      }
      const double b0 = tb[k0 * CBW + cb0], b1 = tb[k0 * CBW + cb1];
      dmma884(e[0][0], a0, b0);
      dmma884(e[0][1], a0, b1);
      dmma884(e[1][0], a1, b0);
      dmma884(e[1][1], a1, b1);
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) {
        e[mi][ni][0] += f[mi][ni][0];
        e[mi][ni][1] += f[mi][ni][1];
      }

    // ---- link: C-fragment (row = lq, chains 2*l4 + {0,1}) -> residual patch of the pair ----
    pair_bar_sync(1 + nq);  // the partner has finished reading the previous patch
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int row = 16 * mp + 8 * mi + lq;
      const double y = tile[ycol + (row ^ ysw)];
      const double aux = FAMILY == FAM_BINOMIAL_LOGIT ? tile[ycol + 32 + (row ^ (((p.K + 1) & 3) << 2))] : 0.0;   // population size
      const bool valid = pi * 32 + row < p.n_rows;
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) {
        double rr[2];
        const double2 al = *reinterpret_cast<const double2*>(sAl + coff + 8 * ni + 2 * l4);
        const double2 is = *reinterpret_cast<const double2*>(sIs + coff + 8 * ni + 2 * l4);
#pragma unroll
        for (int e2 = 0; e2 < 2; ++e2) {
          double lp_i, r_i;
          if (FAMILY == FAM_BINOMIAL_LOGIT) {
            double x_i;
            link_ext<FAMILY>(e[mi][ni][e2] + (e2 ? al.y : al.x), y, aux, LinkConst(), lp_i, r_i, x_i);
          } else {
            link<FAMILY>(e[mi][ni][e2] + (e2 ? al.y : al.x), y, e2 ? is.y : is.x, lp_i, r_i);
          }
          if (!valid) {
            lp_i = 0.0;
            r_i = 0.0;
          }
          lp_acc[ni][e2] += lp_i;
          r_acc[ni][e2] += r_i;
          rr[e2] = r_i;
        }
        // residual patch: element (row, chain c) lives at column c ^ f(row & 3), f = {0, 8, 4, 12}.  Any permutation
        // of {0, 4, 8, 12} keeps the GEMM2 B-fragment reads conflict-free (a half-warp's four rows land in four
        // different column quads); f(0) ^ f(1) and f(2) ^ f(3) having the 8-bit set is what this 16-byte store needs:
        // the two rows of a quarter-warp then cover all 16 columns.  (Round 1 used f = 4 (row & 3): both rows on the
        // same 8 columns, a 2-way conflict on every store -- 65.6 M conflicts per launch at config 3, ncu.)
        *reinterpret_cast<double2*>(sRp + row * 16 + ((8 * ni + 2 * l4) ^ (((row & 1) << 3) | ((row & 2) << 1)))) =
            make_double2(rr[0], rr[1]);
      }
    }
    pair_bar_sync(1 + nq);

    // ---- GEMM2: G (this warp's features x 16 chains) += X^T (features x 32 rows) * R ----
    const double* tf = tile + (size_t)(mb_base * 8 + lq) * 32;
    if (dense) {
      // every m-block slot of this warp is issued: the short warp of the pair (K = 200: 12 of 13 blocks) repeats
      // its block 0 into an accumulator that is never stored.  No guard around the DMMA pair -- guarded, each
      // pair came out as ISETP + predicated WARPSYNC + NOP + 2 DMMA (profiles/r1_sass_mnemonics.txt).
#pragma unroll 2
      for (int kr = 0; kr < 32; kr += 4) {
        const int rowb = kr + l4;
        const double b0 = sRp[rowb * 16 + (lq ^ swr)], b1 = sRp[rowb * 16 + ((8 + lq) ^ swr)];
        const int ao = rowb ^ fsw;
#pragma unroll
        for (int m = 0; m < MBH; ++m) {
          const double a = tf[(m < mb_cnt ? m : 0) * 256 + ao];
          dmma884(g[m][0], a, b0);
          dmma884(g[m][1], a, b1);
        }
      }
    } else {
#pragma unroll 2
      for (int kr = 0; kr < 32; kr += 4) {
        const int rowb = kr + l4;
        const double b0 = sRp[rowb * 16 + (lq ^ swr)], b1 = sRp[rowb * 16 + ((8 + lq) ^ swr)];
        const int ao = rowb ^ fsw;
#pragma unroll
        for (int m = 0; m < MBH; ++m) {
          if (m < mb_cnt) {
            const double a = tf[m * 256 + ao];
            dmma884(g[m][0], a, b0);
            dmma884(g[m][1], a, b1);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }

  // ---- CTA partial -> global ----
  double* part = p.partials + ((size_t)sl * NCB + cb) * (size_t)(K + 2) * CB;
#pragma unroll
  for (int m = 0; m < MBH; ++m) {
    const int f = (mb_base + m) * 8 + lq;
    if (m < mb_cnt && f < K) {
#pragma unroll
      for (int ni = 0; ni < 2; ++ni)
        *reinterpret_cast<double2*>(part + (size_t)f * CB + 16 * nq + 8 * ni + 2 * l4) =
            make_double2(g[m][ni][0], g[m][ni][1]);
    }
  }
  // lp / r sums: over the 8 row-lanes, then over the two warps of the pair
#pragma unroll
  for (int ni = 0; ni < 2; ++ni)
#pragma unroll
    for (int e2 = 0; e2 < 2; ++e2) {
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        lp_acc[ni][e2] += __shfl_xor_sync(0xffffffffu, lp_acc[ni][e2], o);
        r_acc[ni][e2] += __shfl_xor_sync(0xffffffffu, r_acc[ni][e2], o);
      }
    }
  pair_bar_sync(1 + nq);
  if (mp == 1 && lq == 0) {
#pragma unroll
    for (int ni = 0; ni < 2; ++ni)
#pragma unroll
      for (int e2 = 0; e2 < 2; ++e2) {
        sRp[8 * ni + 2 * l4 + e2] = lp_acc[ni][e2];
        sRp[16 + 8 * ni + 2 * l4 + e2] = r_acc[ni][e2];
      }
  }
  pair_bar_sync(1 + nq);
  if (mp == 0 && lq == 0) {
#pragma unroll
    for (int ni = 0; ni < 2; ++ni)
#pragma unroll
      for (int e2 = 0; e2 < 2; ++e2) {
        const int c = 8 * ni + 2 * l4 + e2;
        part[(size_t)K * CB + 16 * nq + c] = lp_acc[ni][e2] + sRp[c];
        part[(size_t)(K + 1) * CB + 16 * nq + c] = r_acc[ni][e2] + sRp[16 + c];
      }
  }
}

// ------------------------------------------------------------------------------------------
// Per-lane prologue / epilogue of a batched evaluation.  Chain state is feature-major
// ([P][ld_state], one column per chain slot) so that every access below is coalesced over chains.
// ------------------------------------------------------------------------------------------
struct BatchedStepParams {
  int n;                    // active lanes
  int P, K, off_beta, family, ldc, ld_state, NCB, NS;
  int mode;                 // MODE_THETA: theta_in -> result; MODE_LEAPFROG: chain state advanced
  const int32_t* chains;    // [n] chain slot of lane i (device); NULL = identity
  const double* eps;        // [n] per-lane step size (device); NULL = eps_scalar
  int eps_by_chain;         // eps is indexed by chain slot, not by lane (device-side NUTS: the state machine sets it)
  double eps_scalar;
  const double* theta_in;   // MODE_THETA: [n][P] chain-major
  double *Q, *Pm, *Gd, *V, *IM;   // chain state [P][ld_state] (V: [ld_state])
  double* theta_c;          // [P][ldc] evaluation points (out of begin, in of main + finish)
  double* p_half;           // [P][ldc] momenta after the first half step
  const double* partials;
  double* reduced;          // [K + 2][ldc]: the slice partials summed over the NS row slices (batched_reduce_kernel)
  int fold;                 // row-split mode: the PA 16-column blocks of a partial row belong to the same 16 chains
  double* result;           // [n][P + 2] chain-major: lp, grad[P], status
  double* state_out;        // MODE_LEAPFROG: [n][3P + 1] chain-major mirror of (q, p, g, V) or NULL
  ModelConst mc;
};

// theta for every lane: MODE_THETA transposes the input; MODE_LEAPFROG does begin_update_p + update_q
// (expl_leapfrog.hpp:16-26).  Lanes [n, ldc) are zero-filled.
__global__ void __launch_bounds__(256) batched_begin_kernel(const BatchedStepParams p) {
  const int lane_total = p.ldc;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < (long long)p.P * lane_total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % lane_total), k = (int)(idx / lane_total);
    double th = 0.0, ph = 0.0;
    if (i < p.n) {
      if (p.mode == MODE_LEAPFROG) {
        const int c = p.chains ? p.chains[i] : i;
        const double eps = p.eps ? p.eps[p.eps_by_chain ? c : i] : p.eps_scalar;
        const size_t o = (size_t)k * p.ld_state + c;
        ph = p.Pm[o] - (0.5 * eps) * p.Gd[o];
        th = p.Q[o] + eps * (p.IM[o] * ph);
      } else {
        th = p.theta_in[(size_t)i * p.P + k];
      }
    }
    p.theta_c[(size_t)k * p.ldc + i] = th;
    p.p_half[(size_t)k * p.ldc + i] = ph;
  }
}

// set_state for n chain slots: chain-major staging [n][4P + 1] = (q, p, g, inv_metric, V) -> feature-major state
__global__ void __launch_bounds__(256) batched_scatter_state_kernel(int n, int P, int ld_state,
                                                                    const int32_t* __restrict__ chains,
                                                                    const double* __restrict__ in, int has_metric,
                                                                    double* Q, double* Pm, double* Gd, double* IM,
                                                                    double* V) {
  const int W = 4 * P + 1;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < (long long)n * (P + 1);
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % n), k = (int)(idx / n);
    const int c = chains ? chains[i] : i;
    const double* src = in + (size_t)i * W;
    if (k == P) {
      V[c] = src[4 * P];
    } else {
      const size_t o = (size_t)k * ld_state + c;
      Q[o] = src[k];
      Pm[o] = src[P + k];
      Gd[o] = src[2 * P + k];
      if (has_metric) IM[o] = src[3 * P + k];
    }
  }
}

// Sum of the NS slice partials of every (row, chain) in fixed slice order (deterministic), 8 loads in flight per
// thread.  One CTA of 64 threads per (row of [0, K+2), chain block).  Its own launch because a small batch -- the
// straggler batches of a lock-step NUTS run -- has NS = 148 slices and a single finish CTA: summing them there was
// a dependent chain of 148 x (K+2)/8 L2 loads per thread, ~1.3 ms per batch at K = 200, twice the DMMA kernel.
__global__ void __launch_bounds__(BATCH_CB) batched_reduce_kernel(const BatchedStepParams p) {
  const int row = blockIdx.x / p.NCB, cb = blockIdx.x % p.NCB, cl = threadIdx.x;
  const size_t stride = (size_t)p.NCB * (p.K + 2) * BATCH_CB;
  const double* src = p.partials + ((size_t)cb * (p.K + 2) + row) * BATCH_CB + cl;
  double a[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  int s = 0;
  for (; s + 8 <= p.NS; s += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] += __ldcg(src + (size_t)(s + u) * stride);
  }
#pragma unroll
  for (int u = 0; u < 7; ++u)
    if (s + u < p.NS) a[u] += __ldcg(src + (size_t)(s + u) * stride);
  double v = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  if (p.fold > 0) {   // row-split mode (NCB == 1): chain cl < 16 is the sum of columns cl, 16 + cl, ... in pair order
    __shared__ double sh[BATCH_CB];
    sh[cl] = v;
    __syncthreads();
    if (cl >= 16) return;
    v = sh[cl];
    for (int q = 1; q < p.fold; ++q) v += sh[16 * q + cl];
  }
  p.reduced[(size_t)row * p.ldc + cb * BATCH_CB + cl] = v;
}

// Slice partials -> per-chain model lp / gradient (priors, Jacobian: same formulas as finish() for
// G == 0) and, in leapfrog mode, end_update_p + state write-back.  One CTA per 32 lanes (lane = chain),
// 8 warps stride over the features.
__global__ void __launch_bounds__(256) batched_finish_kernel(const BatchedStepParams p) {
  __shared__ double sh_b2[8][32], sh_bad[8][32];
  __shared__ double sh_lp[32], sh_dom[32];
  const ModelConst& mc = p.mc;
  const int P = p.P, K = p.K;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const bool act = i < p.n;
  const double ib2 = 1.0 / (mc.prior_beta_sd * mc.prior_beta_sd);
  double* res = p.result + (size_t)(act ? i : 0) * (P + 2);
  auto slice_sum = [&](int row) { return p.reduced[(size_t)row * p.ldc + i]; };

  double sb = 0.0, bad = 0.0;
  if (act) {
    for (int k = w; k < K; k += 8) {
      const double b = p.theta_c[(size_t)(p.off_beta + k) * p.ldc + i];
      const double lik = slice_sum(k);
      sb += b * b;
      if (!isfinite(b) || !isfinite(lik)) bad += 1.0;
      res[1 + p.off_beta + k] = lik - b * ib2;
    }
  }
  sh_b2[w][lane] = sb;
  sh_bad[w][lane] = bad;
  __syncthreads();
  if (w == 0 && act) {
    double sum_b2 = 0.0, n_bad = 0.0;
    for (int j = 0; j < 8; ++j) {
      sum_b2 += sh_b2[j][lane];
      n_bad += sh_bad[j][lane];
    }
    const double S = slice_sum(K), Sr = slice_sum(K + 1);
    const double alpha = p.theta_c[i];
    const double u_s = mc.family == FAM_NORMAL_ID ? p.theta_c[(size_t)(P - 1) * p.ldc + i] : 0.0;
    const double sigma = mc.family == FAM_NORMAL_ID ? exp(u_s) : 1.0;
    if (!isfinite(alpha) || !isfinite(u_s) || !isfinite(S) || !isfinite(Sr)) n_bad += 1.0;
    const bool dens = (!mc.propto) || mc.is_var;
    double lp = 0.0;
    if (mc.jacobian && mc.family == FAM_NORMAL_ID) lp += u_s;  // lb_constrain.hpp:64
    if (dens) {
      const double z0 = alpha / mc.prior_alpha_sd;
      lp += -0.5 * z0 * z0;
      if (!mc.propto) lp += NEG_LOG_SQRT_TWO_PI_D - log(mc.prior_alpha_sd);
      if (K > 0) {
        lp += -0.5 * sum_b2 * ib2;
        if (!mc.propto) lp += K * (NEG_LOG_SQRT_TWO_PI_D - log(mc.prior_beta_sd));
      }
      if (mc.family == FAM_NORMAL_ID) {
        const double z = (sigma - mc.prior_sigma_loc) / mc.prior_sigma_scale;
        lp += -0.5 * z * z;
        if (!mc.propto) lp += NEG_LOG_SQRT_TWO_PI_D - log(mc.prior_sigma_scale);
      }
      if (mc.N_total > 0) {
        if (mc.family == FAM_BERNOULLI_LOGIT) {
          lp += S;
        } else if (mc.family == FAM_POISSON_LOG || mc.family == FAM_BINOMIAL_LOGIT) {
          lp += S;
          if (!mc.propto) lp -= mc.lgamma_sum;   // poisson_log_glm_lpmf.hpp:127-129, binomial_logit_glm_lpmf.hpp:127-130
        } else {
          if (!mc.propto) lp += NEG_LOG_SQRT_TWO_PI_D * mc.N_total;
          lp -= mc.N_total * u_s;
          lp -= 0.5 * S;
        }
      }
    }
    const bool domain = !(isfinite(lp) && n_bad == 0.0);
    res[0] = lp;
    res[1] = Sr - alpha / (mc.prior_alpha_sd * mc.prior_alpha_sd);
    if (mc.family == FAM_NORMAL_ID) {
      const double dlik = mc.N_total > 0 ? (S - mc.N_total) / sigma : 0.0;  // normal_id_glm_lpdf.hpp:181-183
      const double dpri = -(sigma - mc.prior_sigma_loc) / (mc.prior_sigma_scale * mc.prior_sigma_scale);
      res[1 + P - 1] = (dlik + dpri) * sigma + (mc.jacobian ? 1.0 : 0.0);
    }
    res[1 + P] = domain ? (double)ST_DOMAIN : (double)ST_OK;
    sh_lp[lane] = lp;
    sh_dom[lane] = domain ? 1.0 : 0.0;
  }
  __syncthreads();
  if (p.mode != MODE_LEAPFROG || !act) return;

  // ---- end_update_p + write-back (expl_leapfrog.hpp:28-32; base_hamiltonian.hpp:64-69) ----
  const bool domain = sh_dom[lane] != 0.0;
  const int c = p.chains ? p.chains[i] : i;
  const double eps = p.eps ? p.eps[p.eps_by_chain ? c : i] : p.eps_scalar;
  const double he = 0.5 * eps;
  double* so = p.state_out ? p.state_out + (size_t)i * (3 * P + 1) : nullptr;
  for (int k = w; k < P; k += 8) {
    const size_t o = (size_t)k * p.ld_state + c;
    const double qn = p.theta_c[(size_t)k * p.ldc + i];
    const double ph = p.p_half[(size_t)k * p.ldc + i];
    const double gnew = domain ? -p.Gd[o] : -res[1 + k];
    const double pn = ph - he * gnew;
    p.Q[o] = qn;
    p.Gd[o] = gnew;
    p.Pm[o] = pn;
    if (so) {
      so[k] = qn;
      so[P + k] = pn;
      so[2 * P + k] = gnew;
    }
  }
  if (w == 0) {
    const double Vn = domain ? CUDART_INF : -sh_lp[lane];
    p.V[c] = Vn;
    if (so) so[3 * P] = Vn;
  }
}

}  // namespace b200glm
