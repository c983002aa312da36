// glm_kernels.cuh -- hand-written sm_100a kernels for the GLM log-density + gradient hot path.
//
// What is computed is the arithmetic of the reference's three GLM densities
//   stan::math::bernoulli_logit_glm_lpmf  (SM/prim/prob/bernoulli_logit_glm_lpmf.hpp:105-164)
//   stan::math::poisson_log_glm_lpmf      (SM/prim/prob/poisson_log_glm_lpmf.hpp:107-161)
//   stan::math::normal_id_glm_lpdf        (SM/prim/prob/normal_id_glm_lpdf.hpp:117-213)
// and, in the single-chain kernel for K <= 256 ("the remaining GLMs"),
//   stan::math::binomial_logit_glm_lpmf      (SM/prim/prob/binomial_logit_glm_lpmf.hpp:104-157)
//   stan::math::neg_binomial_2_log_glm_lpmf  (SM/prim/prob/neg_binomial_2_log_glm_lpmf.hpp:145-246)
// plus the model wrapper (priors, lb_constrain Jacobian) and the leapfrog update
// (ST/mcmc/hmc/integrators/expl_leapfrog.hpp:16-32).  How it is computed is new: ONE pass over X.
//
// Data layout in HBM ("row-panel format", built once at upload by relayout_kernel):
//   rows are cut into panels of 32; panel p is one contiguous block of C = K + n_aux columns,
//   each column 32 doubles:  panel[p][c][ r ^ swz(c) ],  swz(c) = (c & 3) << 2.
//   aux column K holds y (as double), then the binomial population sizes (binomial_logit only), then
//   the 1-based group id (as double) when G > 0.
//   One panel = C*256 bytes = one cp.async.bulk (TMA, SASS UBLKCP) into one shared-memory stage.
//   The XOR swizzle makes BOTH access patterns below bank-conflict free.
//
// Main kernel (persistent, one CTA per SM, 1 TMA producer warp + 8 consumer warps):
//   producer : streams this CTA's panels (p = cta, cta+grid, ...) through an S-stage mbarrier ring.
//   consumer warp w takes the CTA's n-th panel when n % 8 == w and, from that smem stage only,
//     phase 1 (lane = row):        eta_r = sum_c X[r][c]*beta[c] + alpha|a[group_r];  link -> lp_r, r_r
//     phase 2 (lane = (cg, rg)):   acc[s] += X[rg+4m][cg+8s] * r[rg+4m]   (private accumulators,
//                                   reduced across lanes/warps/CTAs only once, at the end)
//   last CTA (ticket) sums the per-CTA partials in fixed order -> deterministic result, then
//   (single-GPU, no groups) runs the model epilogue: priors, Jacobian, leapfrog half-step.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "glm_link.cuh"   // family ids, link<> / link_ext<> and their special functions (also compiled for the host)
#include "glm_model.cuh"  // KernelParams, ModelConst and the model epilogue finish() (also compiled for the host)

namespace b200glm {

constexpr int PANEL_ROWS = 32;
constexpr int NUM_CONSUMER_WARPS = 8;
constexpr int NUM_THREADS = (NUM_CONSUMER_WARPS + 1) * 32;
constexpr int MAX_STAGES = 32;
constexpr int SMEM_A_MAX_GROUPS = 2048;  // a[G] staged in smem up to this many groups

// aux columns of a panel: y at K, trials at K+1 (binomial_logit only), group id after those (G > 0 only)
__host__ __device__ constexpr int fam_group_col(int f, int K) { return K + 1 + (f == FAM_BINOMIAL_LOGIT ? 1 : 0); }
// doubles per CTA partial row: K beta gradients, lp-sum, r-sum, aux-sum (d/dphi terms of neg_binomial_2_log)
__host__ __device__ constexpr int partial_stride(int K) { return (K + 3 + 15) & ~15; }   // whole 128-byte lines
// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA engine)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// global -> shared bulk copy, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                            uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// Programmatic dependent launch (PDL): consecutive evaluations of a slot are consecutive launches on one stream.
// launch_dependents lets the NEXT launch's CTAs take an SM as soon as this launch's CTA on it has exited (they
// set up barriers and start streaming X, which is constant data); grid_dependency_wait is what orders every read
// of the previous launch's results (theta / leapfrog state) after that launch has completed and flushed.
// Both are no-ops for a launch without the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void named_barrier_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ------------------------------------------------------------------------------------------
// All-reduce (sum) of the P+2 likelihood partials across the row shards, INSIDE the launch that
// produced them.  One CTA per rank pushes its partial into every peer's mailbox with plain 8-byte peer
// stores and reads the world's entries from its own mailbox; the sum runs in rank order -- the same
// order on every rank, so theta stays bitwise replicated.
// Protocol (Lamport-style, no fence, no flag): every mailbox word is self-validating -- it holds
// PEER_EMPTY until the sender's value lands, and an aligned 8-byte store is indivisible -- so the receiver
// simply polls the words it needs.  One NVLink one-way latency per exchange; the first version (payload
// stores, __threadfence_system, st.release flag, ld.acquire spin) paid two round trips.
// Three buffers rotate by evaluation number n: n % 3 is read now, (n + 1) % 3 may already be receiving the
// next evaluation from a faster peer, (n + 2) % 3 -- last read at n - 1, next written at n + 2, which no
// peer can start before it has this rank's n + 1 contribution (a later LAUNCH) -- is re-armed here.
// Replaces ncclAllReduce + a separate epilogue launch (two launches and ~20 us per gradient).
// Called by every thread of that CTA after p.lik is complete; returns false on timeout.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* ptr, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(ptr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* ptr) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Per-phase time stamps of a launch (b200glm_timeline_enable): row `row` of [grid + 1][16] words,
// word k = %globaltimer (ns, comparable across SMs and GPUs of a box), word 8 + k = clock64 of the SM.
__device__ __forceinline__ void tl_stamp(const KernelParams& p, int row, int k) {
  if (p.tl) {
    p.tl[(size_t)row * 16 + k] = globaltimer_ns();
    p.tl[(size_t)row * 16 + 8 + k] = (unsigned long long)clock64();
  }
}
__device__ bool peer_allreduce_lik(const KernelParams& p, int* sh_ok, double* lik /* P + 2, shared or global */) {
  const PeerParams& pe = p.peer;
  const int n = p.P + 2, tid = threadIdx.x, nt = blockDim.x;
  const int buf = (int)(pe.seq % PEER_BUFS), buf_rearm = (int)((pe.seq + 2) % PEER_BUFS);
  const size_t my_off = ((size_t)buf * pe.world + pe.rank) * pe.stride;
  if (tid == 0) *sh_ok = 1;
  __syncthreads();
  unsigned long long* mine = reinterpret_cast<unsigned long long*>(pe.mbox[pe.rank]);
  auto own_word = [&](int j) {
    unsigned long long own = (unsigned long long)__double_as_longlong(lik[j]);
    return own == PEER_EMPTY_BITS ? 0x7FF8000000000000ull : own;
  };
  for (int j = tid; j < n; j += nt) {            // every send is in flight before the first poll
    const unsigned long long own = own_word(j);
#pragma unroll
    for (int r = 0; r < MAX_PEERS; ++r)
      if (r < pe.world && r != pe.rank)
        st_relaxed_sys_u64(reinterpret_cast<unsigned long long*>(pe.mbox[r]) + my_off + j, own);
#pragma unroll
    for (int r = 0; r < MAX_PEERS; ++r)          // re-arm the buffer of evaluation n + 2
      if (r < pe.world) mine[((size_t)buf_rearm * pe.world + r) * pe.stride + j] = PEER_EMPTY_BITS;
  }
  const unsigned long long t0 = globaltimer_ns();
  for (int j = tid; j < n; j += nt) {
    const unsigned long long own = own_word(j);
    const unsigned long long* src = mine + (size_t)buf * pe.world * pe.stride + j;
    unsigned long long w[MAX_PEERS];
    bool all;
    do {
      all = true;
#pragma unroll
      for (int r = 0; r < MAX_PEERS; ++r)
        if (r < pe.world) {
          w[r] = r == pe.rank ? own : ld_relaxed_sys_u64(src + (size_t)r * pe.stride);
          all = all && (w[r] != PEER_EMPTY_BITS);
        }
      if (!all && globaltimer_ns() - t0 > pe.timeout_ns) {
        *sh_ok = 0;
        break;
      }
    } while (!all);
    double v = 0.0;
#pragma unroll
    for (int r = 0; r < MAX_PEERS; ++r)
      if (r < pe.world) v += __longlong_as_double((long long)w[r]);
    lik[j] = v;
  }
  __threadfence();
  __syncthreads();
  return *sh_ok != 0;
}
__device__ void peer_timeout_result(const KernelParams& p) {
  if (threadIdx.x == 0) {
    p.result[0] = CUDART_NAN;
    p.result[1 + p.P] = (double)ST_PEER_TIMEOUT;
    if (p.mode == MODE_LEAPFROG) p.st_out[3 * p.P] = CUDART_NAN;
    if (p.host_out) {
      p.host_out[0] = CUDART_NAN;
      p.host_out[1 + p.P] = (double)ST_PEER_TIMEOUT;
      if (p.mode == MODE_LEAPFROG) p.host_out[(p.P + 2) + 3 * p.P] = CUDART_NAN;
    }
  }
  host_out_publish(p);
}

// ------------------------------------------------------------------------------------------
// Shared tail of both main kernels.  Every CTA has written its partial sums
// (my_part[0..K) = beta gradient, [K] = lp-sum, [K+1] = r-sum) to p.partials; the LAST CTA to
// arrive (ticket) adds them in fixed CTA order -> bitwise deterministic, then (single GPU, no
// groups) runs the model epilogue.  Must be called by all threads of the CTA.
// ------------------------------------------------------------------------------------------
// On-chip copy of the chain state a launch works on (KernelParams::state_in_smem): theta (the evaluated point),
// the momentum after begin_update_p, the gradient the step started from, and room for the likelihood sums --
// everything the epilogue needs, so that the last CTA's tail does not go back to global memory for it.
struct StateSmem {
  double *theta, *ph, *g0, *lik;   // P, P, P, P + 2 doubles (NULL: not staged)
};
__host__ __device__ constexpr int state_smem_doubles(int P) { return 4 * ((P + 3) & ~1); }
__device__ __forceinline__ StateSmem carve_state_smem(double* base, int P, int enabled) {
  const int Pp = (P + 3) & ~1;
  StateSmem st;
  st.theta = enabled ? base : nullptr;
  st.ph = enabled ? base + Pp : nullptr;
  st.g0 = enabled ? base + 2 * Pp : nullptr;
  st.lik = enabled ? base + 3 * Pp : nullptr;
  return st;
}
// theta for this launch (leapfrog: begin_update_p + update_q, expl_leapfrog.hpp:16-26), staged by `nthr` threads:
// beta -> sbeta[0, nb_pad) (zero beyond K), the group intercepts -> sa (if staged), the whole state -> st (if staged),
// theta -> p.theta_used (CTA 0).  The caller synchronises afterwards.
__device__ __forceinline__ void stage_theta(const KernelParams& p, int t, int nthr, double* sbeta, int nb_pad,
                                            double* sa, const StateSmem& st) {
  const int P = p.P, K = p.K, G = p.G;
  const bool lf = p.mode == MODE_LEAPFROG;
  const double he = 0.5 * p.eps;
  for (int i = t; i < P; i += nthr) {
    double q, ph = 0.0, g0 = 0.0;
    if (lf) {
      g0 = p.st_in[2 * P + i];
      ph = p.st_in[P + i] - he * g0;
      q = p.st_in[i] + p.eps * (p.inv_metric[i] * ph);
    } else {
      q = p.theta_inline_n ? p.theta_inline[i] : p.theta_in[i];
    }
    if (i >= p.off_beta && i < p.off_beta + K) sbeta[i - p.off_beta] = q;
    if (sa && i >= 2 && i < 2 + G) sa[i - 2] = q;
    if (st.theta) {
      st.theta[i] = q;
      st.ph[i] = ph;
      st.g0[i] = g0;
    }
    if (blockIdx.x == 0) p.theta_used[i] = q;
  }
  for (int k = K + t; k < nb_pad; k += nthr) sbeta[k] = 0.0;
}

template <int FAMILY>
__device__ __forceinline__ void cross_cta_reduce_and_finish(const KernelParams& p, double* sh_scratch,
                                                            int* sh_is_last, const StateSmem& st,
                                                            double* scratch /* idle smem: >= warps x min(512, K + 34) doubles */) {
  const int tid = threadIdx.x, nt = blockDim.x, grid = gridDim.x;
  const int K = p.K, G = p.G, P = p.P;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int t = atomicAdd(p.ticket, 1u);
    *sh_is_last = (t == (unsigned int)(grid - 1));
  }
  __syncthreads();
  if (tid == 0) tl_stamp(p, blockIdx.x, 5);      // partial written, ticket taken
  if (!*sh_is_last) return;
  if (tid == 0) {
    tl_stamp(p, grid, 0);                        // last CTA: ticket won
    if (p.tl) p.tl[(size_t)grid * 16 + 7] = blockIdx.x;
  }

  __threadfence();
  if (tid == 0) tl_stamp(p, grid + 1, 0);        // acquire fence after the ticket
  // Sum of the grid's partial rows in a FIXED order (bitwise reproducible).  Warp w takes rows w, w + nw, ...; a
  // lane reads 8 consecutive bytes of a 256-byte stretch of the row, so every load instruction of a warp is one
  // fully used, aligned run of sectors (rows are padded to 128 bytes, partial_stride), and eight rows are in flight
  // per thread.  The per-warp column sums meet in shared memory (the TMA ring is idle by now: `scratch`) and are
  // folded in warp order.  This loop is on the critical path of every launch: ~13 us as one dependent chain of 148
  // loads; 4.7 us with lanes striding over rows (two half-used 128-byte segments per instruction: one SM's L2 port,
  // not latency, was the limit -- no gain from 8 -> 40 loads in flight, profiles/r2_timeline_*); ~1.5 us like this.
  double* lik = st.lik ? st.lik : p.lik;         // the sums go to the on-chip copy when there is one
  const int n_sums = fam_has_aux_sum(FAMILY) ? K + 3 : K + 2;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const int n_chunks = (n_sums + 31) >> 5;       // 32-column chunks of a row
  // measurement only: tl_repeat makes the sum run twice (same result), the first pass stamped on its own, to
  // separate what a cold pass costs (instruction fetch, first touch of the partial rows) from a warm one
  for (int rep = (p.tl && p.tl_repeat) ? 0 : 1; rep < 2; ++rep) {
    if (rep == 1 && p.tl && p.tl_repeat) {
      __syncthreads();
      if (tid == 0) tl_stamp(p, grid + 1, 5);    // end of the cold pass
    }
    constexpr int CB = 16;                       // chunks (of 32 columns) summed per block: bounds the scratch
    const int sstride = min(CB, n_chunks) * 32;  // doubles of scratch per warp
    for (int cb = 0; cb < n_chunks; cb += CB) {
      const int cb_n = min(CB, n_chunks - cb);
      if (cb > 0) __syncthreads();               // the previous block's scratch has been folded
      for (int c0 = cb; c0 < cb + cb_n; c0 += 4) {   // up to 4 chunks (128 columns) per sweep over the rows
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        const double* src = p.partials + c0 * 32 + lane;
        int r = warp;
        for (; r + 7 * nw < grid; r += 8 * nw) {      // 8 rows x 4 chunks = 32 loads in flight per thread
          double v[8][4];
#pragma unroll
          for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int q = 0; q < 4; ++q)
              v[u][q] = (c0 + q) * 32 + lane < n_sums ? __ldcg(src + (size_t)(r + u * nw) * p.pstride + q * 32) : 0.0;
#pragma unroll
          for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] += v[u][q];
        }
        for (; r + 3 * nw < grid; r += 4 * nw) {
          double v[4][4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int q = 0; q < 4; ++q)
              v[u][q] = (c0 + q) * 32 + lane < n_sums ? __ldcg(src + (size_t)(r + u * nw) * p.pstride + q * 32) : 0.0;
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] += v[u][q];
        }
        for (; r < grid; r += nw) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if ((c0 + q) * 32 + lane < n_sums) acc[q] += __ldcg(src + (size_t)r * p.pstride + q * 32);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (c0 + q < cb + cb_n) scratch[warp * sstride + (c0 + q - cb) * 32 + lane] = acc[q];
      }
      if (tid == 0 && cb == 0) tl_stamp(p, grid + 1, 1);   // this warp's rows of the first block loaded
      __syncthreads();
      for (int j = cb * 32 + tid; j < min(n_sums, (cb + cb_n) * 32); j += nt) {
        double v = 0.0;
        for (int w = 0; w < nw; ++w) v += scratch[w * sstride + (j - cb * 32)];
        if (j < K)
          lik[p.off_beta + j] = v;
        else if (j == K)
          lik[P] = v;
        else if (j == K + 2)
          lik[P + 1] = v;      // neg_binomial_2_log: sum of the per-row d/dphi terms
        else if (G == 0)
          lik[0] = v;
      }
    }
  }
  if (p.group_fused) {
    // per-group residual sums = the gradient wrt a[g] (rvalue(a, index_multi(group)) scatters r back, rvalue.hpp:154-172)
    for (int g = tid; g < G; g += nt) {
      const int4 m = __ldg(p.gmeta + g);
      double v = 0.0;
      if (m.x <= m.y) {
        v = __ldcg(p.gpart + m.z);
        for (int c = m.x + 1; c <= m.y; ++c) v += __ldcg(p.gpart + (size_t)c * p.Gcs);
      }
      lik[2 + g] = v;
    }
  }
  if (tid == 0) tl_stamp(p, grid + 1, 2);        // sums written
  if (tid == 0) *p.ticket = 0u;
  if (fam_has_scale(FAMILY) && tid == 0) lik[P - 1] = 0.0;  // sigma | phi entry is derived in finish()
  if (!fam_has_aux_sum(FAMILY) && tid == 32) lik[P + 1] = 0.0;
  if (G > 0 && tid < 2) lik[tid] = 0.0;
  if (!st.lik) __threadfence();                  // the sums live in global memory only without the on-chip copy
  __syncthreads();
  if (tid == 0) tl_stamp(p, grid, 1);            // sum of the grid's partial rows done
  if (p.peer_in_main && !peer_allreduce_lik(p, sh_is_last, lik)) {
    peer_timeout_result(p);
    return;
  }
  if (tid == 0) tl_stamp(p, grid, 2);            // peers' partials received and summed
  if (p.fuse_finish) {
    const FinishSrc src = {st.theta ? st.theta : p.theta_used, lik, st.ph, st.g0};
    finish_t<FAMILY>(p, sh_scratch, src);
  } else if (st.lik) {
    for (int j = tid; j < P + 2; j += nt) p.lik[j] = lik[j];   // the separate epilogue launch reads the global copy
  }
  if (tid == 0) tl_stamp(p, grid, 3);            // model epilogue / leapfrog tail written
}

// ------------------------------------------------------------------------------------------
// The main kernel
// ------------------------------------------------------------------------------------------
template <int FAMILY, int CPL>
__global__ void __launch_bounds__(NUM_THREADS, 1) glm_fused_kernel(const __grid_constant__ KernelParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int K = p.K, C = p.C, G = p.G, P = p.P;
  const int S = p.n_stages;
  const int tile_doubles = C * PANEL_ROWS;
  const int Kpad = (K + 3) & ~3;

  // carve shared memory
  double* tiles = reinterpret_cast<double*>(smem_raw);                  // S * tile_doubles
  double* sbeta = tiles + (size_t)S * tile_doubles;                     // Kpad
  double* sr = sbeta + Kpad;                                            // 8 * 32
  double* red = sr + NUM_CONSUMER_WARPS * 32;                           // 8 * (Kpad + 4)
  double* sa = red + NUM_CONSUMER_WARPS * (Kpad + 4);                   // G (optional)
  double* st_base = sa + (p.stage_a_in_smem ? ((G + 1) & ~1) : 0);      // chain state (optional)
  const StateSmem st = carve_state_smem(st_base, P, p.state_in_smem);
  double* sgw = st_base + (p.state_in_smem ? state_smem_doubles(P) : 0);  // 8 warps x Gcs group sums (fused group path)
  double* after_a = sgw + (p.group_fused ? NUM_CONSUMER_WARPS * p.Gcs : 0);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(after_a);            // S
  uint64_t* empty_bar = full_bar + S;                                   // S
  __shared__ double sh_scratch[64];
  __shared__ int sh_is_last;
  const bool GF = p.group_fused != 0;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  pdl_launch_dependents();
  if (tid == 0) {
    tl_stamp(p, blockIdx.x, 0);                  // CTA entry
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();
  // From here the TMA producer streams X (constant data) while the consumer warps wait for the previous launch
  // on this stream -- whose epilogue may still be running in one CTA -- before they read theta / the state.
  if (warp != NUM_CONSUMER_WARPS) pdl_grid_dependency_wait();
  if (tid == 0) tl_stamp(p, blockIdx.x, 1);      // previous launch on the stream complete (PDL wait over)

  // ---- theta for this launch (leapfrog: begin_update_p + update_q, expl_leapfrog.hpp:16-26) ----
  auto theta_at = [&](int i) -> double {
    if (p.mode == MODE_LEAPFROG) {
      const double ph = p.st_in[P + i] - (0.5 * p.eps) * p.st_in[2 * P + i];
      return p.st_in[i] + p.eps * (p.inv_metric[i] * ph);
    }
    return p.theta_inline_n ? p.theta_inline[i] : p.theta_in[i];
  };
  constexpr int NC = NUM_CONSUMER_WARPS * 32;   // the consumer warps stage theta; the producer is already loading
  double alpha = 0.0;
  LinkConst lc;
  if (warp != NUM_CONSUMER_WARPS) {
    stage_theta(p, tid, NC, sbeta, Kpad, p.stage_a_in_smem ? sa : nullptr, st);
    alpha = G > 0 ? 0.0 : theta_at(0);
  }
  lc.inv_sigma = 1.0;
  lc.phi = 1.0;
  lc.log_phi = 0.0;
  lc.dg_phi = 0.0;
  lc.lg_phi = 0.0;
  lc.inc_phi_terms = (!p.mc.propto || !p.mc.lik_only || p.mc.sigma_is_var) ? 1 : 0;
  lc.inc_ytheta = (!p.mc.propto || !p.mc.lik_only || p.mc.sigma_is_var != 2) ? 1 : 0;
  if (warp != NUM_CONSUMER_WARPS) {
    if (FAMILY == FAM_NORMAL_ID) lc.inv_sigma = 1.0 / exp(theta_at(P - 1));  // normal_id_glm_lpdf.hpp:117
    if (FAMILY == FAM_NEG_BINOMIAL_2_LOG) {
      lc.phi = exp(theta_at(P - 1));           // lb_constrain.hpp:65
      lc.log_phi = log(lc.phi);                // neg_binomial_2_log_glm_lpmf.hpp:152
      lc.dg_phi = digamma_pos(lc.phi);
      lc.lg_phi = lgamma(lc.phi);
    }
    named_barrier_sync(1, NC);                 // sbeta / sa visible to every consumer warp
  }
  if (tid == 0) tl_stamp(p, blockIdx.x, 2);      // theta staged

  const long long n_panels = p.n_panels;
  const int grid = gridDim.x;
  // this CTA's panels: n = 0 .. p_count-1 -> panel p_base + n * p_stride.  Interleaved over the grid by default; one
  // contiguous range per CTA on the fused group path (rows are sorted by group, so a range meets few groups)
  const long long p_base = GF ? ((long long)blockIdx.x * n_panels) / grid : (long long)blockIdx.x;
  const long long p_stride = GF ? 1 : grid;
  const long long p_count = GF ? ((long long)(blockIdx.x + 1) * n_panels) / grid - p_base
                               : (n_panels > blockIdx.x ? (n_panels - blockIdx.x + grid - 1) / grid : 0);

  double acc[CPL];
#pragma unroll
  for (int s = 0; s < CPL; ++s) acc[s] = 0.0;
  double lp_acc = 0.0, r_acc = 0.0, x_acc = 0.0;

  if (warp == NUM_CONSUMER_WARPS) {
    // ===================== TMA producer (one elected lane) =====================
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      const uint32_t bytes = (uint32_t)tile_doubles * 8u;
      int s = 0;
      uint32_t round = 0;
      bool dep_waited = false;
      for (long long n = 0; n < p_count; ++n) {
        const long long pi = p_base + n * p_stride;
        // Under programmatic dependent launch this CTA may have started while the PREVIOUS launch's last CTA is
        // still in its serial tail (grid sum, exchange, epilogue).  Filling the whole ring right away floods the
        // memory system with ~30 MB of bulk loads and that tail's few small loads queue behind them (measured:
        // 5.2 us for 2 rounds of L2 loads, profiles/r2_timeline_*); so only pdl_prefetch stages are requested
        // before the previous launch has completed, the rest of the ring after.
        if (n == p.pdl_prefetch && !dep_waited) {
          pdl_grid_dependency_wait();
          dep_waited = true;
        }
        if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        tma_load_1d(tiles + (size_t)s * tile_doubles, p.panels + (size_t)pi * tile_doubles, bytes,
                    &full_bar[s], pol);
        if (++s == S) {
          s = 0;
          ++round;
        }
      }
    }
    pdl_grid_dependency_wait();   // the tail below writes what the previous launch's epilogue may still be reading
  } else if (warp < (S < NUM_CONSUMER_WARPS ? S : NUM_CONSUMER_WARPS)) {
    // ===================== consumers =====================
    // Panel n of the CTA's sequence lives in stage n % S; a STAGE belongs to one warp (stage s to warp s % 8) for the
    // whole launch, so S need not be a multiple of 8 (with 13 stages five warps own two and three own one).  The
    // ownership is what makes the parity wait sound: a warp has consumed fill k - 1 of a stage before it waits for
    // fill k, so the barrier can only be in phase k or k + 1.  (Handing panels out round-robin, n % 8, with S = 9 let
    // a warp wait for fill 1 of a stage whose fill 0 was still in flight -- parity 1 reads as "the phase before phase
    // 0 is complete", the warp ran ahead on stale data and the ring dead-locked.)
    const int rg = lane & 3, cg = lane >> 2, cgl = cg & 3;
    const int o1 = lane ^ 4, o2 = lane ^ 8, o3 = lane ^ 12;
    const int ycol = K * 32 + (lane ^ ((K & 3) << 2));
    const int tcol = (K + 1) * 32 + (lane ^ (((K + 1) & 3) << 2));   // binomial population sizes
    const int Kg = fam_group_col(FAMILY, K);
    const int gcol = Kg * 32 + (lane ^ ((Kg & 3) << 2));
    int off[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) off[m] = rg + 4 * (m ^ cgl);
    double* my_sr = sr + warp * 32;

    // fused group path: running residual sum of the group this warp is in (flushed to the warp's smem row when the
    // group changes: the warp meets groups in ascending order, so once per group)
    const int g_base = GF ? p.cta_g0[blockIdx.x] : 0;
    double* my_sg = sgw + warp * p.Gcs;
    int cur_g = -1;
    double rg_acc = 0.0;
    auto flush_group = [&]() {
      const double v = warp_sum(rg_acc);
      if (lane == 0 && cur_g >= 0) my_sg[cur_g - g_base] += v;
      rg_acc = 0.0;
    };
    if (GF) {
      for (int j = lane; j < p.Gcs; j += 32) my_sg[j] = 0.0;
      __syncwarp();
    }

    uint32_t parity = 0;
    for (long long n0 = 0; n0 < p_count; n0 += S, parity ^= 1u)          // fill round k = n0 / S of the ring
    for (int s = warp; s < S; s += NUM_CONSUMER_WARPS) {                 // this warp's stages, in panel order
      const long long n = n0 + s;                                        // index of the panel in the CTA's sequence
      if (n >= p_count) break;
      const long long pi = p_base + n * p_stride;
      mbar_wait(&full_bar[s], parity);
      if (n == 0 && tid == 0) tl_stamp(p, blockIdx.x, 3);   // first panel landed
      const double* tile = tiles + (size_t)s * tile_doubles;

      // ---- phase 1: eta for row `lane` ----
      double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
      int c = 0;
#pragma unroll 4
      for (; c + 4 <= K; c += 4) {
        const double2 b01 = *reinterpret_cast<const double2*>(sbeta + c);
        const double2 b23 = *reinterpret_cast<const double2*>(sbeta + c + 2);
        e0 = fma(tile[(c + 0) * 32 + lane], b01.x, e0);
        e1 = fma(tile[(c + 1) * 32 + o1], b01.y, e1);
        e2 = fma(tile[(c + 2) * 32 + o2], b23.x, e2);
        e3 = fma(tile[(c + 3) * 32 + o3], b23.y, e3);
      }
      for (; c < K; ++c) e0 = fma(tile[c * 32 + (lane ^ ((c & 3) << 2))], sbeta[c], e0);
      double eta = (e0 + e1) + (e2 + e3);
      const double y = tile[ycol];
      int gi = -1;
      if (G > 0) {
        gi = (int)tile[gcol] - 1;
        const bool gok = gi >= 0 && gi < G;
        eta += p.stage_a_in_smem ? sa[gok ? gi : 0] : theta_at(2 + (gok ? gi : 0));
      } else {
        eta += alpha;
      }
      const bool valid = (pi * PANEL_ROWS + lane) < p.n_rows;
      double lp_i, r_i, x_i;
      if (p.alpha_rows || p.sigma_rows) {
        // function-level call with per-row operands (G == 0: rows are in the caller's order)
        const long long row = pi * PANEL_ROWS + lane;
        if (p.alpha_rows && valid) eta += __ldg(p.alpha_rows + row);
        LinkConst lcr = lc;
        double sg = 1.0;
        if (FAMILY == FAM_NORMAL_ID && p.sigma_rows) {
          sg = valid ? __ldg(p.sigma_rows + row) : 1.0;
          lcr.inv_sigma = 1.0 / sg;
        }
        link_ext<FAMILY>(eta, y, FAMILY == FAM_BINOMIAL_LOGIT ? tile[tcol] : 0.0, lcr, lp_i, r_i, x_i);
        if (FAMILY == FAM_NORMAL_ID && p.sigma_rows) {
          x_i = log(sg);                                              // sum log sigma_i rides in the aux sum
          if (valid) p.s_out[row] = (lp_i - 1.0) * lcr.inv_sigma;    // normal_id_glm_lpdf.hpp:181-183, per row
        }
        if (p.alpha_rows && valid) p.r_out[row] = r_i;
      } else {
        link_ext<FAMILY>(eta, y, FAMILY == FAM_BINOMIAL_LOGIT ? tile[tcol] : 0.0, lc, lp_i, r_i, x_i);
      }
      if (!valid) {
        lp_i = 0.0;
        r_i = 0.0;
        x_i = 0.0;
      }
      lp_acc += lp_i;
      r_acc += r_i;
      if (fam_has_aux_sum(FAMILY)) x_acc += x_i;
      if (GF) {
        const int g_first = __shfl_sync(0xffffffffu, gi, 0);          // row 0 of a panel always exists
        if (__all_sync(0xffffffffu, !valid || gi == g_first)) {
          if (g_first != cur_g) {
            flush_group();
            cur_g = g_first;
          }
          rg_acc += r_i;
        } else {
          // a panel that straddles group boundaries (rare: rows are sorted): one masked warp sum per group, in row order
          unsigned todo = __ballot_sync(0xffffffffu, valid);
          while (todo) {
            const int gsel = __shfl_sync(0xffffffffu, gi, __ffs(todo) - 1);
            const bool mine = valid && gi == gsel;
            const double v = warp_sum(mine ? r_i : 0.0);
            if (lane == 0) my_sg[gsel - g_base] += v;
            todo &= ~__ballot_sync(0xffffffffu, mine);
          }
        }
      } else if (G > 0) {
        p.r_out[pi * PANEL_ROWS + lane] = r_i;
      }

      // ---- phase 2: X^T r from the same smem tile ----
      my_sr[lane] = r_i;
      __syncwarp();
      double rr[8];
#pragma unroll
      for (int m = 0; m < 8; ++m) rr[m] = my_sr[rg + 4 * m];
      const double* base = tile + cg * 32;
#pragma unroll
      for (int s2 = 0; s2 < CPL; ++s2) {
        if (s2 * 8 < K) {
          if (cg + 8 * s2 < K) {
#pragma unroll
            for (int m = 0; m < 8; ++m) acc[s2] = fma(base[s2 * 256 + off[m]], rr[m], acc[s2]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

    if (GF) flush_group();
    // ---- per-warp reduction of the private accumulators ----
#pragma unroll
    for (int s2 = 0; s2 < CPL; ++s2) {
      double v = acc[s2];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (rg == 0 && cg + 8 * s2 < K) red[warp * (Kpad + 4) + cg + 8 * s2] = v;
    }
    lp_acc = warp_sum(lp_acc);
    r_acc = warp_sum(r_acc);
    if (fam_has_aux_sum(FAMILY)) x_acc = warp_sum(x_acc);
    if (lane == 0) {
      red[warp * (Kpad + 4) + Kpad] = lp_acc;
      red[warp * (Kpad + 4) + Kpad + 1] = r_acc;
      red[warp * (Kpad + 4) + Kpad + 2] = x_acc;
    }
  } else {
    // idle consumer warp (fewer stages than warps): contributes zeros to the CTA reduction
    for (int j = lane; j < Kpad + 4; j += 32) red[warp * (Kpad + 4) + j] = 0.0;
    if (GF)
      for (int j = lane; j < p.Gcs; j += 32) sgw[warp * p.Gcs + j] = 0.0;
  }
  __syncthreads();
  if (tid == 0) tl_stamp(p, blockIdx.x, 4);      // every warp has consumed its last panel

  // ---- CTA partial -> global ----
  double* my_part = p.partials + (size_t)blockIdx.x * p.pstride;
  for (int j = tid; j < K + 3; j += NUM_THREADS) {
    const int src = j < K ? j : Kpad + (j - K);
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < NUM_CONSUMER_WARPS; ++w) v += red[w * (Kpad + 4) + src];
    my_part[j] = v;
  }
  if (GF) {   // this CTA's group sums, warps folded in fixed order
    double* my_g = p.gpart + (size_t)blockIdx.x * p.Gcs;
    for (int j = tid; j < p.Gcs; j += NUM_THREADS) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < NUM_CONSUMER_WARPS; ++w) v += sgw[w * p.Gcs + j];
      my_g[j] = v;
    }
  }
  cross_cta_reduce_and_finish<FAMILY>(p, sh_scratch, &sh_is_last, st, tiles);
}

// G > 0: deterministic per-group sums of the residual (rows are sorted by group at upload, so a group
// is one contiguous segment).  One CTA per group; four independent accumulators per thread keep four
// loads in flight, and the order of every addition is fixed -> bitwise reproducible.
__global__ void __launch_bounds__(256) group_reduce_kernel(const double* __restrict__ r,
                                                          const long long* __restrict__ seg_ptr, int G,
                                                          double* __restrict__ lik_a /* lik + 2 */) {
  __shared__ double sh[8];
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    const long long b = seg_ptr[g], e = seg_ptr[g + 1];
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
    long long i = b + threadIdx.x;
    for (; i + 768 < e; i += 1024) {
      v0 += __ldcs(r + i);
      v1 += __ldcs(r + i + 256);
      v2 += __ldcs(r + i + 512);
      v3 += __ldcs(r + i + 768);
    }
    for (; i < e; i += 256) v0 += __ldcs(r + i);
    double v = warp_sum((v0 + v1) + (v2 + v3));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += sh[w];
      lik_a[g] = t;
    }
  }
}

// Leapfrog step of a model whose likelihood term is empty (no rows anywhere, or the binomial size_zero quirk):
// begin_update_p + update_q (expl_leapfrog.hpp:16-26) without a pass over X; finish_kernel does the rest.
__global__ void __launch_bounds__(NUM_THREADS) theta_from_state_kernel(const __grid_constant__ KernelParams p) {
  for (int i = threadIdx.x; i < p.P; i += blockDim.x) {
    const double ph = p.st_in[p.P + i] - (0.5 * p.eps) * p.st_in[2 * p.P + i];
    p.theta_used[i] = p.st_in[i] + p.eps * (p.inv_metric[i] * ph);
  }
}

__global__ void __launch_bounds__(NUM_THREADS) finish_kernel(const __grid_constant__ KernelParams p) {
  __shared__ double sh_scratch[64];
  __shared__ int sh_ok;
  if (p.peer_in_finish && !peer_allreduce_lik(p, &sh_ok, p.lik)) {
    peer_timeout_result(p);
    return;
  }
  finish(p, sh_scratch);
}

// class-outcome models whose likelihood term is empty (no rows anywhere, one class): the epilogue alone
__global__ void __launch_bounds__(NUM_THREADS) class_finish_kernel(const __grid_constant__ KernelParams p) {
  __shared__ double sh_scratch[64];
  class_epilogue(p, sh_scratch);
}

// ------------------------------------------------------------------------------------------
// One-time re-layout: column-major X (+ y, group) -> row-panel format (optionally through a row
// permutation that sorts rows by group).  One thread per (panel, column, row).
//   narrow format: PR = 32 rows per panel, Cs = C columns, XOR swizzle (swz = 1)
//   wide format:   PR = 16 rows per panel, Cs = Cpad columns (zero padded), no swizzle
// ------------------------------------------------------------------------------------------
__global__ void relayout_kernel(const double* __restrict__ X, long long ldx, long long x_row0,
                                const int32_t* __restrict__ y_int, const double* __restrict__ y_real,
                                const int32_t* __restrict__ group, const int32_t* __restrict__ trials,
                                const long long* __restrict__ perm, long long row0, long long n_rows_chunk, long long n_rows_total, int K, int C,
                                int c_begin, int c_end, double* __restrict__ panels, int PR, int Cs, int swz) {
  // Writes columns [c_begin, c_end) of destination rows [row0, row0 + n_rows_chunk), row0 % PR == 0.
  // X may be a chunk whose first row is source row x_row0 (leading dimension ldx).
  const long long n_panels_chunk = (n_rows_chunk + PR - 1) / PR;
  const int nc = c_end - c_begin;
  const long long total = n_panels_chunk * nc * PR;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % PR);
    const long long t = idx / PR;
    const int c = c_begin + (int)(t % nc);
    const long long pl = t / nc;
    const long long dst_row = row0 + pl * PR + r;
    double v = 0.0;
    if (dst_row < n_rows_total && dst_row < row0 + n_rows_chunk && c < C) {
      const long long src = perm ? perm[dst_row] : dst_row;
      if (c < K)
        v = X[(src - x_row0) + (long long)c * ldx];
      else if (c == K)
        v = y_real ? y_real[src] : (double)y_int[src];
      else if (trials && c == K + 1)
        v = (double)trials[src];
      else
        v = (double)group[src];
    }
    const int rr = swz ? (r ^ ((c & 3) << 2)) : r;
    panels[(row0 / PR + pl) * (long long)Cs * PR + (long long)c * PR + rr] = v;
  }
}

// data checks the reference performs on every call (bernoulli :85 check_bounded, poisson :84
// check_nonnegative) + the propto=false constant sum lgamma(y+1) (poisson :127-129)
// binomial_coefficient_log.hpp:81-112 (value; the symmetric branch keeps k <= n/2)
__device__ __forceinline__ double binomial_coefficient_log_d(double n, double k) {
  if (k > n / 2.0 + 1e-8) k = n - k;
  if (k == 0.0) return 0.0;
  return lgamma(n + 1.0) - lgamma(k + 1.0) - lgamma(n + 1.0 - k);
}
// + binomial_logit :101-102 check_bounded(n, 0, N) / check_nonnegative(N) and :127-130 the propto=false
// constant sum binomial_coefficient_log(N, n), stored NEGATED (finish() subtracts the constant);
// neg_binomial_2_log :130 check_nonnegative(y), :163-169 sum lgamma(y+1)
__global__ void __launch_bounds__(256) y_stats_kernel(const int32_t* __restrict__ y, const int32_t* __restrict__ trials,
                                                      long long n, int family,
                                                      double* out /* [gridDim.x][2] = bad, lgamma_sum */, int n_classes) {
  __shared__ double sh[16];
  double bad = 0.0, lg = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int v = y[i];
    if (family == FAM_BERNOULLI_LOGIT) {
      if (v < 0 || v > 1) bad += 1.0;
    } else if (fam_is_class(family)) {        // check_bounded(y, 1, N_classes)
      if (v < 1 || v > n_classes) bad += 1.0;
    } else if (family == FAM_BINOMIAL_LOGIT) {
      const int t = trials[i];
      if (v < 0 || v > t || t < 0) bad += 1.0;
      else lg -= binomial_coefficient_log_d((double)t, (double)v);
    } else {
      if (v < 0) bad += 1.0;
      else lg += lgamma((double)v + 1.0);
    }
  }
  bad = warp_sum(bad);
  lg = warp_sum(lg);
  if ((threadIdx.x & 31) == 0) {
    sh[threadIdx.x >> 5] = bad;
    sh[8 + (threadIdx.x >> 5)] = lg;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = 0.0, l = 0.0;
    for (int w = 0; w < 8; ++w) {
      b += sh[w];
      l += sh[8 + w];
    }
    out[2 * blockIdx.x] = b;
    out[2 * blockIdx.x + 1] = l;
  }
}

}  // namespace b200glm
