// glm_model.cuh -- what a launch is told (KernelParams) and the MODEL EPILOGUE finish(): priors, lb_constrain
// Jacobians, assembly of the gradient with respect to the unconstrained parameters, status, and the second half
// of the leapfrog step.  Run by ONE CTA of the gradient kernel (or by finish_kernel) once the likelihood sums are
// complete.  Like glm_link.cuh it is kept free of anything a host compiler cannot take: tests/host/model_host.cpp
// compiles it with g++ as a one-thread CTA (threadIdx.x = 0, blockDim.x = 1, warp_sum = identity) and
// tests/test_model_epilogue_host.py checks finish() -- fed with likelihood sums built from glm_link.cuh on the
// host -- against the CPU oracle: the whole arithmetic of the device path, minus the parallel reductions, has a
// regression test that needs no GPU.
#pragma once

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#include <math_constants.h>
#endif
#include <stdint.h>

#include "glm_link.cuh"

namespace b200glm {

#if !defined(__CUDACC__)
struct int4 {   // the host build (tests/host) has no CUDA vector types
  int x, y, z, w;
};
#endif

enum { MODE_THETA = 0, MODE_LEAPFROG = 1 };
enum { ST_OK = 0, ST_DOMAIN = 1, ST_PEER_TIMEOUT = 2 };

#define NEG_LOG_SQRT_TWO_PI_D (-0.91893853320467274178032973640562)

// Everything the epilogue needs to turn likelihood sums into the model's lp / gradient.
struct ModelConst {
  int family, K, G, P, off_beta;
  int propto, jacobian, is_var;  // semantics of this evaluation (see include/b200glm.h)
  int lik_only;                  // function-level call (b200glm_glm_lpmf): the GLM term alone -- no priors, no
                                 // Jacobian; the sigma entry of the gradient is d/d sigma, not d/d log sigma
  int sigma_rows;                // lik_only + normal_id with a per-row scale: lik[P + 1] holds sum log sigma_i
  int sigma_is_var;              // lik_only + normal_id: keep -N log sigma under propto (normal_id_glm_lpdf.hpp:205);
                                 // lik_only + neg_binomial_2_log: phi is an autodiff variable (the phi-only terms
                                 // stay); 2 = phi is the ONLY variable operand (y * theta drops under propto)
  double N_total;                // rows over all shards
  double lgamma_sum;             // propto=0 constant over all shards, subtracted: sum lgamma(y+1) (poisson,
                                 // neg_binomial_2_log), -sum binomial_coefficient_log(trials, y) (binomial_logit)
  double prior_alpha_sd, prior_beta_sd, prior_sigma_loc, prior_sigma_scale, prior_sigma_a_scale;
};

// Row-sharded operation without a separate collective launch: every rank owns a mailbox in its HBM,
// mapped into every peer (CUDA IPC over NVLink / NVSwitch).  [PEER_BUFS][world][stride] doubles.
constexpr int MAX_PEERS = 8;
constexpr int THETA_INLINE_MAX = 256;
constexpr int PEER_BUFS = 3;       // mailbox buffers rotated by evaluation number (see peer_allreduce_lik)
// Every 8-byte word of a mailbox is self-validating: it holds PEER_EMPTY until the sender's value lands.
// PEER_EMPTY is a NaN bit pattern no arithmetic produces; a payload that happens to carry it (a NaN handed in
// by the caller) is sent as the canonical quiet NaN instead.
#define PEER_EMPTY_BITS 0xFFFFFFFFFFFFFFFFull
struct PeerParams {
  int enabled, world, rank, stride;
  unsigned long long seq;          // 1, 2, 3, ... identical on every rank for the same evaluation
  unsigned long long timeout_ns;
  double* mbox[MAX_PEERS];         // this slot's mailbox on rank r (mbox[rank] is local memory)
};

struct KernelParams {
  const double* panels;
  long long n_rows;    // local rows
  long long n_panels;  // local panels
  int K, C, G, family, P, off_beta;
  int n_stages;        // narrow kernel: panel stages; wide kernel: sub-panel slots in the ring
  int Cpad, Kc, J;     // wide kernel: padded column count, sub-panel width, sub-panels per row panel
  int mode;
  int fuse_finish;     // last CTA also runs finish() (G == 0 and: world == 1, or peers connected)
  int peer_in_main;    // last CTA of the main kernel exchanges the likelihood partials with the peers
  int peer_in_finish;  // finish_kernel does (G > 0: the group sums are only complete by then)
  PeerParams peer;
  int stage_a_in_smem;
  int state_in_smem;   // the launch keeps (theta, half-updated momentum, old gradient, likelihood sums) in shared memory
  const double* theta_in;   // MODE_THETA: P doubles (device)
  const double* st_in;      // MODE_LEAPFROG: [q(P) p(P) g(P) V]
  double* st_out;
  const double* inv_metric; // P doubles
  double eps;
  double* partials;         // [grid][pstride]: [0,K) beta grads, [K] lp-sum, [K+1] r-sum
  int pstride;
  unsigned int* ticket;
  double* r_out;            // G > 0, unfused group path: residual per (sorted) row
  // G > 0, fused group path (narrow kernel): rows are sorted by group and every CTA owns a CONTIGUOUS panel range,
  // so a CTA meets few groups: it accumulates their residual sums on chip and writes them to gpart[cta][0, Gcs);
  // the last CTA folds them in CTA order through gmeta[g] = {first CTA, last CTA, index of the first CTA's entry}
  // (in every later CTA the group is that CTA's first: entry 0).  cta_g0[c] = first group (0-based) of CTA c's range.
  // function-level entry with per-row operands (b200glm_glm_lpmf_rows; narrow kernel, G == 0): intercept alpha_i and /
  // or scale sigma_i per row, in the caller's row order; r_out then receives d logp / d alpha_i (the residual), s_out
  // d logp / d sigma_i = (z_i^2 - 1) / sigma_i   (SM/opencl/prim/normal_id_glm_lpdf.hpp:68-84 and :117-150)
  const double* alpha_rows;
  const double* sigma_rows;
  double* s_out;
  int n_classes;            // class-outcome models (ordered_logistic, categorical_logit): number of classes C
  double* cuts;             // ordered_logistic: 2 (C - 1) doubles of scratch for the epilogue
  int group_fused, Gcs;
  double* gpart;
  const int4* gmeta;
  const int* cta_g0;
  double* lik;              // [P] likelihood gradient aligned with theta, [P] lp-sum, [P+1] spare
  double* result;           // [lp, grad(P), status]
  double* theta_used;       // P doubles: the theta this launch evaluated (q_new in leapfrog mode)
  int pdl_prefetch;         // ring stages the TMA producer may request before the previous launch has completed
  int tl_repeat;            // measurement only, see cross_cta_reduce_and_finish
  int wide_producer;        // wide kernel: 0 = one lane issues every bulk copy, 1 = lane j streams sub-panel j
  unsigned long long* tl;   // NULL, or the per-phase time stamps of this launch (b200glm_timeline_*): [grid + 1][16]
  // Host-facing calls (b200glm_log_prob_grad, b200glm_leapfrog): the epilogue also writes its outputs straight into
  // pinned host memory -- [result (P + 2)] [state (3P + 1)] [sequence word] -- and the host polls the sequence word,
  // instead of a device-to-host copy plus a stream synchronisation behind the launch (NULL for device-resident callers).
  double* host_out;
  unsigned long long host_seq;
  // MODE_THETA with a small model: theta travels in the kernel's parameter block instead of a host-to-device copy
  // ahead of the launch (theta_inline_n = P, or 0: read theta_in)
  int theta_inline_n;
  ModelConst mc;
  double theta_inline[THETA_INLINE_MAX];
};

#if defined(__CUDACC__)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif   // the host build supplies a one-lane warp_sum

// ------------------------------------------------------------------------------------------
// Model epilogue, run by ONE CTA once the likelihood sums are complete.
//   src.theta : the evaluated point            src.lik : [P] likelihood gradient aligned with theta, [P] lp-sum,
//   src.ph    : leapfrog mode, the momentum after begin_update_p (NULL: recomputed from p.st_in)
//   src.g0    : leapfrog mode, the gradient the step started from (NULL: p.st_in + 2P)
// The fused kernels hand in shared-memory copies (the state was read when theta was staged, the sums were formed
// by this CTA), so the epilogue on the critical path of every launch has no dependent global-memory round trip;
// finish_kernel and the host build pass the global arrays.
// ------------------------------------------------------------------------------------------
// All of the CTA's writes to the pinned host block are done: make them visible system-wide, then release the
// sequence word the host polls.  Called by every thread of the CTA.
__device__ inline void host_out_publish(const KernelParams& p) {
#if defined(__CUDACC__)
  if (p.host_out) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(p.host_out + (p.P + 2) + (3 * p.P + 1));
      *flag = p.host_seq;
    }
  }
#else
  (void)p;
#endif
}

// measurement only (b200glm_timeline_*): stamp k of the extra tail row
__device__ inline void finish_stamp(const KernelParams& p, int k) {
#if defined(__CUDACC__)
  if (p.tl && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.tl[(size_t)(gridDim.x + 1) * 16 + k] = t;
  }
#else
  (void)p;
  (void)k;
#endif
}

struct FinishSrc {
  const double* theta;
  const double* lik;
  const double* ph;
  const double* g0;
};

// FAMILY >= 0: the family is a compile-time constant (the fused kernels: only that family's branch is compiled in, which
// keeps the once-per-launch, cold-instruction-cache epilogue short); FAMILY < 0: taken from p.mc.family at run time.
template <int FAMILY>
__device__ void finish_t(const KernelParams& p, double* sh /* >= 64 doubles scratch */, const FinishSrc& src) {
  const ModelConst& mc = p.mc;
  const int family = FAMILY >= 0 ? FAMILY : mc.family;
  const int P = mc.P, K = mc.K, G = mc.G;
  const double* theta = src.theta;
  const double* lik = src.lik;
  const int tid = threadIdx.x, nt = blockDim.x;
  const bool dens = (!mc.propto) || mc.is_var;  // include_summand: anything left to compute?

  const double alpha = G > 0 ? 0.0 : theta[0];
  const double mu_a = G > 0 ? theta[0] : 0.0;
  const double u_sa = G > 0 ? theta[1] : 0.0;
  const double sigma_a = G > 0 ? exp(u_sa) : 1.0;
  const bool has_scale = fam_has_scale(family);   // sigma (normal_id) | phi (neg_binomial_2_log)
  const double u_s = has_scale ? theta[P - 1] : 0.0;
  const double sigma = has_scale ? exp(u_s) : 1.0;
  const double ib2 = 1.0 / (mc.prior_beta_sd * mc.prior_beta_sd);
  const double isa2 = 1.0 / (sigma_a * sigma_a);

  // block-wide sums: sum beta^2, sum (a-mu)^2, sum (a-mu), non-finite count
  double sb = 0.0, sa = 0.0, sd = 0.0, bad = 0.0;
  for (int k = tid; k < K; k += nt) {
    const double b = theta[mc.off_beta + k];
    sb += b * b;
  }
  for (int g = tid; g < G; g += nt) {
    const double d = theta[2 + g] - mu_a;
    sa += d * d;
    sd += d;
  }
  for (int i = tid; i < P; i += nt) {
    if (!isfinite(theta[i]) || !isfinite(lik[i])) bad += 1.0;
  }
  sb = warp_sum(sb);
  sa = warp_sum(sa);
  sd = warp_sum(sd);
  bad = warp_sum(bad);
  __syncthreads();
  const int w = tid >> 5, nw = (nt + 31) >> 5;
  if ((tid & 31) == 0) {
    sh[w] = sb;
    sh[12 + w] = sa;
    sh[24 + w] = bad;
    sh[36 + w] = sd;
  }
  __syncthreads();
  double sum_b2 = 0.0, sum_d2 = 0.0, n_bad = 0.0, sum_d = 0.0;   // every thread folds the (<= 12) warp sums itself
  for (int i = 0; i < nw; ++i) {
    sum_b2 += sh[i];
    sum_d2 += sh[12 + i];
    n_bad += sh[24 + i];
    sum_d += sh[36 + i];
  }

  // ---- value (computed redundantly by every thread: no broadcast round) ----
  double lp = 0.0;
  {
    if (mc.jacobian && !mc.lik_only) {
      if (G > 0) lp += u_sa;                           // lb_constrain.hpp:64
      if (has_scale) lp += u_s;
    }
    if (dens && !mc.lik_only) {
      // priors: normal_lpdf.hpp:81-88
      if (G > 0) {
        const double z0 = mu_a / mc.prior_alpha_sd;
        lp += -0.5 * z0 * z0;
        const double z1 = sigma_a / mc.prior_sigma_a_scale;
        lp += -0.5 * z1 * z1;
        lp += -0.5 * sum_d2 * isa2 - G * u_sa;         // -G log sigma_a (sigma_a is a parameter)
        if (!mc.propto)
          lp += 2.0 * NEG_LOG_SQRT_TWO_PI_D - log(mc.prior_alpha_sd) - log(mc.prior_sigma_a_scale)
                + G * NEG_LOG_SQRT_TWO_PI_D;
      } else {
        const double z0 = alpha / mc.prior_alpha_sd;
        lp += -0.5 * z0 * z0;
        if (!mc.propto) lp += NEG_LOG_SQRT_TWO_PI_D - log(mc.prior_alpha_sd);
      }
      if (K > 0) {
        lp += -0.5 * sum_b2 * ib2;
        if (!mc.propto) lp += K * (NEG_LOG_SQRT_TWO_PI_D - log(mc.prior_beta_sd));
      }
      if (has_scale) {
        const double z = (sigma - mc.prior_sigma_loc) / mc.prior_sigma_scale;
        lp += -0.5 * z * z;
        if (!mc.propto) lp += NEG_LOG_SQRT_TWO_PI_D - log(mc.prior_sigma_scale);
      }
    }
    if (dens) {
      // likelihood
      if (mc.N_total > 0) {
        const double S = lik[P];
        if (family == FAM_BERNOULLI_LOGIT) {
          lp += S;
        } else if (family == FAM_POISSON_LOG || family == FAM_BINOMIAL_LOGIT) {
          lp += S;
          if (!mc.propto) lp -= mc.lgamma_sum;         // poisson_log_glm_lpmf.hpp:127-129, binomial_logit_glm_lpmf.hpp:127-130
        } else if (family == FAM_NEG_BINOMIAL_2_LOG) {
          lp += S;                                     // neg_binomial_2_log_glm_lpmf.hpp:186-197 (row terms)
          if (!mc.propto) lp -= mc.lgamma_sum;         // :163-169
          if (!mc.lik_only || !mc.propto || mc.sigma_is_var)
            lp += mc.N_total * (sigma * log(sigma) - lgamma(sigma));   // :170-185 multiply_log(phi, phi) - lgamma(phi)
        } else {
          if (!mc.propto) lp += NEG_LOG_SQRT_TWO_PI_D * mc.N_total;   // normal_id_glm_lpdf.hpp:202-204
          if (!mc.lik_only || !mc.propto || mc.sigma_is_var)
            lp -= mc.sigma_rows ? lik[P + 1] : mc.N_total * u_s;      // :205-212, log sigma = u_s | sum log sigma_i
          lp -= 0.5 * S;                                              // :213
        }
      }
    }
  }
  const bool domain = !(isfinite(lp) && n_bad == 0.0);
  finish_stamp(p, 3);                             // block sums + value done

  // ---- gradient wrt unconstrained theta and, in leapfrog mode, the second half of the step
  //      (expl_leapfrog.hpp:28-32 end_update_p; base_hamiltonian.hpp:64-69), one thread per entry ----
  const bool lf = p.mode == MODE_LEAPFROG;
  const double he = 0.5 * p.eps;
  double* hres = p.host_out;                                   // pinned host mirror of result / state (or NULL)
  double* hst = p.host_out ? p.host_out + (P + 2) : nullptr;
  for (int i = tid; i < P; i += nt) {
    double g = lik[i];
    if (mc.lik_only) {
      if (family == FAM_NORMAL_ID && i == P - 1)
        g = mc.N_total > 0 ? (lik[P] - mc.N_total) / sigma : 0.0;    // normal_id_glm_lpdf.hpp:181-183
      else if (family == FAM_NEG_BINOMIAL_2_LOG && i == P - 1)
        g = mc.N_total > 0 ? lik[P + 1] : 0.0;                       // neg_binomial_2_log_glm_lpmf.hpp:240-245
      else if (G > 0 && i < 2)
        g = 0.0;
    } else {
      if (G > 0) {
        if (i == 0) {
          g = -mu_a / (mc.prior_alpha_sd * mc.prior_alpha_sd) + 0.0;
          g += sum_d * isa2;                                          // sum_g (a_g - mu) / sigma_a^2
        } else if (i == 1) {
          const double dsa = -sigma_a / (mc.prior_sigma_a_scale * mc.prior_sigma_a_scale)
                             + sum_d2 * isa2 / sigma_a - G / sigma_a;
          g = dsa * sigma_a + (mc.jacobian ? 1.0 : 0.0);
        } else if (i < 2 + G) {
          g += -(theta[i] - mu_a) * isa2;
        }
      } else if (i == 0) {
        g += -alpha / (mc.prior_alpha_sd * mc.prior_alpha_sd);
      }
      if (i >= mc.off_beta && i < mc.off_beta + K) g += -theta[i] * ib2;
      if (has_scale && i == P - 1) {
        double dlik = 0.0;
        if (mc.N_total > 0)
          dlik = family == FAM_NORMAL_ID ? (lik[P] - mc.N_total) / sigma   // normal_id_glm_lpdf.hpp:181-183
                                            : lik[P + 1];                     // neg_binomial_2_log_glm_lpmf.hpp:240-245
        const double dpri = -(sigma - mc.prior_sigma_loc) / (mc.prior_sigma_scale * mc.prior_sigma_scale);
        g = (dlik + dpri) * sigma + (mc.jacobian ? 1.0 : 0.0);
      }
    }
    p.result[1 + i] = g;
    if (hres) hres[1 + i] = g;
    if (lf) {
      const double g0 = src.g0 ? src.g0[i] : p.st_in[2 * P + i];
      const double ph = src.ph ? src.ph[i] : p.st_in[P + i] - he * g0;
      const double gnew = domain ? -g0 : -g;
      const double pnew = ph - he * gnew;
      p.st_out[i] = theta[i];
      p.st_out[2 * P + i] = gnew;
      p.st_out[P + i] = pnew;
      if (hst) {
        hst[i] = theta[i];
        hst[2 * P + i] = gnew;
        hst[P + i] = pnew;
      }
    }
  }
  finish_stamp(p, 4);                             // gradient + leapfrog tail written
  if (tid == 0) {
    const double status = domain ? (double)ST_DOMAIN : (double)ST_OK;
    p.result[0] = lp;
    p.result[1 + P] = status;
    if (lf) p.st_out[3 * P] = domain ? CUDART_INF : -lp;
    if (hres) {
      hres[0] = lp;
      hres[1 + P] = status;
      if (lf) hst[3 * P] = domain ? CUDART_INF : -lp;
    }
  }
  host_out_publish(p);
}

__device__ inline void finish(const KernelParams& p, double* sh, const FinishSrc& src) { finish_t<-1>(p, sh, src); }
// the epilogue from the global arrays (finish_kernel, host build)
__device__ inline void finish(const KernelParams& p, double* sh) {
  const FinishSrc src = {p.theta_used, p.lik, nullptr, nullptr};
  finish_t<-1>(p, sh, src);
}

// ------------------------------------------------------------------------------------------
// Model epilogues of the two class models (ordered_logistic_glm, categorical_logit_glm).  No kernel calls them
// yet: like their row arithmetic in glm_link.cuh they are written ahead of the kernels
// and pinned to the oracle through the host build (tests/test_model_epilogue_host.py).  The Stan programs:
//   ordered_logistic:   parameters { vector[K] beta; ordered[C-1] c; }            theta = [beta, c unconstrained]
//                       beta ~ normal(0, prior_beta_sd); c ~ normal(0, prior_alpha_sd);
//   categorical_logit:  parameters { vector[C] alpha; matrix[K, C] beta; }        theta = [alpha, beta col-major]
//                       alpha ~ normal(0, prior_alpha_sd); to_vector(beta) ~ normal(0, prior_beta_sd);
// ------------------------------------------------------------------------------------------
struct ClassModelParams {
  int family_ordered;        // 1 = ordered_logistic, 0 = categorical_logit
  int K, C, P;               // attributes, classes, parameters
  int propto, jacobian, is_var;
  int mode;                  // MODE_THETA / MODE_LEAPFROG
  double N_total;            // rows over all shards
  double prior_alpha_sd, prior_beta_sd;
  double eps;
  const double* theta_used;  // P
  const double* lik;         // P + 1: d logp / d (constrained parameter) sums aligned with theta, then the lp-sum.
                             // ordered_logistic: entries [K, K + C - 1) are the partials wrt the cut-points c
  double* cuts;              // ordered_logistic: 2 (C - 1) doubles of scratch (cut-points, their partials)
  double* result;            // [lp, grad(P), status]
  const double* st_in;       // MODE_LEAPFROG: [q(P) p(P) g(P) V]
  double* st_out;
};

__device__ void finish_class_model(const ClassModelParams& p, double* sh /* >= 8 doubles scratch */) {
  const int K = p.K, C = p.C, P = p.P;
  const double* theta = p.theta_used;
  const double* lik = p.lik;
  const int tid = threadIdx.x, nt = blockDim.x;
  const bool dens = (!p.propto) || p.is_var;   // include_summand: anything left to compute?
  const double ia2 = 1.0 / (p.prior_alpha_sd * p.prior_alpha_sd), ib2 = 1.0 / (p.prior_beta_sd * p.prior_beta_sd);
  const int nc = p.family_ordered ? (C > 0 ? C - 1 : 0) : 0;
  // the likelihood term is empty without rows, with one class (categorical...:70-72) or without cut-points (ordered...:93-95)
  const bool lik_on = dens && p.N_total > 0 && C > 1;

  if (tid == 0) {
    double lp = 0.0, bad = 0.0;
    for (int i = 0; i < P; ++i)
      if (!isfinite(theta[i]) || !isfinite(lik[i])) bad += 1.0;
    double sb = 0.0, sa = 0.0;
    if (p.family_ordered) {
      // ordered_constrain.hpp:34-37 (+ Jacobian :56-58), then check_ordered / check_finite (ordered...:85-91)
      for (int k = 0; k < nc; ++k) {
        p.cuts[k] = k == 0 ? theta[K] : p.cuts[k - 1] + exp(theta[K + k]);
        if (p.jacobian && k > 0) lp += theta[K + k];
        if (k > 0 && !(p.cuts[k] > p.cuts[k - 1])) bad += 1.0;
        sa += p.cuts[k] * p.cuts[k];
      }
      if (nc > 0 && (!isfinite(p.cuts[0]) || !isfinite(p.cuts[nc - 1]))) bad += 1.0;
      for (int k = 0; k < K; ++k) sb += theta[k] * theta[k];
    } else {
      for (int c = 0; c < C; ++c) sa += theta[c] * theta[c];
      for (int i = C; i < P; ++i) sb += theta[i] * theta[i];
    }
    if (dens) {   // normal_lpdf.hpp:81-88 on vectors, the scale is data
      const int n_a = p.family_ordered ? nc : C, n_b = p.family_ordered ? K : K * C;
      if (n_a > 0) {
        lp += -0.5 * sa * ia2;
        if (!p.propto) lp += n_a * (NEG_LOG_SQRT_TWO_PI_D - log(p.prior_alpha_sd));
      }
      if (n_b > 0) {
        lp += -0.5 * sb * ib2;
        if (!p.propto) lp += n_b * (NEG_LOG_SQRT_TWO_PI_D - log(p.prior_beta_sd));
      }
    }
    if (lik_on) lp += lik[P];
    sh[0] = lp;
    sh[1] = (isfinite(lp) && bad == 0.0) ? 0.0 : 1.0;
    // ordered_logistic: partials wrt the cut-points (likelihood + prior), then the chain rule through
    // ordered_constrain: d c_j / d u_k = exp(u_k) for j >= k (1 for k = 0), + the Jacobian term
    if (p.family_ordered) {
      double tail = 0.0;
      for (int k = nc - 1; k >= 0; --k) {
        tail += (lik_on ? lik[K + k] : 0.0) - (dens ? p.cuts[k] * ia2 : 0.0);
        p.cuts[nc + k] = k == 0 ? tail : tail * exp(theta[K + k]) + (p.jacobian ? 1.0 : 0.0);
      }
    }
  }
  __syncthreads();
  const double lp = sh[0];
  const bool domain = sh[1] != 0.0;
  for (int i = tid; i < P; i += nt) {
    double g;
    if (p.family_ordered) {
      g = i < K ? (lik_on ? lik[i] : 0.0) - (dens ? theta[i] * ib2 : 0.0) : p.cuts[nc + (i - K)];
    } else {
      g = (lik_on ? lik[i] : 0.0) - (dens ? theta[i] * (i < C ? ia2 : ib2) : 0.0);
    }
    p.result[1 + i] = g;
  }
  __syncthreads();
  if (tid == 0) {
    p.result[0] = lp;
    p.result[1 + P] = domain ? (double)ST_DOMAIN : (double)ST_OK;
  }
  if (p.mode == MODE_LEAPFROG) {   // expl_leapfrog.hpp:28-32 end_update_p; base_hamiltonian.hpp:64-69
    const double* p0 = p.st_in + P;
    const double* g0 = p.st_in + 2 * P;
    const double he = 0.5 * p.eps;
    for (int i = tid; i < P; i += nt) {
      const double ph = p0[i] - he * g0[i];
      const double gnew = domain ? -g0[i] : -p.result[1 + i];
      p.st_out[i] = theta[i];
      p.st_out[2 * P + i] = gnew;
      p.st_out[P + i] = ph - he * gnew;
    }
    if (tid == 0) p.st_out[3 * P] = domain ? CUDART_INF : -lp;
  }
}

// What the class-model kernels run once the likelihood sums are complete in p.lik: the epilogue above on the launch's
// KernelParams, then the mirror of result / state into pinned host memory for host-facing calls.
__device__ inline void class_epilogue(const KernelParams& p, double* sh) {
  ClassModelParams cp;
  cp.family_ordered = p.family == FAM_ORDERED_LOGISTIC ? 1 : 0;
  cp.K = p.K;
  cp.C = p.n_classes;
  cp.P = p.P;
  cp.propto = p.mc.propto;
  cp.jacobian = p.mc.jacobian;
  cp.is_var = p.mc.is_var;
  cp.mode = p.mode;
  cp.N_total = p.mc.N_total;
  cp.prior_alpha_sd = p.mc.prior_alpha_sd;
  cp.prior_beta_sd = p.mc.prior_beta_sd;
  cp.eps = p.eps;
  cp.theta_used = p.theta_used;
  cp.lik = p.lik;
  cp.cuts = p.cuts;
  cp.result = p.result;
  cp.st_in = p.st_in;
  cp.st_out = p.st_out;
  finish_class_model(cp, sh);
  __syncthreads();
  if (p.host_out) {
    const int P = p.P;
    for (int i = threadIdx.x; i < P + 2; i += blockDim.x) p.host_out[i] = p.result[i];
    if (p.mode == MODE_LEAPFROG)
      for (int i = threadIdx.x; i < 3 * P + 1; i += blockDim.x) p.host_out[(P + 2) + i] = p.st_out[i];
  }
  host_out_publish(p);
}

}  // namespace b200glm
