// tests/host/link_host.cpp -- TEST INFRASTRUCTURE: the kernels' per-row arithmetic (stan_b200/csrc/glm_link.cuh,
// the very source the GPU executes) compiled for the HOST, so tests/test_link_math_host.py can check it row by
// row against the CPU oracle without a GPU.  Built by the test with g++ into a temporary shared object.
#include <math.h>
#include <stddef.h>
#include <string.h>

// one-thread CTA: what the CUDA built-ins finish() uses mean when a single host thread runs the whole block
#define __device__
#define __forceinline__ inline
#define CUDART_INF INFINITY
#define CUDART_NAN NAN
namespace b200glm {
struct host_dim3 {
  int x;
};
static const host_dim3 threadIdx = {0}, blockDim = {1};
inline void __syncthreads() {}
inline double warp_sum(double v) { return v; }   // a warp of one lane
}  // namespace b200glm

#include "../../stan_b200/csrc/glm_link.cuh"
#include "../../stan_b200/csrc/glm_model.cuh"

using namespace b200glm;

template <int FAMILY>
static void rows(int n, const double* eta, const double* y, const double* aux, const LinkConst& lc, double* lp,
                 double* r, double* x) {
  for (int i = 0; i < n; ++i) link_ext<FAMILY>(eta[i], y[i], aux ? aux[i] : 0.0, lc, lp[i], r[i], x[i]);
}

extern "C" {

// what the kernel prologue derives from the scale parameter (glm_kernels.cuh) + one link_ext call per row
int link_rows(int family, int n, const double* eta, const double* y, const double* aux, double scale,
              int inc_phi_terms, int inc_ytheta, double* lp, double* r, double* x) {
  LinkConst lc;
  lc.inv_sigma = 1.0;
  lc.phi = 1.0;
  lc.log_phi = lc.dg_phi = lc.lg_phi = 0.0;
  lc.inc_phi_terms = inc_phi_terms;
  lc.inc_ytheta = inc_ytheta;
  if (family == FAM_NORMAL_ID) lc.inv_sigma = 1.0 / scale;
  if (family == FAM_NEG_BINOMIAL_2_LOG) {
    lc.phi = scale;
    lc.log_phi = log(scale);
    lc.dg_phi = digamma_pos(scale);
    lc.lg_phi = lgamma(scale);
  }
  switch (family) {
    case FAM_BERNOULLI_LOGIT: rows<FAM_BERNOULLI_LOGIT>(n, eta, y, aux, lc, lp, r, x); return 0;
    case FAM_POISSON_LOG: rows<FAM_POISSON_LOG>(n, eta, y, aux, lc, lp, r, x); return 0;
    case FAM_NORMAL_ID: rows<FAM_NORMAL_ID>(n, eta, y, aux, lc, lp, r, x); return 0;
    case FAM_BINOMIAL_LOGIT: rows<FAM_BINOMIAL_LOGIT>(n, eta, y, aux, lc, lp, r, x); return 0;
    case FAM_NEG_BINOMIAL_2_LOG: rows<FAM_NEG_BINOMIAL_2_LOG>(n, eta, y, aux, lc, lp, r, x); return 0;
  }
  return 1;
}

double digamma_host(double v) { return digamma_pos(v); }

// link<> and its branch-free form link_bf<> (glm_multi_kernel.cuh uses the latter) on the same rows, and the special
// functions of link_bf on their own
// (binomial_logit: aux = population sizes, link_ext<> is the reference form)
int link_pair_rows(int family, int n, const double* eta, const double* y, double inv_sigma, double* lp, double* r,
                   double* lp_bf, double* r_bf, const double* aux) {
  LinkConst lc;
  memset(&lc, 0, sizeof(lc));
  for (int i = 0; i < n; ++i) {
    switch (family) {
      case FAM_BINOMIAL_LOGIT: {
        double x;
        link_ext<FAM_BINOMIAL_LOGIT>(eta[i], y[i], aux[i], lc, lp[i], r[i], x);
        link_bf<FAM_BINOMIAL_LOGIT>(eta[i], y[i], inv_sigma, lp_bf[i], r_bf[i], aux[i]);
        break;
      }
      case FAM_BERNOULLI_LOGIT:
        link<FAM_BERNOULLI_LOGIT>(eta[i], y[i], inv_sigma, lp[i], r[i]);
        link_bf<FAM_BERNOULLI_LOGIT>(eta[i], y[i], inv_sigma, lp_bf[i], r_bf[i]);
        break;
      case FAM_POISSON_LOG:
        link<FAM_POISSON_LOG>(eta[i], y[i], inv_sigma, lp[i], r[i]);
        link_bf<FAM_POISSON_LOG>(eta[i], y[i], inv_sigma, lp_bf[i], r_bf[i]);
        break;
      case FAM_NORMAL_ID:
        link<FAM_NORMAL_ID>(eta[i], y[i], inv_sigma, lp[i], r[i]);
        link_bf<FAM_NORMAL_ID>(eta[i], y[i], inv_sigma, lp_bf[i], r_bf[i]);
        break;
      default: return 1;
    }
  }
  return 0;
}
void fm_exp_rows(int n, const double* x, double* out) {
  for (int i = 0; i < n; ++i) out[i] = fm_exp(x[i]);
}
void fm_log1p_rows(int n, const double* u, double* out, double* rw) {
  for (int i = 0; i < n; ++i) out[i] = fm_log1p(u[i], rw[i]);
}

// The model epilogue finish() (glm_model.cuh) as one host thread.
//   ic = {family, K, G, P, off_beta, propto, jacobian, is_var, lik_only, sigma_is_var, mode}
//   dc = {N_total, lgamma_sum, prior_alpha_sd, prior_beta_sd, prior_sigma_loc, prior_sigma_scale, prior_sigma_a_scale, eps}
//   lik: P + 2 likelihood sums laid out as the kernels leave them; result: P + 2; st_in / st_out: 3P + 1 (leapfrog mode)
void finish_host(const int* ic, const double* dc, const double* theta_used, double* lik, double* result,
                 const double* st_in, double* st_out) {
  KernelParams p;
  memset(&p, 0, sizeof(p));
  ModelConst& mc = p.mc;
  mc.family = ic[0];
  mc.K = ic[1];
  mc.G = ic[2];
  mc.P = ic[3];
  mc.off_beta = ic[4];
  mc.propto = ic[5];
  mc.jacobian = ic[6];
  mc.is_var = ic[7];
  mc.lik_only = ic[8];
  mc.sigma_is_var = ic[9];
  p.mode = ic[10];
  mc.N_total = dc[0];
  mc.lgamma_sum = dc[1];
  mc.prior_alpha_sd = dc[2];
  mc.prior_beta_sd = dc[3];
  mc.prior_sigma_loc = dc[4];
  mc.prior_sigma_scale = dc[5];
  mc.prior_sigma_a_scale = dc[6];
  p.eps = dc[7];
  p.K = mc.K;
  p.G = mc.G;
  p.P = mc.P;
  p.family = mc.family;
  p.off_beta = mc.off_beta;
  p.theta_used = const_cast<double*>(theta_used);
  p.lik = lik;
  p.result = result;
  p.st_in = st_in;
  p.st_out = st_out;
  double scratch[64];
  finish(p, scratch);
}

// ordered_logistic: out[4 * i + {0,1,2,3}] = lp_i, w_i, d1, d2
void ordered_logistic_rows(int n, const double* loc, const int* c, int C, const double* cuts, double* out) {
  for (int i = 0; i < n; ++i)
    ordered_logistic_row(loc[i], c[i], C, cuts, out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
}

// categorical_logit: lin is n x C row-major, overwritten with the per-class weights; lp[i] = the row's log-density
void categorical_logit_rows(int n, int C, const int* y, double* lin, double* lp) {
  for (int i = 0; i < n; ++i) lp[i] = categorical_logit_row(C, y[i], lin + (size_t)i * C);
}

// finish_class_model() as one host thread.
//   ic = {family_ordered, K, C, P, propto, jacobian, is_var, mode};  dc = {N_total, prior_alpha_sd, prior_beta_sd, eps}
void finish_class_model_host(const int* ic, const double* dc, const double* theta_used, const double* lik, double* cuts,
                             double* result, const double* st_in, double* st_out) {
  ClassModelParams p;
  memset(&p, 0, sizeof(p));
  p.family_ordered = ic[0];
  p.K = ic[1];
  p.C = ic[2];
  p.P = ic[3];
  p.propto = ic[4];
  p.jacobian = ic[5];
  p.is_var = ic[6];
  p.mode = ic[7];
  p.N_total = dc[0];
  p.prior_alpha_sd = dc[1];
  p.prior_beta_sd = dc[2];
  p.eps = dc[3];
  p.theta_used = theta_used;
  p.lik = lik;
  p.cuts = cuts;
  p.result = result;
  p.st_in = st_in;
  p.st_out = st_out;
  double scratch[8];
  finish_class_model(p, scratch);
}

}  // extern "C"
