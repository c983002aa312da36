/*
 * oracle/glm_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Data/model description shared by the two CPU checkers under oracle/:
 *   - oracle/glm_oracle.c      : plain-C restatement ("port") of the reference arithmetic
 *   - oracle/ref/ref_oracle.cpp: the reference itself (Stan + Stan Math headers compiled
 *                                from /root/reference), exposed through the same C calls
 *
 * Nothing in the product (stan_b200/, include/) may include, link or call anything in
 * this directory; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do, and only as the checker or the timed CPU baseline.
 *
 * The model the spec describes is the hand-written Stan program (no stanc offline):
 *
 *   data { int N; int K; matrix[N,K] X; y; [int G; array[N] int group;] }
 *   parameters {                       // unconstrained layout, in this order
 *     G == 0:  real alpha;                                  theta[0]
 *     G  > 0:  real mu_a; real<lower=0> sigma_a; vector[G] a;   theta[0], theta[1]=log sigma_a, theta[2..2+G)
 *     vector[K] beta;
 *     family == NORMAL_ID: real<lower=0> sigma;             last entry = log sigma
 *     family == NEG_BINOMIAL_2_LOG: real<lower=0> phi;      last entry = log phi
 *   }
 *   model {
 *     G == 0:  alpha ~ normal(0, prior_alpha_sd);
 *     G  > 0:  mu_a ~ normal(0, prior_alpha_sd); sigma_a ~ normal(0, prior_sigma_a_scale);
 *              a ~ normal(mu_a, sigma_a);
 *     beta ~ normal(0, prior_beta_sd);
 *     NORMAL_ID: sigma ~ normal(prior_sigma_loc, prior_sigma_scale);
 *     NEG_BINOMIAL_2_LOG: phi ~ normal(prior_sigma_loc, prior_sigma_scale);
 *     y ~ <family>_glm(X, G == 0 ? alpha : a[group], beta [, sigma]);
 *   }
 */
#ifndef ORACLE_GLM_ORACLE_H
#define ORACLE_GLM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { GLM_BERNOULLI_LOGIT = 0, GLM_POISSON_LOG = 1, GLM_NORMAL_ID = 2,
       GLM_BINOMIAL_LOGIT = 3,      /* y ~ binomial_logit_glm(trials, X, alpha, beta)          */
       GLM_NEG_BINOMIAL_2_LOG = 4,  /* y ~ neg_binomial_2_log_glm(X, alpha, beta, phi); phi is the
                                       last parameter (log phi unconstrained), phi ~ normal(prior_sigma_loc,
                                       prior_sigma_scale) -- the slot sigma has for NORMAL_ID */
       /* The two GLMs below have their own parameter blocks (G must be 0); the oracle for them is ahead of the
        * CUDA path (not built on the device yet). */
       GLM_ORDERED_LOGISTIC = 5,    /* parameters { vector[K] beta; ordered[C-1] c; }   theta = [beta, c unconstrained]
                                       beta ~ normal(0, prior_beta_sd); c ~ normal(0, prior_alpha_sd);
                                       y ~ ordered_logistic_glm(X, beta, c);           y in 1..C */
       GLM_CATEGORICAL_LOGIT = 6    /* parameters { vector[C] alpha; matrix[K, C] beta; }  theta = [alpha, beta col-major]
                                       alpha ~ normal(0, prior_alpha_sd); to_vector(beta) ~ normal(0, prior_beta_sd);
                                       y ~ categorical_logit_glm(X, alpha, beta);      y in 1..C */ };

typedef struct glm_spec {
  int32_t family;
  int32_t K;
  int64_t N;
  const double* X; /* column-major N x K, leading dimension ldx (>= N) */
  int64_t ldx;
  const int32_t* y_int; /* bernoulli (0/1), poisson (>=0); NULL for normal */
  const double* y_real; /* normal; NULL otherwise */
  int32_t G;            /* 0: scalar intercept; >0: hierarchical a[group] */
  int32_t _pad;
  const int32_t* group; /* N entries, 1-based (Stan indexing); NULL when G == 0 */
  double prior_alpha_sd;
  double prior_beta_sd;
  double prior_sigma_loc;
  double prior_sigma_scale;
  double prior_sigma_a_scale;
  const int32_t* trials; /* BINOMIAL_LOGIT: population sizes (N entries); NULL otherwise */
  int32_t n_classes;     /* ORDERED_LOGISTIC / CATEGORICAL_LOGIT: number of outcome classes C (>= 1) */
  int32_t _pad2;
} glm_spec;

/* number of unconstrained parameters of the model the spec describes */
int32_t glm_oracle_num_params(const glm_spec* s);

/*
 * C port.  log_prob (+gradient when grad != NULL) at unconstrained theta.
 *  propto=1 & grad!=NULL : what stan::model::log_prob_grad<true,jacobian> returns
 *  propto=0              : what model.log_prob<false,jacobian>(double) returns (all constants)
 * returns 0 OK, 1 domain error (message in errbuf), 2 invalid argument.
 */
int glm_oracle_log_prob_grad(const glm_spec* s, const double* theta, int propto,
                             int jacobian, double* lp, double* grad,
                             char* errbuf, int errlen);

/* double semantics: Model::log_prob<propto,jacobian>(double); propto=1 drops every density */
int glm_oracle_log_prob(const glm_spec* s, const double* theta, int propto,
                        int jacobian, double* lp, char* errbuf, int errlen);

/* one explicit leapfrog step on (q,p,g,V) with diagonal inverse metric, in place
 * (expl_leapfrog.hpp:16-32 + base_hamiltonian.hpp:61-70) */
int glm_oracle_leapfrog(const glm_spec* s, double eps, const double* inv_metric,
                        double* q, double* p, double* g, double* V,
                        char* errbuf, int errlen);

#ifdef __cplusplus
}
#endif
#endif
