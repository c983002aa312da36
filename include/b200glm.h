/*
 * b200glm.h -- C ABI of the B200-native GLM log-density + gradient backend.
 *
 * This is the drop-in boundary.  Every entry point
 * names the reference interface it replaces; paths are relative to the reference tree
 * (ST = src/stan, SM = lib/stan_math/stan/math).  Plain pointers and sizes only: no C++ types,
 * no exceptions and no torch types cross this line.  The reference-side binding (a
 * stan::model::model_base_crtp model + explicit specialisations of stan::model::gradient and
 * stan::mcmc::expl_leapfrog) is stan_b200/cpp/b200/stan_glm_model.hpp, see INTEGRATION.md.
 *
 * There is NO CPU fallback: every compute entry point returns B200GLM_CUDA if no sm_100
 * device is usable.
 *
 * Model (the hand-written Stan program):
 *   G == 0: theta = [alpha, beta_1..K (, log sigma)]
 *   G  > 0: theta = [mu_a, log sigma_a, a_1..G, beta_1..K (, log sigma)]
 *   (neg_binomial_2_log: the trailing entry is log phi, the precision, with the prior sigma has for normal_id)
 *   priors alpha|mu_a ~ N(0, prior_alpha_sd), sigma_a ~ N(0, prior_sigma_a_scale),
 *   a ~ N(mu_a, sigma_a), beta ~ N(0, prior_beta_sd), sigma ~ N(prior_sigma_loc, prior_sigma_scale)
 *   likelihood y ~ {bernoulli_logit,poisson_log,normal_id}_glm(X, alpha | a[group], beta [, sigma]),
 *              y ~ binomial_logit_glm(trials, X, alpha | a[group], beta),
 *              y ~ neg_binomial_2_log_glm(X, alpha | a[group], beta, phi).
 *
 * Environment switches (read at b200glm_create / b200glm_batch_reserve; for A/B measurements only):
 *   B200GLM_NO_PDL=1       launch without the programmatic-dependent-launch attribute
 *   B200GLM_NO_ROWSPLIT=1  serve batches of <= 16 lanes with the normal batched kernel, not its row-split variant
 *   B200GLM_NO_MULTI=1     serve batches of <= 8 lanes (K <= 128) with the DMMA kernels, not the few-chain FMA kernel
 *   B200GLM_WIDE_PRODUCER=single|lanes   wide-matrix kernel: one lane issues every bulk copy | lane j streams sub-panel j
 *   B200GLM_NO_STATE_SMEM=1  keep the chain state of the fused epilogue in global memory
 *   B200GLM_NO_GROUP_FUSION=1, B200GLM_PDL_PREFETCH=<stages>, B200GLM_WIDE_ROWS=16|8|4, B200GLM_TL_REPEAT=1  
 */
#ifndef B200GLM_H
#define B200GLM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200GLM_ABI_VERSION 5   /* 2: b200glm_desc gained `trials` (appended), families 3 and 4
                                   3: b200glm_abi_version, shard constants, timeline, (no struct change)
                                   4: b200glm_desc gained `n_classes` (appended), families 5 and 6, streamed build
                                   5: b200glm_nuts_config gained `stepsize_jitter` (appended); a chain's row of the
                                      `uniforms` buffer is 72 doubles (the 64-entry ring + the jitter variate) */

/* status codes; the C++ shim maps them to the exceptions the reference throws:
 * DOMAIN -> std::domain_error (recoverable: base_hamiltonian.hpp:65-68, initialize.hpp:104-112),
 * INVALID -> std::invalid_argument (check_consistent_size), CUDA -> std::runtime_error (fatal). */
enum { B200GLM_OK = 0, B200GLM_DOMAIN = 1, B200GLM_INVALID = 2, B200GLM_CUDA = 3 };

/* SM/prim/prob/{bernoulli_logit_glm_lpmf.hpp:49, poisson_log_glm_lpmf.hpp:51, normal_id_glm_lpdf.hpp:54,
 * binomial_logit_glm_lpmf.hpp:55, neg_binomial_2_log_glm_lpmf.hpp:64}.  The last two run in
 * the single-chain kernels (narrow and wide-matrix); the batched kernels (DMMA and few-chain) serve families 0-3. */
enum { B200GLM_BERNOULLI_LOGIT = 0, B200GLM_POISSON_LOG = 1, B200GLM_NORMAL_ID = 2, B200GLM_BINOMIAL_LOGIT = 3,
       B200GLM_NEG_BINOMIAL_2_LOG = 4,
       /* SM/prim/prob/ordered_logistic_glm_lpmf.hpp:49 and categorical_logit_glm_lpmf.hpp:47 -- class-outcome models
        * with their own parameter blocks (y in 1..n_classes, no intercept vector a[group]):
        *   ordered_logistic:  parameters { vector[K] beta; ordered[C-1] c; }   theta = [beta, c unconstrained]
        *                      beta ~ N(0, prior_beta_sd); c ~ N(0, prior_alpha_sd); y ~ ordered_logistic_glm(X, beta, c)
        *   categorical_logit: parameters { vector[C] alpha; matrix[K, C] beta; }   theta = [alpha, beta column-major]
        *                      alpha ~ N(0, prior_alpha_sd); to_vector(beta) ~ N(0, prior_beta_sd);
        *                      y ~ categorical_logit_glm(X, alpha, beta)
        * Served by glm_class_kernel (single chain, K <= 256 | K x classes limits in b200glm_create's error text);
        * not by the batched, wide-matrix or function-level entry points. */
       B200GLM_ORDERED_LOGISTIC = 5, B200GLM_CATEGORICAL_LOGIT = 6 };

/* desc.flags: use the wide-matrix kernel (16-row panels split over the CTA; the default for K > 256)
 * even for a narrow X -- for tests of that kernel at small K */
#define B200GLM_FLAG_FORCE_WIDE 1
/* desc.flags: streamed construction -- b200glm_create only reserves the panels for desc.N rows (X, y, trials are
 * ignored and may be NULL; G must be 0); the rows then arrive through b200glm_append_rows in chunks, each re-laid out
 * into the panels at once, and b200glm_finalize makes the handle usable.  For matrices that fill the HBM: X is never
 * resident twice (BASELINE configs[4]: 8 x ~160 GB). */
#define B200GLM_FLAG_STREAMED 2

typedef struct b200glm_handle b200glm_handle;

typedef struct b200glm_desc {
  int32_t family;
  int32_t K;            /* columns of X ("attributes") */
  int64_t N;            /* rows held by THIS handle (the local shard when world > 1) */
  const double* X;      /* column-major N x K (Eigen::MatrixXd layout), leading dimension ldx */
  int64_t ldx;
  const int32_t* y_int; /* bernoulli / poisson / binomial (successes) / neg_binomial_2 */
  const double* y_real; /* normal */
  int32_t G;            /* 0 = scalar intercept, >0 = a[group] (ST/model/indexing/rvalue.hpp:154-172) */
  int32_t data_on_device; /* 0: X,y,group are host pointers (copied); 1: device pointers on `device` */
  const int32_t* group; /* 1-based */
  double prior_alpha_sd, prior_beta_sd, prior_sigma_loc, prior_sigma_scale, prior_sigma_a_scale;
  int32_t device;       /* CUDA device ordinal */
  int32_t n_slots;      /* independent chain slots (each has its own stream + workspace); >= 1 */
  int32_t rank, world;  /* row shard `rank` of `world` (world <= 1: unsharded).  With world > 1 the
                           likelihood partials are summed across ranks (b200glm_comm_*) and the
                           priors are added once, identically, on every rank. */
  int64_t N_total;      /* rows over all shards (normal_id needs N for -N log sigma); 0 => N */
  int32_t grid_ctas;    /* 0 = one persistent CTA per SM */
  int32_t flags;        /* B200GLM_FLAG_* */
  const int32_t* trials; /* binomial_logit: population sizes (N entries; host or device like y_int); else NULL */
  int32_t n_classes;    /* ordered_logistic / categorical_logit: number of outcome classes C (1..16); else 0 */
} b200glm_desc;

/* Data upload + one-time re-layout of X into the row-panel format the kernel streams
 * (replaces the Model constructor copying data out of a var_context, and SM/opencl/copy.hpp:45
 * to_matrix_cl).  The caller keeps ownership of every pointer in desc. */
int b200glm_create(const b200glm_desc* desc, b200glm_handle** out);
void b200glm_destroy(b200glm_handle* h);
/* Streamed construction (B200GLM_FLAG_STREAMED): the next n rows of the design matrix -- X column-major n x K with
 * leading dimension ldx, y_int or y_real, trials (binomial_logit) -- host or device pointers as desc.data_on_device
 * says.  Every chunk but the last must be a multiple of 32 rows (the panel height).  The buffers may be reused as soon
 * as the call returns.  (The stanc-generated constructor reads the whole data block at once; this is the analogue for
 * a matrix that only fits once: SM/opencl/copy.hpp:45 to_matrix_cl, chunk by chunk.) */
int b200glm_append_rows(b200glm_handle* h, int64_t n, const double* X, int64_t ldx, const int32_t* y_int,
                        const double* y_real, const int32_t* trials);
int b200glm_finalize(b200glm_handle* h);

/* prob_grad::num_params_r()  (ST/model/prob_grad.hpp:19-84) */
int32_t b200glm_num_params(const b200glm_handle* h);

/* Value and gradient with autodiff-variable semantics:
 *   propto=1, jacobian=1 == stan::model::log_prob_grad<true,true>(model, theta, ., grad)
 *   (ST/model/log_prob_grad.hpp:29-50) == stan::model::gradient (ST/model/gradient.hpp:22-35).
 * theta/lp/grad are HOST pointers; grad may be NULL.  Synchronous on the slot's stream. */
int b200glm_log_prob_grad(b200glm_handle* h, int32_t slot, const double* theta, int32_t propto,
                          int32_t jacobian, double* lp, double* grad);

/* Value with plain-double semantics: Model::log_prob<propto,jacobian>(vector<double>&, ...)
 * as called from ST/services/util/initialize.hpp:128 (<false,jacobian>: all constants kept;
 * propto=1 drops every density term, exactly as include_summand does for doubles). */
int b200glm_log_prob(b200glm_handle* h, int32_t slot, const double* theta, int32_t propto,
                     int32_t jacobian, double* lp);

/* Function-level entry: the GLM term ALONE (no priors, no Jacobian), the slot where the reference's OpenCL
 * backend plugs in -- an overload of the density selected by argument type that returns
 * ops_partials.build(logp) (SM/opencl/prim/bernoulli_logit_glm_lpmf.hpp:52-58, :105-138).  Value and
 * partials of {bernoulli_logit,poisson_log,normal_id}_glm_lp*f<propto>(y, X, alpha, beta [, sigma]),
 * binomial_logit_glm_lpmf<propto>(y, trials, X, alpha, beta) or neg_binomial_2_log_glm_lpmf<propto>(y, X, alpha,
 * beta, phi) (phi passed as `sigma`, its partial returned in d_sigma, sigma_is_var = 1 "phi is a var", 2 "phi is
 * the ONLY var operand": y * theta then drops under propto, neg_binomial_2_log_glm_lpmf.hpp:188-190) for the
 * (y, X [, trials] [, group]) the handle holds; alpha: 1 value (G == 0) or G values (alpha = a[group]).
 * operands_are_var = 0 with propto = 1 returns 0 as the reference does (all-constant, include_summand);
 * sigma_is_var only matters for normal_id under propto (-N log sigma kept iff sigma is an autodiff
 * variable, normal_id_glm_lpdf.hpp:205-212).  d_alpha / d_beta / d_sigma may be NULL.
 * The C++ binding (stan::math overloads on b200::glm_data views) is stan_b200/cpp/b200/glm_functions.hpp. */
int b200glm_glm_lpmf(b200glm_handle* h, int32_t slot, int32_t propto, int32_t operands_are_var,
                     int32_t sigma_is_var, const double* alpha, const double* beta, double sigma,
                     double* logp, double* d_alpha, double* d_beta, double* d_sigma);

/* The same entry with PER-ROW operands, the other forms the reference's overloads accept
 * (SM/opencl/prim/normal_id_glm_lpdf.hpp:68-84: is_alpha_vector, is_sigma_vector; the prim versions likewise):
 * alpha_rows != NULL: the intercept is an N-vector (alpha is then ignored), its partials -- the per-row residual --
 * come back in d_alpha_rows (N); sigma_rows != NULL (normal_id only): the scale is an N-vector (sigma ignored), value
 * -1/2 sum z_i^2 - sum log sigma_i [+ constants], partials (z_i^2 - 1) / sigma_i in d_sigma_rows (N).  All pointers
 * are HOST pointers; NULL outputs are skipped.  Costs an N-vector upload per vector operand and an N-vector download
 * per vector of partials on top of the single pass over X.  Scalar-intercept handles (G == 0), K <= 256, unsharded. */
int b200glm_glm_lpmf_rows(b200glm_handle* h, int32_t slot, int32_t propto, int32_t operands_are_var,
                          int32_t sigma_is_var, const double* alpha_rows, double alpha, const double* beta,
                          const double* sigma_rows, double sigma, double* logp, double* d_alpha_rows, double* d_alpha,
                          double* d_beta, double* d_sigma_rows, double* d_sigma);

/* Device-resident leapfrog (replaces expl_leapfrog::evolve, ST/mcmc/hmc/integrators/
 * base_leapfrog.hpp:17-22 + expl_leapfrog.hpp:16-32, and the update_potential_gradient it calls,
 * base_hamiltonian.hpp:61-70).  set_state uploads z = (q, p, g, V) for a slot;
 * leapfrog advances it by eps in ONE launch (half p, full q, gradient, half p) and mirrors the
 * new (q, p, g, V) to the host pointers (any may be NULL).  inv_metric NULL = keep the last one
 * (initially all ones).  On a domain error V=+inf and g is negated, as the reference does. */
int b200glm_set_state(b200glm_handle* h, int32_t slot, const double* q, const double* p,
                      const double* g, double V);
int b200glm_leapfrog(b200glm_handle* h, int32_t slot, double eps, const double* inv_metric,
                     double* q, double* p, double* g, double* V);

/* Asynchronous forms for callers that keep theta on the device (bench `value`, batched driver):
 * enqueue on the slot's stream, no host copies; b200glm_stream returns that cudaStream_t. */
int b200glm_leapfrog_async(b200glm_handle* h, int32_t slot, double eps);
int b200glm_grad_async(b200glm_handle* h, int32_t slot, const double* theta_device);
int b200glm_sync(b200glm_handle* h, int32_t slot);
void* b200glm_stream(b200glm_handle* h, int32_t slot);
/* device pointer to the slot's result block [lp, grad[P], status] (doubles) */
const double* b200glm_result_device(b200glm_handle* h, int32_t slot);

/* Batched chains: n chains evaluated in ONE pass over X; the per-chain GEMV pair becomes a pair of
 * fp64 GEMMs on the DMMA tensor path (BASELINE configs[2]).  This is the device side of a multi-chain
 * driver in ST/services/sample (hmc_nuts_diag_e_adapt.hpp:364-401 runs chains as independent TBB
 * tasks, each calling stan::model::gradient on its own; here the calls of all chains that are waiting
 * for a leapfrog step are served together).  Requires K <= 208, G == 0; on a row-sharded handle (world > 1)
 * b200glm_comm_init must have been called: every rank sums its rows for all chains and one NCCL all-reduce of the
 * (K + 2) x chains block per batched evaluation combines them (identical on every rank).  Batches of <= 8 lanes with
 * K <= 128 take glm_multi_kernel: four chains per pass on the FMA path (5-8 lanes: two passes), HBM-bound like the
 * single-chain kernel.
 * All per-chain arrays are chain-major HOST arrays: theta[i*P + k] belongs to lane i.
 *   batch_reserve            allocate state for chain slots [0, max_chains) (inverse metric = 1)
 *   log_prob_grad_batched    == n calls of b200glm_log_prob_grad; status[i] (may be NULL) gets the
 *                            per-chain code (OK / DOMAIN); with status == NULL any DOMAIN is returned
 *   set_state_batched        upload z = (q, p, g, V) [and the diagonal inverse metric] of n chain slots
 *   leapfrog_batched         lane i advances chain slot chains[i] (NULL: slot i) by eps[i]; mirrors the
 *                            new (q, p, g, V) to the host (any of q, p, g, V, status may be NULL).
 *                            Domain error in a lane: V = +inf, g negated, status[i] = DOMAIN.
 *   leapfrog_batched_async   slots [0, n) by the same eps, device-resident (bench `value`) */
int b200glm_batch_reserve(b200glm_handle* h, int32_t max_chains);
int b200glm_log_prob_grad_batched(b200glm_handle* h, int32_t n, const double* theta, int32_t propto,
                                  int32_t jacobian, double* lp, double* grad, int32_t* status);
int b200glm_set_state_batched(b200glm_handle* h, int32_t n, const int32_t* chains, const double* q,
                              const double* p, const double* g, const double* V, const double* inv_metric);
int b200glm_leapfrog_batched(b200glm_handle* h, int32_t n, const int32_t* chains, const double* eps,
                             double* q, double* p, double* g, double* V, int32_t* status);
int b200glm_leapfrog_batched_async(b200glm_handle* h, int32_t n, double eps);
int b200glm_batch_sync(b200glm_handle* h);
void* b200glm_batch_stream(b200glm_handle* h);

/* Device-side NUTS transition + adaptation for the batch.  Replaces, per chain,
 * stan::mcmc::adapt_diag_e_nuts<Model, rng_t>::transition (ST/mcmc/hmc/nuts/adapt_diag_e_nuts.hpp:26-44), i.e.
 * base_nuts::transition + build_tree (base_nuts.hpp:78-204, 247-352), base_hmc::init_stepsize (base_hmc.hpp:78-143),
 * stepsize_adaptation::learn_stepsize / complete_adaptation (stepsize_adaptation.hpp:55-71) and
 * var_adaptation::learn_variance (var_adaptation.hpp:17-46), run by a per-chain state machine right behind the batched
 * leapfrog: q, p, g, the tree, the dual-averaging state and the Welford accumulators stay on the device; only draws,
 * 40 bytes of status per chain and round, and the adapted metric leave it.  Randomness stays the caller's (the
 * reference's per-chain boost engine): P normal variates per momentum refresh and the uniform variates of the direction
 * / multinomial decisions, written into the pinned buffers below in the order the reference draws them; the status
 * says how many uniform variates were consumed.  Host driver: b200::hmc_nuts_diag_e_adapt_device
 * (stan_b200/cpp/b200/device_nuts.hpp), the argument list of ST/services/sample/hmc_nuts_diag_e_adapt.hpp:331-404.
 *   nuts_reserve     batch_reserve(n_chains) + the per-chain tree state ((19 + 6 max_depth) P doubles)
 *   nuts_buffers     pinned host buffers shared with the kernels: normals [n][P], uniforms [n][72] (entries [0, 64): a ring indexed
 *                    by the running count; entry 64: the step-size jitter variate of the transition about to start), status [n], draws [n][3 P + 8] (parameters, lp__, accept_stat__, stepsize__,
 *                    treedepth__, n_leapfrog__, divergent__, energy__, iteration, then the selected state's momentum
 *                    and gradient: the diagnostic writer's columns), metric [n][P]
 *   nuts_init_chain  initial point, diagonal inverse metric, nominal step size of one chain
 *   nuts_round       one round for the listed chains: begin (chains whose normal variates were supplied) ->
 *                    ONE batched leapfrog -> tree / adaptation step; returns when the status is up to date */
typedef struct b200glm_nuts_config {
  int32_t max_depth, num_warmup, num_samples;
  /* stan::mcmc::windowed_adaptation after set_window_params(): num_warmup_, adapt_init_buffer_, adapt_term_buffer_,
   * adapt_base_window_, adapt_window_size_, adapt_next_window_ */
  uint32_t w_num_warmup, w_init_buffer, w_term_buffer, w_base_window, w_size0, w_next0;
  double max_deltaH, delta, gamma, kappa, t0;
  double stepsize_jitter;   /* base_hmc::epsilon_jitter_ (base_hmc.hpp:195-200); appended in ABI 5 */
} b200glm_nuts_config;
typedef struct b200glm_nuts_status {
  int32_t phase;          /* 1 initial gradient, 2 / 3 init_stepsize, 4 inside a transition, 5 done, 6 failed */
  int32_t need_normals;   /* the chain waits for P fresh normal variates in its row of `normals` */
  int32_t iter;           /* transitions completed */
  int32_t fail_code;      /* 1 posterior improper, 2 no acceptably small step size, 3 metric overflow (base_hmc.hpp:131-140) */
  int32_t adapt_done, reserved;
  uint64_t n_unif;        /* uniform variates consumed so far */
  double eps_nom;         /* nominal step size */
} b200glm_nuts_status;
int b200glm_nuts_reserve(b200glm_handle* h, int32_t n_chains, const b200glm_nuts_config* cfg);
int b200glm_nuts_buffers(b200glm_handle* h, double** normals, double** uniforms, b200glm_nuts_status** status,
                         double** draws, double** metric);
int b200glm_nuts_init_chain(b200glm_handle* h, int32_t chain, const double* q0, const double* inv_metric,
                            double stepsize);
int b200glm_nuts_round(b200glm_handle* h, int32_t n_lanes, const int32_t* chains);

/* Row-sharded operation: one process per GPU, likelihood partials combined by one NCCL
 * all-reduce of P+2 doubles per gradient (replaces nothing in the reference's GLM path; the
 * analogue is map_rect's gatherv, SM/prim/functor/mpi_parallel_call.hpp:354-392).
 * unique_id is the 128-byte ncclUniqueId obtained on rank 0 and broadcast by the caller. */
int b200glm_comm_unique_id(void* unique_id_128);
int b200glm_comm_init(b200glm_handle* h, const void* unique_id_128, int32_t rank, int32_t world);

/* Row-sharded operation WITHOUT a collective launch (preferred on NVLink / NVSwitch boxes): every rank
 * exports a small mailbox in its HBM (peer_export: 64-byte cudaIpcMemHandle_t), the caller gathers the
 * world's handles (rank order) and hands them to peer_connect.  From then on the LAST CTA of each
 * gradient launch pushes its likelihood partials into every peer's mailbox, waits for theirs and sums
 * in rank order, so a gradient (or fused leapfrog step) is ONE launch per rank and no NCCL call; theta
 * stays bitwise replicated.  All ranks must issue the same sequence of evaluations per slot; a rank
 * that does not makes the others return B200GLM_CUDA after a 20 s timeout (no hang).
 * The poisson propto=false constant sum lgamma(y+1) is per shard: read it with lgamma_sum_local, add
 * over ranks on the host and store the total with set_lgamma_sum_total (comm_init does this itself). */
int b200glm_peer_export(b200glm_handle* h, void* ipc_handle_64);
int b200glm_peer_connect(b200glm_handle* h, const void* all_handles, int32_t world);
double b200glm_lgamma_sum_local(const b200glm_handle* h);
int b200glm_set_lgamma_sum_total(b200glm_handle* h, double total);
/* Both create-time per-shard constants at once: out[0] = the propto=false constant of this shard (poisson_log,
 * binomial_logit, neg_binomial_2_log; 0 otherwise), out[1] = 1 if this shard holds an out-of-range y (the data
 * checks bernoulli_logit_glm_lpmf.hpp:85, poisson_log_glm_lpmf.hpp:84, ... perform on every call).  Sum both over
 * the ranks and store the totals, so that every rank subtracts the same constant and reports the same
 * B200GLM_DOMAIN status (b200glm_comm_init does this itself over NCCL). */
int b200glm_shard_constants_local(const b200glm_handle* h, double out[2]);
int b200glm_set_shard_constants_total(b200glm_handle* h, const double total[2]);

/* launch accounting for the bench (`gpu_launches`) and algorithmic bytes per gradient */
int64_t b200glm_launch_count(const b200glm_handle* h);
int64_t b200glm_bytes_per_gradient(const b200glm_handle* h);
/* Per-phase time stamps of a slot's gradient launches (measurement only; narrow kernel + the shared tail).
 * enable(on=1) allocates the buffer, from then on every launch of the slot overwrites it; read copies the
 * stamps of the LAST launch: rows = grid + 2 rows of 16 words, word k = %globaltimer in ns (comparable across
 * the SMs and GPUs of one box), word 8 + k = clock64 of that SM.  Rows [0, grid) are the CTAs: 0 entry,
 * 1 previous launch complete (griddepcontrol.wait over), 2 theta staged, 3 first panel landed, 4 last panel
 * consumed by every warp, 5 partial row written + ticket taken.  Row `grid` is the last CTA's tail: 0 ticket
 * won (word 7 = its CTA id), 1 sum of the grid's partial rows done, 2 peers' partials received and summed,
 * 3 model epilogue / leapfrog tail written.  Row grid + 1 refines the tail (word k only): 0 acquire fence after the
 * ticket, 1 first batch of partial rows loaded, 2 sums written, 3 epilogue: block sums + value done, 4 gradient and
 * leapfrog tail written.  read(out = NULL) only returns the row count. */
int b200glm_timeline_enable(b200glm_handle* h, int32_t slot, int32_t on);
int b200glm_timeline_read(b200glm_handle* h, int32_t slot, uint64_t* out, int32_t* rows);
/* Roofline denominators measured on `device` in the caller's process (either pointer may be NULL):
 * read_gbs    = bandwidth of a read-only stream over 4 GiB (the gradient kernels read X once and write nothing;
 *               the driver's MEASURED_PEAKS hbm_gbs is a read+write copy);
 * dmma_tflops = register-resident mma.sync.m8n8k4.f64 loop, the fp64 tensor-pipe peak that bounds the batched kernel. */
int b200glm_measure_peaks(int32_t device, double* read_gbs, double* dmma_tflops);
const char* b200glm_last_error(const b200glm_handle* h);
const char* b200glm_version(void);
/* B200GLM_ABI_VERSION the library was built with; bindings refuse to run on a mismatch */
int32_t b200glm_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif
