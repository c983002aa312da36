// Device-side NUTS transition + adaptation for a batch of chains: the per-chain state machine.
//
// What it restates (reference files relative to src/stan):
//   mcmc/hmc/nuts/base_nuts.hpp:78-204    transition(): doubling loop, multinomial sampling across subtrees, U-turn checks
//   mcmc/hmc/nuts/base_nuts.hpp:247-352   build_tree(): here ITERATIVE -- one leapfrog per call, the recursion's frames are
//                                         an explicit stack of completed-subtree summaries (one per level)
//   mcmc/hmc/base_hmc.hpp:78-143          init_stepsize(): the doubling / halving search, one iteration per call
//   mcmc/hmc/hamiltonians/diag_e_metric.hpp:20-50   T(z), dtau_dp, sample_p (normal variates come from the host's engine)
//   mcmc/stepsize_adaptation.hpp:55-71    learn_stepsize / complete_adaptation (dual averaging)
//   mcmc/var_adaptation.hpp:17-46, windowed_adaptation.hpp:82-113, math welford_var_estimator.hpp   metric windows
//   mcmc/hmc/nuts/adapt_diag_e_nuts.hpp:26-44   what follows a warm-up transition
//
// A chain advances by ROUNDS: [nuts_begin: start a transition or an init_stepsize iteration when the host has supplied the
// normal variates it needs] -> one leapfrog step of every live chain (the batched gradient kernels) -> [nuts_after_leapfrog:
// everything the reference does between two evolve() calls].  The leapfrog state (q, p, g, V, inverse metric) is the batched
// kernels' feature-major chain slot; the tree state lives in a chain-major block of (19 + 6 max_depth) P doubles.
// Randomness is the reference's: P normal variates per momentum refresh and the uniform variates of the direction /
// multinomial draws are produced by the host from the chain's own boost engine IN THE REFERENCE'S ORDER and consumed here
// in that order, so a chain reproduces the reference's draws for the same seed (up to summation order in dot products).
//
// The code is lane-generic: on the device one warp works on one chain (LN = lanes of a warp, features strided over lanes,
// sums by shuffles); tests/ build the same header for the host with a one-lane policy and run it against the reference's
// sampler without a GPU.  Scalar logic is computed redundantly by every lane from identical inputs (no broadcasts); lane 0
// publishes it.
#pragma once
#include <math.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define NT_HD __host__ __device__ __forceinline__
#else
#define NT_HD inline
#endif

namespace b200glm {

constexpr int NUTS_DEPTH_CAP = 16;    // max_depth <= 16
constexpr int NUTS_UNIF_CAP = 64;     // per-chain ring of uniform variates (a round consumes at most max_depth + 3)
constexpr int NUTS_UNIF_STRIDE = NUTS_UNIF_CAP + 8;   // a chain's row of `uniforms`: the ring, then at [NUTS_UNIF_CAP] the
                                      // step-size jitter variate of the transition about to start (drawn BEFORE the
                                      // momentum: base_nuts.hpp:80 precedes :84), then padding to a 64-byte multiple
constexpr int NUTS_DRAW_EXTRA = 8;    // after the P parameters: lp__, accept_stat__, stepsize__, treedepth__, n_leapfrog__,
                                      // divergent__, energy__, iteration; then the selected state's momentum (P) and
                                      // gradient (P) for the diagnostic writer (ps_point::get_params)
NT_HD size_t nuts_draw_doubles(int P) { return (size_t)3 * P + NUTS_DRAW_EXTRA; }

enum { NPH_IDLE = 0, NPH_INIT_GRAD = 1, NPH_SS_FIRST = 2, NPH_SS_LOOP = 3, NPH_TREE = 4, NPH_DONE = 5, NPH_FAILED = 6 };
enum { NFAIL_NONE = 0, NFAIL_IMPROPER = 1, NFAIL_NO_STEPSIZE = 2, NFAIL_METRIC_OVERFLOW = 3 };

struct NutsConfig {
  int P, max_depth;
  double max_deltaH;                     // base_nuts: 1000
  double delta, gamma, kappa, t0;        // stepsize_adaptation
  // windowed_adaptation members as set_window_params() leaves them (all zero when num_warmup < 20)
  unsigned w_num_warmup, w_init_buffer, w_term_buffer, w_base_window;
  // ... and its window cursor at that point: set_window_params() only calls restart() on its regular path, so with the
  // 15 % / 75 % / 10 % schedule the cursor is still the constructor's (size 0, next window = UINT_MAX: no metric update ever)
  unsigned w_size0, w_next0;
  int num_warmup, num_samples;           // transitions with / without adaptation
  double stepsize_jitter;                // base_hmc::epsilon_jitter_ (0: sample_stepsize() draws nothing)
};

// offsets (in units of P doubles) into a chain's vector block
enum {
  NV_ZE = 0,      // trajectory ends: [0] backward (q, p, g), [1] forward (q, p, g)
  NV_RHO = 6,     // rho of the whole trajectory
  NV_QS = 7,      // current sample z_sample.q (between transitions: the chain's position)
  NV_GS = 8,      // ... its gradient
  NV_PS = 9,      // ... its momentum (only reported: the diagnostic writer's p columns)
  NV_IM = 10,     // chain-major copy of the diagonal inverse metric
  NV_WM = 11,     // Welford mean
  NV_WM2 = 12,    // Welford sum of squares
  NV_CUR = 13,    // summary of the subtree being completed: rho, p_beg, p_end, z_propose (q, g, p)
  NV_STACK = 19   // [level] summaries of completed subtrees waiting for their sibling
};
enum { NS_RHO = 0, NS_PBEG = 1, NS_PEND = 2, NS_QP = 3, NS_GP = 4, NS_PP = 5, NS_VECS = 6 };
NT_HD size_t nuts_vec_doubles(int P, int max_depth) { return (size_t)(NV_STACK + NS_VECS * max_depth) * (size_t)P; }

struct NutsChain {
  int phase, need_normals, iter, depth, dir, leaf, n_leapfrog, divergent;
  int ss_direction, after_ss, adapt_done, fail_code;
  unsigned long long n_unif;          // uniform variates consumed so far (index into the ring)
  double eps_nom, eps, H0, lsw, sum_metro;
  double Vs, hs;                      // z_sample: potential and Hamiltonian
  double V_end[2];
  double cur_lsw, cur_V, cur_h;
  double st_lsw[NUTS_DEPTH_CAP], st_V[NUTS_DEPTH_CAP], st_h[NUTS_DEPTH_CAP];
  double sa_counter, sa_sbar, sa_xbar, sa_mu;
  unsigned w_counter, w_next, w_size;
  double wf_n;
  double lane_eps;                    // step of the chain's next leapfrog lane: 0 (gradient only), eps_nom, +-eps
};

// what the host sees after every round (pinned memory)
struct NutsStatus {
  int phase, need_normals, iter, fail_code;
  int adapt_done, reserved;
  unsigned long long n_unif;
  double eps_nom;
};

// the batched kernels' chain slot (feature-major, one column per chain)
struct NutsSlot {
  double *Q, *Pm, *Gd, *IM, *V;
  size_t ld;
  int c;
  NT_HD double& q(int k) const { return Q[(size_t)k * ld + c]; }
  NT_HD double& p(int k) const { return Pm[(size_t)k * ld + c]; }
  NT_HD double& g(int k) const { return Gd[(size_t)k * ld + c]; }
  NT_HD double& im(int k) const { return IM[(size_t)k * ld + c]; }
  NT_HD double& v() const { return V[c]; }
};

// one-lane policy (host build; a device thread working alone)
struct NutsOneLane {
  static NT_HD int lane() { return 0; }
  static NT_HD int lanes() { return 1; }
  static NT_HD double sum(double v) { return v; }
};

// SM/prim/fun/log1p_exp.hpp, log_sum_exp.hpp
NT_HD double nuts_log1p_exp(double a) { return a > 0.0 ? a + log1p(exp(-a)) : log1p(exp(a)); }
NT_HD double nuts_log_sum_exp(double a, double b) {
  if (a == -INFINITY) return b;
  if (a == INFINITY && b == INFINITY) return INFINITY;
  if (a > b) return a + nuts_log1p_exp(b - a);
  return b + nuts_log1p_exp(a - b);
}

NT_HD void nuts_window_restart(const NutsConfig& cfg, NutsChain& ch) {   // the cursor the reference starts warm-up with
  ch.w_counter = 0;
  ch.w_size = cfg.w_size0;
  ch.w_next = cfg.w_next0;
}

// state of a chain before its first round; the caller has put q0 into the slot's Q column (p = g = 0) and the inverse
// metric into its IM column
template <class LN>
NT_HD void nuts_chain_init(const NutsConfig& cfg, NutsChain& ch, double* v, const NutsSlot& s, double stepsize) {
  const int P = cfg.P;
  for (int k = LN::lane(); k < P; k += LN::lanes()) {
    v[(size_t)NV_IM * P + k] = s.im(k);
    v[(size_t)NV_WM * P + k] = 0.0;
    v[(size_t)NV_WM2 * P + k] = 0.0;
  }
  ch.phase = NPH_INIT_GRAD;
  ch.need_normals = 0;
  ch.iter = 0;
  ch.depth = ch.dir = ch.leaf = ch.n_leapfrog = ch.divergent = 0;
  ch.ss_direction = 0;
  ch.after_ss = 0;
  ch.adapt_done = 0;
  ch.fail_code = NFAIL_NONE;
  ch.n_unif = 0;
  ch.eps_nom = stepsize > 0 ? stepsize : 0.1;   // base_hmc: nom_epsilon_(0.1), set_nominal_stepsize ignores e <= 0
  ch.eps = ch.eps_nom;
  ch.H0 = ch.lsw = ch.sum_metro = 0.0;
  ch.Vs = ch.hs = 0.0;
  ch.sa_counter = ch.sa_sbar = ch.sa_xbar = 0.0;
  ch.sa_mu = log(10 * stepsize);                // hmc_nuts_diag_e_adapt.hpp:94 (the caller's stepsize, before init_stepsize)
  ch.wf_n = 0.0;
  nuts_window_restart(cfg, ch);
  ch.lane_eps = 0.0;
}

NT_HD void nuts_next_or_done(const NutsConfig& cfg, NutsChain& ch) {
  if (!ch.adapt_done && ch.iter == cfg.num_warmup) {   // disengage_adaptation -> complete_adaptation
    ch.eps_nom = exp(ch.sa_xbar);
    ch.adapt_done = 1;
  }
  if (ch.iter == cfg.num_warmup + cfg.num_samples) {
    ch.phase = NPH_DONE;
    ch.lane_eps = 0.0;
  } else {
    ch.phase = NPH_TREE;
    ch.need_normals = 1;
  }
}

NT_HD void nuts_ss_done(const NutsConfig& cfg, NutsChain& ch) {
  if (ch.after_ss) {   // adapt_diag_e_nuts.hpp:37-40
    ch.sa_mu = log(10 * ch.eps_nom);
    ch.sa_counter = ch.sa_sbar = ch.sa_xbar = 0.0;
  }
  nuts_next_or_done(cfg, ch);
}

NT_HD void nuts_start_init_stepsize(const NutsConfig& cfg, NutsChain& ch, int after) {
  ch.after_ss = after;
  if (ch.eps_nom == 0 || ch.eps_nom > 1e7 || isnan(ch.eps_nom)) {   // base_hmc.hpp:88-91
    nuts_ss_done(cfg, ch);
    return;
  }
  ch.phase = NPH_SS_FIRST;
  ch.need_normals = 1;
}

// Start of a transition (base_nuts.hpp:80-121) or of an init_stepsize iteration (base_hmc.hpp:93-97, 109-114): fresh
// momentum p = normal / sqrt(inv_metric) (diag_e_metric::sample_p), the position and its gradient are the current sample's
// (hamiltonian.init() would recompute the same values), H0.
template <class LN>
NT_HD void nuts_begin(const NutsConfig& cfg, NutsChain& ch, double* v, const NutsSlot& s, const double* normals,
                      const double* unif) {
  const int P = cfg.P;
  const bool tree = ch.phase == NPH_TREE;
  double* ze = v + (size_t)NV_ZE * P;
  double t = 0.0;
  for (int k = LN::lane(); k < P; k += LN::lanes()) {
    const double m = v[(size_t)NV_IM * P + k];
    const double p = normals[k] / sqrt(m);
    const double q = v[(size_t)NV_QS * P + k], g = v[(size_t)NV_GS * P + k];
    s.q(k) = q;
    s.p(k) = p;
    s.g(k) = g;
    t += p * (m * p);
    if (tree) {
      ze[k] = q;
      ze[(size_t)P + k] = p;
      ze[(size_t)2 * P + k] = g;
      ze[(size_t)3 * P + k] = q;
      ze[(size_t)4 * P + k] = p;
      ze[(size_t)5 * P + k] = g;
      v[(size_t)NV_RHO * P + k] = p;
      v[(size_t)NV_PS * P + k] = p;
    }
  }
  t = 0.5 * LN::sum(t);
  if (LN::lane() == 0) s.v() = ch.Vs;
  ch.H0 = t + ch.Vs;
  ch.need_normals = 0;
  if (tree) {
    ch.eps = ch.eps_nom;   // sample_stepsize() (base_hmc.hpp:195-200): the variate is only drawn when jitter != 0
    if (cfg.stepsize_jitter != 0.0) ch.eps *= 1.0 + cfg.stepsize_jitter * (2.0 * unif[NUTS_UNIF_CAP] - 1.0);
    ch.lsw = 0.0;
    ch.n_leapfrog = 0;
    ch.sum_metro = 0.0;
    ch.depth = 0;
    ch.divergent = 0;
    ch.hs = ch.H0;
    ch.V_end[0] = ch.V_end[1] = ch.Vs;
    const double u = unif[ch.n_unif % NUTS_UNIF_CAP];
    ch.n_unif++;
    ch.dir = u > 0.5 ? 1 : 0;
    ch.leaf = 0;
    ch.lane_eps = ch.dir ? ch.eps : -ch.eps;
  } else {
    ch.lane_eps = ch.eps_nom;
  }
}

// The three U-turn criteria of a merge (base_nuts.hpp:173-188 and :336-349) for two adjacent spans A (built first) and B:
//   (sharp(A.p_beg), sharp(B.p_end)) . (rho_A + rho_B),  (sharp(A.p_beg), sharp(B.p_beg)) . (rho_A + B.p_beg),
//   (sharp(A.p_end), sharp(B.p_end)) . (rho_B + A.p_end);   sharp(p) = inv_metric o p (dtau_dp).
// Writes rho_A + rho_B to out_rho (may alias either input) and, when a_pbeg_to is given, A.p_beg there.
template <class LN>
NT_HD bool nuts_merge(int P, const double* im, const double* a_rho, const double* a_pbeg, const double* a_pend,
                      const double* b_rho, const double* b_pbeg, const double* b_pend, double* out_rho,
                      double* a_pbeg_to) {
  double d1 = 0, d2 = 0, d3 = 0, d4 = 0, d5 = 0, d6 = 0;
  for (int k = LN::lane(); k < P; k += LN::lanes()) {
    const double m = im[k];
    const double ra = a_rho[k], rb = b_rho[k];
    const double pab = a_pbeg[k], pae = a_pend[k], pbb = b_pbeg[k], pbe = b_pend[k];
    const double sab = m * pab, sae = m * pae, sbb = m * pbb, sbe = m * pbe;
    const double rs = ra + rb, re1 = ra + pbb, re2 = rb + pae;
    d1 += sab * rs;
    d2 += sbe * rs;
    d3 += sab * re1;
    d4 += sbb * re1;
    d5 += sae * re2;
    d6 += sbe * re2;
    out_rho[k] = rs;
    if (a_pbeg_to) a_pbeg_to[k] = pab;
  }
  d1 = LN::sum(d1);
  d2 = LN::sum(d2);
  d3 = LN::sum(d3);
  d4 = LN::sum(d4);
  d5 = LN::sum(d5);
  d6 = LN::sum(d6);
  return d1 > 0 && d2 > 0 && d3 > 0 && d4 > 0 && d5 > 0 && d6 > 0;
}

template <class LN>
NT_HD void nuts_copy(int P, double* dst, const double* src, int n_vecs) {
  // feature k of every vector is always touched by lane k mod lanes(): no value crosses lanes through memory
  for (int j = 0; j < n_vecs; ++j)
    for (int k = LN::lane(); k < P; k += LN::lanes()) dst[(size_t)j * P + k] = src[(size_t)j * P + k];
}

// End of a transition (base_nuts.hpp:193-203) and what adapt_diag_e_nuts::transition does after it (:29-43).
// draw: nuts_draw_doubles(P) doubles; metric_out: P doubles, rewritten when a window ends.
template <class LN>
NT_HD void nuts_end_transition(const NutsConfig& cfg, NutsChain& ch, double* v, const NutsSlot& s, double* draw,
                               double* metric_out) {
  const int P = cfg.P;
  const double accept = ch.sum_metro / (double)ch.n_leapfrog;
  const double* qs = v + (size_t)NV_QS * P;
  for (int k = LN::lane(); k < P; k += LN::lanes()) {
    draw[k] = qs[k];
    draw[(size_t)P + NUTS_DRAW_EXTRA + k] = v[(size_t)NV_PS * P + k];
    draw[(size_t)2 * P + NUTS_DRAW_EXTRA + k] = v[(size_t)NV_GS * P + k];
  }
  if (LN::lane() == 0) {
    draw[P + 0] = -ch.Vs;
    draw[P + 1] = accept;
    draw[P + 2] = ch.eps;
    draw[P + 3] = ch.depth;
    draw[P + 4] = ch.n_leapfrog;
    draw[P + 5] = ch.divergent;
    draw[P + 6] = ch.hs;
    draw[P + 7] = ch.iter;
  }
  const bool adapting = ch.iter < cfg.num_warmup;
  ch.iter++;
  ch.lane_eps = 0.0;
  if (!adapting) {
    nuts_next_or_done(cfg, ch);
    return;
  }
  // ---- stepsize_adaptation::learn_stepsize ----
  {
    ch.sa_counter += 1.0;
    const double a = accept > 1 ? 1 : accept;
    const double eta = 1.0 / (ch.sa_counter + cfg.t0);
    ch.sa_sbar = (1.0 - eta) * ch.sa_sbar + eta * (cfg.delta - a);
    const double x = ch.sa_mu - ch.sa_sbar * sqrt(ch.sa_counter) / cfg.gamma;
    const double x_eta = pow(ch.sa_counter, -cfg.kappa);
    ch.sa_xbar = (1.0 - x_eta) * ch.sa_xbar + x_eta * x;
    ch.eps_nom = exp(x);
  }
  // ---- var_adaptation::learn_variance ----
  const unsigned wc = ch.w_counter, nw = cfg.w_num_warmup;
  const bool in_window = wc >= cfg.w_init_buffer && wc < nw - cfg.w_term_buffer && wc != nw;
  const bool end_window = wc == ch.w_next && wc != nw;
  double* wm = v + (size_t)NV_WM * P;
  double* wm2 = v + (size_t)NV_WM2 * P;
  double* im = v + (size_t)NV_IM * P;
  if (in_window) ch.wf_n += 1.0;
  const double n = ch.wf_n;
  double bad = 0.0;
  if (in_window || end_window) {
    for (int k = LN::lane(); k < P; k += LN::lanes()) {
      double m = wm[k], m2 = wm2[k];
      if (in_window) {   // welford_var_estimator::add_sample
        const double q = qs[k];
        const double delta = q - m;
        m += delta / n;
        m2 += delta * (q - m);
      }
      if (end_window) {
        double var = im[k];
        if (n > 1) var = m2 / (n - 1.0);
        var = (n / (n + 5.0)) * var + 1e-3 * (5.0 / (n + 5.0));
        if (!isfinite(var)) bad += 1.0;
        im[k] = var;
        s.im(k) = var;
        metric_out[k] = var;
        m = 0.0;
        m2 = 0.0;
      }
      wm[k] = m;
      wm2[k] = m2;
    }
  }
  if (end_window) {
    bad = LN::sum(bad);
    // windowed_adaptation::compute_next_window
    const unsigned last = nw - cfg.w_term_buffer - 1u;
    if (ch.w_next != last) {
      ch.w_size *= 2u;
      ch.w_next = ch.w_counter + ch.w_size;
      if (ch.w_next != last) {
        const unsigned boundary = ch.w_next + 2u * ch.w_size;
        if (boundary >= nw - cfg.w_term_buffer) ch.w_next = last;
      }
    }
    ch.wf_n = 0.0;
    ch.w_counter++;
    if (bad != 0.0) {
      ch.phase = NPH_FAILED;
      ch.fail_code = NFAIL_METRIC_OVERFLOW;
      return;
    }
    nuts_start_init_stepsize(cfg, ch, 1);
    return;
  }
  ch.w_counter++;
  nuts_next_or_done(cfg, ch);
}

// Everything between two leapfrog steps of one chain.  The slot holds the state the leapfrog produced.
template <class LN>
NT_HD void nuts_after_leapfrog(const NutsConfig& cfg, NutsChain& ch, double* v, const NutsSlot& s, const double* unif,
                               double* draw, double* metric_out) {
  const int P = cfg.P;
  const double* im = v + (size_t)NV_IM * P;

  if (ch.phase == NPH_INIT_GRAD) {   // the gradient at the initial point (hamiltonian.init in the first init_stepsize)
    for (int k = LN::lane(); k < P; k += LN::lanes()) {
      v[(size_t)NV_QS * P + k] = s.q(k);
      v[(size_t)NV_GS * P + k] = s.g(k);
    }
    ch.Vs = s.v();
    nuts_start_init_stepsize(cfg, ch, 0);
    return;
  }

  // H(z) = T(z) + V(z) of the new state
  double t = 0.0;
  for (int k = LN::lane(); k < P; k += LN::lanes()) {
    const double p = s.p(k);
    t += p * (im[k] * p);
  }
  const double Vn = s.v();
  double h = 0.5 * LN::sum(t) + Vn;
  if (isnan(h)) h = INFINITY;

  if (ch.phase == NPH_SS_FIRST || ch.phase == NPH_SS_LOOP) {   // base_hmc.hpp:99-137
    const double delta_H = ch.H0 - h;
    const double thr = log(0.8);
    if (ch.phase == NPH_SS_FIRST) {
      ch.ss_direction = delta_H > thr ? 1 : -1;
      ch.phase = NPH_SS_LOOP;
      ch.need_normals = 1;
      return;
    }
    if ((ch.ss_direction == 1 && !(delta_H > thr)) || (ch.ss_direction == -1 && !(delta_H < thr))) {
      nuts_ss_done(cfg, ch);
      return;
    }
    ch.eps_nom = ch.ss_direction == 1 ? 2.0 * ch.eps_nom : 0.5 * ch.eps_nom;
    if (ch.eps_nom > 1e7) {
      ch.phase = NPH_FAILED;
      ch.fail_code = NFAIL_IMPROPER;
      return;
    }
    if (ch.eps_nom == 0) {
      ch.phase = NPH_FAILED;
      ch.fail_code = NFAIL_NO_STEPSIZE;
      return;
    }
    ch.need_normals = 1;
    return;
  }

  // ---------------- NPH_TREE: a new leaf (build_tree base case, base_nuts.hpp:253-282) ----------------
  ch.n_leapfrog++;
  if ((h - ch.H0) > cfg.max_deltaH) ch.divergent = 1;
  const double w = ch.H0 - h;
  ch.sum_metro += w > 0 ? 1.0 : exp(w);
  if (ch.divergent) {   // every enclosing build_tree returns false, transition() leaves its loop
    nuts_end_transition<LN>(cfg, ch, v, s, draw, metric_out);
    return;
  }
  double* cur = v + (size_t)NV_CUR * P;
  for (int k = LN::lane(); k < P; k += LN::lanes()) {
    const double p = s.p(k);
    cur[(size_t)NS_RHO * P + k] = p;
    cur[(size_t)NS_PBEG * P + k] = p;
    cur[(size_t)NS_PEND * P + k] = p;
    cur[(size_t)NS_QP * P + k] = s.q(k);
    cur[(size_t)NS_GP * P + k] = s.g(k);
    cur[(size_t)NS_PP * P + k] = p;
  }
  ch.cur_lsw = nuts_log_sum_exp(-INFINITY, w);
  ch.cur_V = Vn;
  ch.cur_h = h;

  // ---- complete every enclosing subtree this leaf is the last leaf of (the recursion's unwinding, :284-351) ----
  int level = 0;
  while ((ch.leaf >> level) & 1) {
    double* a = v + (size_t)(NV_STACK + NS_VECS * level) * P;   // the initial half (built first)
    const double lsw_sub = nuts_log_sum_exp(ch.st_lsw[level], ch.cur_lsw);
    bool take_final;
    if (ch.cur_lsw > lsw_sub) {
      take_final = true;
    } else {
      const double u = unif[ch.n_unif % NUTS_UNIF_CAP];
      ch.n_unif++;
      take_final = u < exp(ch.cur_lsw - lsw_sub);
    }
    const bool persist = nuts_merge<LN>(P, im, a + (size_t)NS_RHO * P, a + (size_t)NS_PBEG * P, a + (size_t)NS_PEND * P,
                                        cur + (size_t)NS_RHO * P, cur + (size_t)NS_PBEG * P, cur + (size_t)NS_PEND * P,
                                        cur + (size_t)NS_RHO * P, cur + (size_t)NS_PBEG * P);
    if (!take_final) {
      nuts_copy<LN>(P, cur + (size_t)NS_QP * P, a + (size_t)NS_QP * P, 3);
      ch.cur_V = ch.st_V[level];
      ch.cur_h = ch.st_h[level];
    }
    ch.cur_lsw = lsw_sub;
    if (!persist) {   // invalid subtree
      nuts_end_transition<LN>(cfg, ch, v, s, draw, metric_out);
      return;
    }
    ++level;
  }
  ch.leaf++;
  if (level < ch.depth) {   // an initial half is complete: park it until its sibling is
    nuts_copy<LN>(P, v + (size_t)(NV_STACK + NS_VECS * level) * P, cur, NS_VECS);
    ch.st_lsw[level] = ch.cur_lsw;
    ch.st_V[level] = ch.cur_V;
    ch.st_h[level] = ch.cur_h;
    return;   // next leaf, same direction
  }

  // ---------------- the new subtree of 2^depth leaves is complete and valid (base_nuts.hpp:156-191) ----------------
  ch.depth++;
  bool take;
  if (ch.cur_lsw > ch.lsw) {
    take = true;
  } else {
    const double u = unif[ch.n_unif % NUTS_UNIF_CAP];
    ch.n_unif++;
    take = u < exp(ch.cur_lsw - ch.lsw);
  }
  if (take) {
    nuts_copy<LN>(P, v + (size_t)NV_QS * P, cur + (size_t)NS_QP * P, 1);
    nuts_copy<LN>(P, v + (size_t)NV_GS * P, cur + (size_t)NS_GP * P, 1);
    nuts_copy<LN>(P, v + (size_t)NV_PS * P, cur + (size_t)NS_PP * P, 1);
    ch.Vs = ch.cur_V;
    ch.hs = ch.cur_h;
  }
  ch.lsw = nuts_log_sum_exp(ch.lsw, ch.cur_lsw);
  double* ze = v + (size_t)NV_ZE * P;
  double* end_this = ze + (size_t)3 * P * ch.dir;          // the end that was extended (still its old state)
  double* end_other = ze + (size_t)3 * P * (1 - ch.dir);
  const bool persist = nuts_merge<LN>(P, im, v + (size_t)NV_RHO * P, end_other + P, end_this + P, cur + (size_t)NS_RHO * P,
                                      cur + (size_t)NS_PBEG * P, cur + (size_t)NS_PEND * P, v + (size_t)NV_RHO * P,
                                      (double*)0);
  for (int k = LN::lane(); k < P; k += LN::lanes()) {      // z_fwd / z_bck = z_
    end_this[k] = s.q(k);
    end_this[(size_t)P + k] = s.p(k);
    end_this[(size_t)2 * P + k] = s.g(k);
  }
  ch.V_end[ch.dir] = Vn;
  if (!persist || ch.depth >= cfg.max_depth) {
    nuts_end_transition<LN>(cfg, ch, v, s, draw, metric_out);
    return;
  }
  // next doubling: direction, and the integrator state moved to that end of the trajectory
  const double u = unif[ch.n_unif % NUTS_UNIF_CAP];
  ch.n_unif++;
  const int nd = u > 0.5 ? 1 : 0;
  if (nd != ch.dir) {
    const double* e = ze + (size_t)3 * P * nd;
    for (int k = LN::lane(); k < P; k += LN::lanes()) {
      s.q(k) = e[k];
      s.p(k) = e[(size_t)P + k];
      s.g(k) = e[(size_t)2 * P + k];
    }
    if (LN::lane() == 0) s.v() = ch.V_end[nd];
  }
  ch.dir = nd;
  ch.leaf = 0;
  ch.lane_eps = nd ? ch.eps : -ch.eps;
}

NT_HD void nuts_publish(const NutsChain& ch, NutsStatus& st) {
  st.phase = ch.phase;
  st.need_normals = ch.need_normals;
  st.iter = ch.iter;
  st.fail_code = ch.fail_code;
  st.adapt_done = ch.adapt_done;
  st.reserved = 0;
  st.n_unif = ch.n_unif;
  st.eps_nom = ch.eps_nom;
}

}  // namespace b200glm
