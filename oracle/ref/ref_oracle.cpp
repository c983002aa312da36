// oracle/ref/ref_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/glm_oracle.h).
//
// C entry points over the UNMODIFIED reference, compiled from /root/reference by
// oracle/Makefile into oracle/_ref/libref_oracle_<isa>.so:
//   ref_glm_log_prob_grad -> stan::model::log_prob_grad<propto,jacobian>  (src/stan/model/log_prob_grad.hpp:29-50)
//   ref_glm_log_prob      -> Model::log_prob<propto,jacobian>(double)     (as called at services/util/initialize.hpp:128)
//   ref_glm_gradient      -> stan::model::gradient                        (src/stan/model/gradient.hpp:22-35)
//   ref_glm_leapfrog      -> expl_leapfrog<diag_e_metric>::evolve         (mcmc/hmc/integrators/base_leapfrog.hpp:17-22)
//   ref_glm_nuts          -> services::sample::hmc_nuts_diag_e_adapt      (services/sample/hmc_nuts_diag_e_adapt.hpp:58,331)
//   ref_glm_nuts_device_host -> the PRODUCT's device-NUTS driver + state machine (b200/device_nuts.hpp, nuts_tree.cuh) built for
//                            the host over the reference's model and integrator (nuts_host_backend.hpp): CPU check of SURVEY 8f row 2
//   ref_ess / ref_mcse_*  -> stan::analyze::{ess, mcse_mean, mcse_sd, rhat, split_rank_normalized_ess}
#include "ref_glm_model.hpp"
#include "nuts_host_backend.hpp"

#include <stan/analyze/mcmc/ess.hpp>
#include <stan/analyze/mcmc/mcse.hpp>
#include <stan/analyze/mcmc/rhat.hpp>
#include <stan/analyze/mcmc/split_rank_normalized_ess.hpp>
#include <stan/callbacks/interrupt.hpp>
#include <stan/callbacks/logger.hpp>
#include <stan/callbacks/structured_writer.hpp>
#include <stan/callbacks/writer.hpp>
#include <stan/io/empty_var_context.hpp>
#include <stan/mcmc/hmc/hamiltonians/diag_e_metric.hpp>
#include <stan/mcmc/hmc/integrators/expl_leapfrog.hpp>
#include <stan/model/gradient.hpp>
#include <stan/model/log_prob_grad.hpp>
#include <stan/services/sample/hmc_nuts_diag_e_adapt.hpp>
#include <stan/services/util/create_unit_e_diag_inv_metric.hpp>

#include <chrono>
#include <cstring>
#include <memory>
#include <mutex>
#include <sstream>

namespace {

using oracle_ref::ref_glm_model;

void set_err(char* buf, int len, const char* msg) {
  if (buf && len > 0) {
    std::strncpy(buf, msg, len - 1);
    buf[len - 1] = 0;
  }
}

template <typename F>
int guarded(char* err, int errlen, F&& f) {
  // With STAN_THREADS the AD tape is thread_local and must be created in every thread that
  // evaluates a density (lib/stan_math/stan/math/rev/core/autodiffstackstorage.hpp:17-21, 60-130;
  // TBB workers get theirs from init_chainablestack.hpp, foreign threads need this).
  static thread_local stan::math::ChainableStack thread_tape;
  try {
    f();
    return 0;
  } catch (const std::domain_error& e) {
    set_err(err, errlen, e.what());
    return 1;
  } catch (const std::invalid_argument& e) {
    set_err(err, errlen, e.what());
    return 2;
  } catch (const std::exception& e) {
    set_err(err, errlen, e.what());
    return 3;
  }
}

struct mem_writer : public stan::callbacks::writer {
  std::vector<std::string> names;
  std::vector<std::vector<double>> rows;
  void operator()(const std::vector<std::string>& n) override { names = n; }
  void operator()(const std::vector<double>& s) override { rows.push_back(s); }
  void operator()() override {}
  void operator()(const std::string&) override {}
};

struct mem_metric_writer : public stan::callbacks::structured_writer {
  double stepsize = 0;
  Eigen::VectorXd inv_metric;
  void write(const std::string& key, double value) override {
    if (key == "stepsize")
      stepsize = value;
  }
  void write(const std::string& key, const Eigen::VectorXd& vec) override {
    if (key == "inv_metric")
      inv_metric = vec;
  }
};

struct err_logger : public stan::callbacks::logger {
  std::mutex m;
  std::string errors;
  void error(const std::string& s) override {
    std::lock_guard<std::mutex> g(m);
    errors += s + "\n";
  }
  void error(const std::stringstream& s) override { error(s.str()); }
  void fatal(const std::string& s) override { error(s); }
  void fatal(const std::stringstream& s) override { error(s.str()); }
};

template <bool propto, bool jacobian>
double lp_grad(const ref_glm_model& m, std::vector<double>& th,
               std::vector<double>& grad) {
  std::vector<int> pi;
  return stan::model::log_prob_grad<propto, jacobian>(m, th, pi, grad, nullptr);
}
template <bool propto, bool jacobian>
double lp_only(const ref_glm_model& m, std::vector<double>& th) {
  std::vector<int> pi;
  return m.template log_prob<propto, jacobian>(th, pi, nullptr);
}

}  // namespace

extern "C" {

void* ref_glm_create(const glm_spec* s) {
  try {
    return new ref_glm_model(*s);
  } catch (...) {
    return nullptr;
  }
}
void ref_glm_destroy(void* h) { delete static_cast<ref_glm_model*>(h); }
int ref_glm_num_params(void* h) {
  return static_cast<int>(static_cast<ref_glm_model*>(h)->num_params_r());
}

int ref_glm_log_prob_grad(void* h, const double* theta, int propto,
                          int jacobian, double* lp, double* grad, char* err,
                          int errlen) {
  auto& m = *static_cast<ref_glm_model*>(h);
  const size_t P = m.num_params_r();
  return guarded(err, errlen, [&] {
    std::vector<double> th(theta, theta + P), g;
    double v;
    if (propto)
      v = jacobian ? lp_grad<true, true>(m, th, g) : lp_grad<true, false>(m, th, g);
    else
      v = jacobian ? lp_grad<false, true>(m, th, g)
                   : lp_grad<false, false>(m, th, g);
    *lp = v;
    if (grad)
      std::memcpy(grad, g.data(), P * sizeof(double));
  });
}

int ref_glm_log_prob(void* h, const double* theta, int propto, int jacobian,
                     double* lp, char* err, int errlen) {
  auto& m = *static_cast<ref_glm_model*>(h);
  const size_t P = m.num_params_r();
  return guarded(err, errlen, [&] {
    std::vector<double> th(theta, theta + P);
    if (propto)
      *lp = jacobian ? lp_only<true, true>(m, th) : lp_only<true, false>(m, th);
    else
      *lp = jacobian ? lp_only<false, true>(m, th)
                     : lp_only<false, false>(m, th);
  });
}

int ref_glm_gradient(void* h, const double* theta, double* lp, double* grad,
                     char* err, int errlen) {
  auto& m = *static_cast<ref_glm_model*>(h);
  const size_t P = m.num_params_r();
  return guarded(err, errlen, [&] {
    Eigen::VectorXd x = Eigen::Map<const Eigen::VectorXd>(theta, P), g;
    double f;
    stan::callbacks::logger logger;
    stan::model::gradient(m, x, f, g, logger);
    *lp = f;
    std::memcpy(grad, g.data(), P * sizeof(double));
  });
}

// One step of the reference integrator on a diag_e_point.  On entry g,V may be
// anything if init != 0 (hamiltonian.init recomputes them as base_nuts.hpp:85 does).
int ref_glm_leapfrog(void* h, double eps, const double* inv_metric, int init,
                     double* q, double* p, double* g, double* V, char* err,
                     int errlen) {
  auto& m = *static_cast<ref_glm_model*>(h);
  const int P = static_cast<int>(m.num_params_r());
  return guarded(err, errlen, [&] {
    using H = stan::mcmc::diag_e_metric<ref_glm_model, stan::rng_t>;
    H ham(m);
    stan::mcmc::expl_leapfrog<H> integrator;
    stan::mcmc::diag_e_point z(P);
    stan::callbacks::logger logger;
    z.q = Eigen::Map<const Eigen::VectorXd>(q, P);
    z.p = Eigen::Map<const Eigen::VectorXd>(p, P);
    z.g = Eigen::Map<const Eigen::VectorXd>(g, P);
    z.V = *V;
    if (inv_metric)
      z.inv_e_metric_ = Eigen::Map<const Eigen::VectorXd>(inv_metric, P);
    if (init)
      ham.init(z, logger);
    integrator.evolve(z, ham, eps, logger);
    std::memcpy(q, z.q.data(), P * sizeof(double));
    std::memcpy(p, z.p.data(), P * sizeof(double));
    std::memcpy(g, z.g.data(), P * sizeof(double));
    *V = z.V;
  });
}

// stepsize_jitter of the two NUTS entry points below (set before the call; kept out of their argument lists)
static double g_stepsize_jitter = 0.0;
void ref_set_stepsize_jitter(double j) { g_stepsize_jitter = j; }
// gamma, kappa, t0, init_buffer, term_buffer, window of the two NUTS entry points below (the services' defaults)
static double g_gamma = 0.05, g_kappa = 0.75, g_t0 = 10.0;
static unsigned g_init_buffer = 75, g_term_buffer = 50, g_window = 25;
void ref_set_adapt_params(double gamma, double kappa, double t0, unsigned init_buffer, unsigned term_buffer,
                          unsigned window) {
  g_gamma = gamma;
  g_kappa = kappa;
  g_t0 = t0;
  g_init_buffer = init_buffer;
  g_term_buffer = term_buffer;
  g_window = window;
}
// initial diagonal inverse metric of every chain for the two NUTS entry points below (n = 0: the unit metric)
static std::vector<double> g_init_inv_metric;
void ref_set_init_inv_metric(const double* m, int n) { g_init_inv_metric.assign(m, m + (n > 0 ? n : 0)); }
static stan::io::array_var_context init_metric_context(int P) {
  if (static_cast<int>(g_init_inv_metric.size()) != P)
    return stan::services::util::create_unit_e_diag_inv_metric(P);
  return stan::io::array_var_context(std::vector<std::string>{"inv_metric"}, g_init_inv_metric,
                                     std::vector<std::vector<size_t>>{{static_cast<size_t>(P)}});
}

// Full NUTS through the reference entry point.  draws: [chain][warmup+sample][7 + P] doubles
// (lp__, accept_stat__, stepsize__, treedepth__, n_leapfrog__, divergent__, energy__, params...).
// warm_leapfrogs[chain] receives sum(n_leapfrog__) over warm-up (save_warmup is forced on
// internally so the column can be read; warm-up rows are not returned).
int ref_glm_nuts(void* h, int num_chains, unsigned seed, unsigned init_chain_id,
                 double init_radius, int num_warmup, int num_samples,
                 double stepsize, int max_depth, double delta, int num_threads,
                 double* draws, double* stepsize_out, double* inv_metric_out,
                 double* warm_leapfrogs, double* wall_seconds, char* err,
                 int errlen) {
  auto& m = *static_cast<ref_glm_model*>(h);
  const int P = static_cast<int>(m.num_params_r());
  int rc = 0;
  int g = guarded(err, errlen, [&] {
    stan::math::init_threadpool_tbb(num_threads > 0 ? num_threads : num_chains);
    std::vector<std::shared_ptr<stan::io::var_context>> inits, metrics;
    for (int c = 0; c < num_chains; ++c) {
      inits.emplace_back(std::make_shared<stan::io::empty_var_context>());
      metrics.emplace_back(std::make_shared<stan::io::array_var_context>(init_metric_context(P)));
    }
    stan::callbacks::interrupt interrupt;
    err_logger logger;
    std::vector<stan::callbacks::writer> init_w(num_chains), diag_w(num_chains);
    std::vector<mem_writer> sample_w(num_chains);
    std::vector<mem_metric_writer> metric_w(num_chains);
    auto t0 = std::chrono::steady_clock::now();
    rc = stan::services::sample::hmc_nuts_diag_e_adapt(
        m, num_chains, inits, metrics, seed, init_chain_id, init_radius,
        num_warmup, num_samples, 1, true, 0, stepsize, g_stepsize_jitter, max_depth, delta,
        g_gamma, g_kappa, g_t0, g_init_buffer, g_term_buffer, g_window, interrupt, logger, init_w, sample_w,
        diag_w, metric_w);
    auto t1 = std::chrono::steady_clock::now();
    if (wall_seconds)
      *wall_seconds = std::chrono::duration<double>(t1 - t0).count();
    if (rc != 0)
      throw std::runtime_error("hmc_nuts_diag_e_adapt rc=" + std::to_string(rc)
                               + ": " + logger.errors);
    const int W = 7 + P;
    for (int c = 0; c < num_chains; ++c) {
      auto& rows = sample_w[c].rows;
      if (static_cast<int>(rows.size()) != num_warmup + num_samples)
        throw std::runtime_error("unexpected number of draws");
      double wl = 0;
      for (int i = 0; i < num_warmup; ++i)
        wl += rows[i][4];
      if (warm_leapfrogs)
        warm_leapfrogs[c] = wl;
      const int T = num_warmup + num_samples;
      for (int i = 0; i < T; ++i) {
        auto& r = rows[i];
        if (static_cast<int>(r.size()) != W)
          throw std::runtime_error("unexpected draw width");
        std::memcpy(draws + (static_cast<size_t>(c) * T + i) * W, r.data(),
                    W * sizeof(double));
      }
      if (stepsize_out)
        stepsize_out[c] = metric_w[c].stepsize;
      if (inv_metric_out)
        std::memcpy(inv_metric_out + static_cast<size_t>(c) * P,
                    metric_w[c].inv_metric.data(), P * sizeof(double));
    }
  });
  return g ? g : rc;
}

// The product's device-side NUTS (driver + per-chain state machine) on the host backend; same outputs as ref_glm_nuts.
// stats: 4 longs {rounds, lanes, uniform variates generated, normal vectors generated}.
int ref_glm_nuts_device_host(void* h, int num_chains, unsigned seed, unsigned init_chain_id, double init_radius,
                             int num_warmup, int num_samples, double stepsize, int max_depth, double delta,
                             double* draws, double* stepsize_out, double* inv_metric_out, long* stats, char* err,
                             int errlen) {
  auto& m = *static_cast<ref_glm_model*>(h);
  const int P = static_cast<int>(m.num_params_r());
  int rc = 0;
  int g = guarded(err, errlen, [&] {
    std::vector<std::shared_ptr<stan::io::var_context>> inits, metrics;
    for (int c = 0; c < num_chains; ++c) {
      inits.emplace_back(std::make_shared<stan::io::empty_var_context>());
      metrics.emplace_back(std::make_shared<stan::io::array_var_context>(init_metric_context(P)));
    }
    stan::callbacks::interrupt interrupt;
    err_logger logger;
    std::vector<stan::callbacks::writer> init_w(num_chains), diag_w(num_chains);
    std::vector<mem_writer> sample_w(num_chains);
    std::vector<mem_metric_writer> metric_w(num_chains);
    oracle_ref::nuts_host_backend<ref_glm_model> backend(m);
    b200::nuts_backend be = backend.table();
    rc = b200::hmc_nuts_diag_e_adapt_device(m, be, num_chains, inits, metrics, seed, init_chain_id, init_radius,
                                            num_warmup, num_samples, 1, true, 0, stepsize, g_stepsize_jitter, max_depth, delta, g_gamma,
                                            g_kappa, g_t0, g_init_buffer, g_term_buffer, g_window, interrupt, logger, init_w, sample_w, diag_w,
                                            metric_w, stats);
    if (rc != 0)
      throw std::runtime_error("hmc_nuts_diag_e_adapt_device rc=" + std::to_string(rc) + ": " + logger.errors);
    const int W = 7 + P, T = num_warmup + num_samples;
    for (int c = 0; c < num_chains; ++c) {
      auto& rows = sample_w[c].rows;
      if (static_cast<int>(rows.size()) != T)
        throw std::runtime_error("unexpected number of draws");
      for (int i = 0; i < T; ++i) {
        if (static_cast<int>(rows[i].size()) != W)
          throw std::runtime_error("unexpected draw width");
        std::memcpy(draws + (static_cast<size_t>(c) * T + i) * W, rows[i].data(), W * sizeof(double));
      }
      if (stepsize_out)
        stepsize_out[c] = metric_w[c].stepsize;
      if (inv_metric_out)
        std::memcpy(inv_metric_out + static_cast<size_t>(c) * P, metric_w[c].inv_metric.data(), P * sizeof(double));
    }
  });
  return g ? g : rc;
}

// Transcript of a run: everything the sample writer, the diagnostic writer and the logger receive, as text, for
//   which = 0: stan::services::sample::hmc_nuts_diag_e_adapt (the reference), which = 1: the product's device-NUTS driver
// on the host backend -- with the service options the other entry points fix (num_thin, save_warmup, refresh).
// Lines: "S<chain>|..." sample writer, "D<chain>|..." diagnostic writer, "L|..." logger info.  Numbers with 17 digits.
namespace {
struct text_writer : public stan::callbacks::writer {
  std::string tag;
  std::string* out = nullptr;
  std::mutex* mu = nullptr;
  void put(const std::string& line) {
    std::lock_guard<std::mutex> g(*mu);
    *out += tag + "|" + line + "\n";
  }
  void operator()(const std::vector<std::string>& n) override {
    std::string l;
    for (size_t i = 0; i < n.size(); ++i)
      l += (i ? "," : "") + n[i];
    put(l);
  }
  void operator()(const std::vector<double>& v) override {
    std::stringstream ss;
    ss.precision(17);
    for (size_t i = 0; i < v.size(); ++i)
      ss << (i ? "," : "") << v[i];
    put(ss.str());
  }
  void operator()() override { put(""); }
  void operator()(const std::string& m) override { put("#" + m); }
};
struct text_logger : public stan::callbacks::logger {
  std::string* out = nullptr;
  std::mutex* mu = nullptr;
  std::string errors;
  void info(const std::string& m) override {
    std::lock_guard<std::mutex> g(*mu);
    *out += "L|" + m + "\n";
  }
  void info(const std::stringstream& m) override { info(m.str()); }
  void error(const std::string& m) override {
    std::lock_guard<std::mutex> g(*mu);
    errors += m + "\n";
  }
  void error(const std::stringstream& m) override { error(m.str()); }
};
}  // namespace

int ref_glm_nuts_transcript(void* h, int which, int num_chains, unsigned seed, unsigned init_chain_id, int num_warmup,
                            int num_samples, int num_thin, int save_warmup, int refresh, double stepsize, int max_depth,
                            char* out, long out_len, char* err, int errlen) {
  auto& m = *static_cast<ref_glm_model*>(h);
  const int P = static_cast<int>(m.num_params_r());
  int rc = 0;
  int g = guarded(err, errlen, [&] {
    std::vector<std::shared_ptr<stan::io::var_context>> inits, metrics;
    for (int c = 0; c < num_chains; ++c) {
      inits.emplace_back(std::make_shared<stan::io::empty_var_context>());
      metrics.emplace_back(std::make_shared<stan::io::array_var_context>(
          stan::services::util::create_unit_e_diag_inv_metric(P)));
    }
    stan::callbacks::interrupt interrupt;
    std::string text;
    std::mutex mu;
    text_logger logger;
    logger.out = &text;
    logger.mu = &mu;
    std::vector<stan::callbacks::writer> init_w(num_chains);
    std::vector<text_writer> sample_w(num_chains), diag_w(num_chains);
    for (int c = 0; c < num_chains; ++c) {
      sample_w[c].tag = "S" + std::to_string(c);
      diag_w[c].tag = "D" + std::to_string(c);
      sample_w[c].out = diag_w[c].out = &text;
      sample_w[c].mu = diag_w[c].mu = &mu;
    }
    std::vector<mem_metric_writer> metric_w(num_chains);
    if (which == 0) {
      stan::math::init_threadpool_tbb(1);
      rc = stan::services::sample::hmc_nuts_diag_e_adapt(   // the multi-chain overload (:331-404)
          m, num_chains, inits, metrics, seed, init_chain_id, 2.0, num_warmup, num_samples, num_thin, save_warmup != 0,
          refresh, stepsize, 0.0, max_depth, 0.8, 0.05, 0.75, 10.0, 75, 50, 25, interrupt, logger, init_w, sample_w,
          diag_w, metric_w);
    } else {
      oracle_ref::nuts_host_backend<ref_glm_model> backend(m);
      b200::nuts_backend be = backend.table();
      rc = b200::hmc_nuts_diag_e_adapt_device(m, be, num_chains, inits, metrics, seed, init_chain_id, 2.0, num_warmup,
                                              num_samples, num_thin, save_warmup != 0, refresh, stepsize, 0.0, max_depth,
                                              0.8, 0.05, 0.75, 10.0, 75, 50, 25, interrupt, logger, init_w, sample_w,
                                              diag_w, metric_w);
    }
    if (rc != 0)
      throw std::runtime_error("sampler rc=" + std::to_string(rc) + ": " + logger.errors);
    if (static_cast<long>(text.size()) + 1 > out_len)
      throw std::runtime_error("transcript buffer too small");
    std::memcpy(out, text.c_str(), text.size() + 1);
  });
  return g ? g : rc;
}

// draws: column-major n_draws x n_chains (one parameter)
// The bare reference GLM densities (function level), for the parity test of b200glm_glm_lpmf:
//   stan::math::{bernoulli_logit,poisson_log,normal_id,binomial_logit,neg_binomial_2_log}_glm_lp*f<propto>(
//       y, [trials,] X, alpha | a[group], beta [, sigma | phi])
// with alpha/beta (and sigma iff sigma_is_var) as reverse-mode vars when operands_are_var, else doubles.
int ref_glm_function(int family, int propto, int operands_are_var, int sigma_is_var, long long N, int K,
                     const double* X, const int* y_int, const double* y_real, const int* group, int G,
                     const double* alpha, const double* beta, double sigma, double* logp, double* d_alpha,
                     double* d_beta, double* d_sigma, char* err, int errlen, const int* trials) {
  return guarded(err, errlen, [&] {
    using stan::math::var;
    std::vector<int> nt(trials ? trials : nullptr, trials ? trials + N : nullptr);
    Eigen::Map<const Eigen::MatrixXd> Xm(X, N, K);
    Eigen::MatrixXd Xc = Xm;
    std::vector<int> yi(y_int ? y_int : nullptr, y_int ? y_int + N : nullptr);
    Eigen::VectorXd yr = y_real ? Eigen::VectorXd(Eigen::Map<const Eigen::VectorXd>(y_real, N)) : Eigen::VectorXd();
    std::vector<int> grp(group ? group : nullptr, group ? group + N : nullptr);
    const int nA = G > 0 ? G : 1;
    auto run = [&](auto tag_alpha, auto tag_sigma) {
      using TA = decltype(tag_alpha);
      using TS = decltype(tag_sigma);
      stan::math::nested_rev_autodiff nested;
      Eigen::Matrix<TA, -1, 1> a(nA), b(K);
      for (int g = 0; g < nA; ++g) a[g] = alpha[g];
      for (int k = 0; k < K; ++k) b[k] = beta[k];
      TS sg = sigma;
      auto call = [&](const auto& intercept) -> stan::return_type_t<TA, TS> {
        if (family == 0)
          return propto ? stan::math::bernoulli_logit_glm_lpmf<true>(yi, Xc, intercept, b)
                        : stan::math::bernoulli_logit_glm_lpmf<false>(yi, Xc, intercept, b);
        if (family == 1)
          return propto ? stan::math::poisson_log_glm_lpmf<true>(yi, Xc, intercept, b)
                        : stan::math::poisson_log_glm_lpmf<false>(yi, Xc, intercept, b);
        if (family == 3)
          return propto ? stan::math::binomial_logit_glm_lpmf<true>(yi, nt, Xc, intercept, b)
                        : stan::math::binomial_logit_glm_lpmf<false>(yi, nt, Xc, intercept, b);
        if (family == 4)
          return propto ? stan::math::neg_binomial_2_log_glm_lpmf<true>(yi, Xc, intercept, b, sg)
                        : stan::math::neg_binomial_2_log_glm_lpmf<false>(yi, Xc, intercept, b, sg);
        return propto ? stan::math::normal_id_glm_lpdf<true>(yr, Xc, intercept, b, sg)
                      : stan::math::normal_id_glm_lpdf<false>(yr, Xc, intercept, b, sg);
      };
      stan::return_type_t<TA, TS> lp = 0;
      if (G > 0) {
        Eigen::Matrix<TA, -1, 1> an(N);
        for (long long i = 0; i < N; ++i) an[i] = a[grp[i] - 1];
        lp = call(an);
      } else {
        TA a0 = a[0];
        lp = call(a0);
        a[0] = a0;
      }
      *logp = stan::math::value_of(lp);
      if constexpr (std::is_same<TA, var>::value || std::is_same<TS, var>::value) {
        lp.grad();
        if constexpr (std::is_same<TA, var>::value) {
          for (int g = 0; g < nA; ++g) d_alpha[g] = a[g].adj();
          for (int k = 0; k < K; ++k) d_beta[k] = b[k].adj();
        }
        if constexpr (std::is_same<TS, var>::value) *d_sigma = sg.adj();
      }
    };
    for (int g = 0; g < nA; ++g) d_alpha[g] = 0;
    for (int k = 0; k < K; ++k) d_beta[k] = 0;
    *d_sigma = 0;
    if (operands_are_var) {
      if (sigma_is_var && (family == 2 || family == 4))
        run(var(0), var(0));
      else
        run(var(0), double(0));
    } else if (sigma_is_var && (family == 2 || family == 4)) {
      run(double(0), var(0));     // the scale / precision is the only autodiff operand
    } else {
      run(double(0), double(0));
    }
  });
}

double ref_ess(const double* d, int n_draws, int n_chains) {
  Eigen::MatrixXd m = Eigen::Map<const Eigen::MatrixXd>(d, n_draws, n_chains);
  return stan::analyze::ess(m);
}
double ref_rhat(const double* d, int n_draws, int n_chains) {
  Eigen::MatrixXd m = Eigen::Map<const Eigen::MatrixXd>(d, n_draws, n_chains);
  return stan::analyze::rhat(m);
}
double ref_mcse_mean(const double* d, int n_draws, int n_chains) {
  Eigen::MatrixXd m = Eigen::Map<const Eigen::MatrixXd>(d, n_draws, n_chains);
  return stan::analyze::mcse_mean(m);
}
double ref_mcse_sd(const double* d, int n_draws, int n_chains) {
  Eigen::MatrixXd m = Eigen::Map<const Eigen::MatrixXd>(d, n_draws, n_chains);
  return stan::analyze::mcse_sd(m);
}
void ref_split_rank_normalized_ess(const double* d, int n_draws, int n_chains,
                                   double* bulk, double* tail) {
  Eigen::MatrixXd m = Eigen::Map<const Eigen::MatrixXd>(d, n_draws, n_chains);
  auto r = stan::analyze::split_rank_normalized_ess(m);
  *bulk = r.first;
  *tail = r.second;
}

const char* ref_oracle_version() {
  return "stan-dev/stan@9048555 + stan-dev/math@2fdd3ed, compiled from /root/reference";
}

}  // extern "C"
