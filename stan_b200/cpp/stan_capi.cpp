// stan_capi.cpp -- C entry points over the UNMODIFIED reference services driving the B200 backend.
//
// Built only where the reference headers exist (the build container); the resulting
// stan_b200/lib/libb200stan.so travels to the GPU box.  It is the executable proof of the drop-in:
//   b200stan_nuts          -> stan::services::sample::hmc_nuts_diag_e_adapt  (ST/services/sample/hmc_nuts_diag_e_adapt.hpp:58,331)
//                             with Model = b200::glm_model; base_nuts, diag_e_metric, adaptation, writers untouched
//   b200stan_nuts_batched  -> b200::hmc_nuts_diag_e_adapt_batched (b200/batched_nuts.hpp): the same single-chain service per
//                             chain, all chains served by one batched DMMA launch per leapfrog step
//   b200stan_nuts_device   -> b200::hmc_nuts_diag_e_adapt_device (b200/device_nuts.hpp): transition + adaptation in the per-chain
//                             state machine on the device (b200glm_nuts_*), the host keeps the chains' boost engines
//   b200stan_func_eval     -> stan::math::{bernoulli_logit,poisson_log,normal_id}_glm_lp*f overloads on b200::glm_data views
//                             (b200/glm_functions.hpp): the function-level slot the OpenCL backend uses
//   b200stan_log_prob_grad -> stan::model::log_prob_grad<propto,jacobian>    (tape path via precomputed_gradients)
//   b200stan_gradient      -> stan::model::gradient                          (explicit specialisation, no tape)
//   b200stan_leapfrog      -> stan::mcmc::expl_leapfrog<diag_e_metric<...>>::evolve (device-resident specialisation)
#include <b200/stan_glm_model.hpp>
#include <b200/batched_nuts.hpp>
#include <b200/device_nuts.hpp>
#include <b200/glm_functions.hpp>

#include <stan/analyze/mcmc/ess.hpp>
#include <stan/analyze/mcmc/mcse.hpp>
#include <stan/analyze/mcmc/rhat.hpp>
#include <stan/callbacks/interrupt.hpp>
#include <stan/callbacks/logger.hpp>
#include <stan/callbacks/structured_writer.hpp>
#include <stan/callbacks/writer.hpp>
#include <stan/io/empty_var_context.hpp>
#include <stan/io/dump.hpp>
#include <stan/io/json/json_data.hpp>
#include <stan/model/log_prob_propto.hpp>
#include <stan/io/stan_csv_reader.hpp>
#include <stan/callbacks/json_writer.hpp>
#include <stan/callbacks/unique_stream_writer.hpp>
#include <stan/model/log_prob_grad.hpp>
#include <stan/services/sample/hmc_nuts_diag_e_adapt.hpp>
#include <stan/services/util/create_unit_e_diag_inv_metric.hpp>

#include <chrono>
#include <cstring>
#include <fstream>
#include <memory>
#include <mutex>
#include <sstream>

namespace {

using b200::glm_model;

void set_err(char* buf, int len, const char* msg) {
  if (buf && len > 0) {
    std::strncpy(buf, msg, len - 1);
    buf[len - 1] = 0;
  }
}

template <typename F>
int guarded(char* err, int errlen, F&& f) {
  static thread_local stan::math::ChainableStack thread_tape;  // STAN_THREADS: one AD tape per thread
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

struct draw_writer : public stan::callbacks::writer {
  std::vector<std::string> names;
  std::vector<std::vector<double>> rows;
  void operator()(const std::vector<std::string>& n) override { names = n; }
  void operator()(const std::vector<double>& s) override { rows.push_back(s); }
  void operator()() override {}
  void operator()(const std::string&) override {}
};

struct metric_writer : public stan::callbacks::structured_writer {
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

struct collecting_logger : public stan::callbacks::logger {
  std::mutex m;
  std::string errors;
  long n_reject = 0;
  void error(const std::string& s) override {
    std::lock_guard<std::mutex> g(m);
    if (s.find("about to be rejected") != std::string::npos)
      ++n_reject;
    if (errors.size() < 4000)
      errors += s + "\n";
  }
  void error(const std::stringstream& s) override { error(s.str()); }
  void fatal(const std::string& s) override { error(s); }
  void fatal(const std::stringstream& s) override { error(s.str()); }
};

template <bool propto, bool jacobian>
double lp_grad(const glm_model& m, std::vector<double>& th, std::vector<double>& grad) {
  std::vector<int> pi;
  return stan::model::log_prob_grad<propto, jacobian>(m, th, pi, grad, nullptr);
}

}  // namespace

extern "C" {

void* b200stan_create(const b200glm_desc* d, char* err, int errlen) {
  glm_model* m = nullptr;
  guarded(err, errlen, [&] { m = new glm_model(*d); });
  return m;
}
// The stanc-style constructor: data block read by the reference's own JSON var_context
// (stan::json::json_data, ST/io/json/json_data.hpp:42-73) from a file in CmdStan's data format.
// names: NULL or "" keeps the default ("N", "K", "X", "y", "G", "group").
void* b200stan_create_from_json(const char* path, int family, const char* name_y, const char* name_X, int center_x,
                                double prior_alpha_sd, double prior_beta_sd, double prior_sigma_loc,
                                double prior_sigma_scale, double prior_sigma_a_scale, int device, int n_slots,
                                char* err, int errlen) {
  glm_model* m = nullptr;
  guarded(err, errlen, [&] {
    std::ifstream in(path);
    if (!in.good())
      throw std::invalid_argument(std::string("cannot open ") + path);
    stan::json::json_data context(in);
    b200::glm_config cfg;
    cfg.family = family;
    if (name_y && *name_y)
      cfg.name_y = name_y;
    if (name_X && *name_X)
      cfg.name_X = name_X;
    cfg.center_x = center_x != 0;
    cfg.prior_alpha_sd = prior_alpha_sd;
    cfg.prior_beta_sd = prior_beta_sd;
    cfg.prior_sigma_loc = prior_sigma_loc;
    cfg.prior_sigma_scale = prior_sigma_scale;
    cfg.prior_sigma_a_scale = prior_sigma_a_scale;
    cfg.device = device;
    cfg.n_slots = n_slots;
    m = new glm_model(context, cfg);
  });
  return m;
}
// The same constructor fed by the reference's R-dump var_context (stan::io::dump, ST/io/dump.hpp), the other data
// format SURVEY 8f row 4 names.
void* b200stan_create_from_dump(const char* path, int family, int device, int n_slots, char* err, int errlen) {
  glm_model* m = nullptr;
  guarded(err, errlen, [&] {
    std::ifstream in(path);
    if (!in.good())
      throw std::invalid_argument(std::string("cannot open ") + path);
    stan::io::dump context(in);
    b200::glm_config cfg;
    cfg.family = family;
    cfg.device = device;
    cfg.n_slots = n_slots;
    m = new glm_model(context, cfg);
  });
  return m;
}
// stan::model::log_prob_propto<jacobian> (ST/model/log_prob_propto.hpp:32-52 and the Eigen overload :75-96): the
// call base_hamiltonian::update_potential makes (base_hamiltonian.hpp:54).  which = 0: std::vector signature,
// 1: Eigen signature.
int b200stan_log_prob_propto(void* h, const double* theta, int jacobian, int which, double* lp, char* err,
                             int errlen) {
  return guarded(err, errlen, [&] {
    const glm_model& m = *static_cast<glm_model*>(h);
    const size_t P = m.num_params_r();
    if (which == 0) {
      std::vector<double> th(theta, theta + P);
      std::vector<int> pi;
      *lp = jacobian ? stan::model::log_prob_propto<true>(m, th, pi, nullptr)
                     : stan::model::log_prob_propto<false>(m, th, pi, nullptr);
    } else {
      Eigen::VectorXd th = Eigen::Map<const Eigen::VectorXd>(theta, P);
      *lp = jacobian ? stan::model::log_prob_propto<true>(m, th, nullptr)
                     : stan::model::log_prob_propto<false>(m, th, nullptr);
    }
  });
}
// column means removed by center_x (K doubles); returns the number written (0 if not centred)
int b200stan_means_x(void* h, double* out) {
  const auto& mu = static_cast<glm_model*>(h)->means_x();
  for (size_t k = 0; k < mu.size(); ++k)
    out[k] = mu[k];
  return static_cast<int>(mu.size());
}
void b200stan_destroy(void* h) { delete static_cast<glm_model*>(h); }
// the b200glm_handle* behind the model (row-sharded runs: b200glm_peer_export / peer_connect / comm_init on it)
void* b200stan_backend_handle(void* h) { return static_cast<glm_model*>(h)->handle(); }
int b200stan_num_params(void* h) { return static_cast<int>(static_cast<glm_model*>(h)->num_params_r()); }
void b200stan_counters(void* h, long* n_gradients, long* n_leapfrogs, long* n_uploads) {
  auto* m = static_cast<glm_model*>(h);
  *n_gradients = m->n_gradients();
  *n_leapfrogs = m->n_leapfrogs();
  *n_uploads = m->n_uploads();
}

int b200stan_log_prob_grad(void* h, const double* theta, int propto, int jacobian, double* lp, double* grad,
                           char* err, int errlen) {
  auto& m = *static_cast<glm_model*>(h);
  const size_t P = m.num_params_r();
  return guarded(err, errlen, [&] {
    std::vector<double> th(theta, theta + P), g;
    if (propto)
      *lp = jacobian ? lp_grad<true, true>(m, th, g) : lp_grad<true, false>(m, th, g);
    else
      *lp = jacobian ? lp_grad<false, true>(m, th, g) : lp_grad<false, false>(m, th, g);
    if (grad)
      std::memcpy(grad, g.data(), P * sizeof(double));
  });
}

int b200stan_log_prob(void* h, const double* theta, int propto, int jacobian, double* lp, char* err, int errlen) {
  auto& m = *static_cast<glm_model*>(h);
  const size_t P = m.num_params_r();
  return guarded(err, errlen, [&] {
    std::vector<double> th(theta, theta + P);
    std::vector<int> pi;
    if (propto)
      *lp = jacobian ? m.template log_prob<true, true>(th, pi, nullptr) : m.template log_prob<true, false>(th, pi, nullptr);
    else
      *lp = jacobian ? m.template log_prob<false, true>(th, pi, nullptr)
                     : m.template log_prob<false, false>(th, pi, nullptr);
  });
}

int b200stan_gradient(void* h, const double* theta, double* lp, double* grad, char* err, int errlen) {
  auto& m = *static_cast<glm_model*>(h);
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

// n_steps consecutive evolve() calls of the (specialised) reference integrator on one diag_e_point,
// after hamiltonian.init(z) as base_nuts.hpp:85 does.
int b200stan_leapfrog(void* h, double eps, const double* inv_metric, int n_steps, double* q, double* p, double* g,
                      double* V, char* err, int errlen) {
  auto& m = *static_cast<glm_model*>(h);
  const int P = static_cast<int>(m.num_params_r());
  return guarded(err, errlen, [&] {
    using H = stan::mcmc::diag_e_metric<glm_model, stan::rng_t>;
    H ham(m);
    stan::mcmc::expl_leapfrog<H> integrator;
    stan::mcmc::diag_e_point z(P);
    stan::callbacks::logger logger;
    z.q = Eigen::Map<const Eigen::VectorXd>(q, P);
    z.p = Eigen::Map<const Eigen::VectorXd>(p, P);
    if (inv_metric)
      z.inv_e_metric_ = Eigen::Map<const Eigen::VectorXd>(inv_metric, P);
    ham.init(z, logger);
    for (int i = 0; i < n_steps; ++i)
      integrator.evolve(z, ham, eps, logger);
    std::memcpy(q, z.q.data(), P * sizeof(double));
    std::memcpy(p, z.p.data(), P * sizeof(double));
    std::memcpy(g, z.g.data(), P * sizeof(double));
    *V = z.V;
  });
}

// stepsize_jitter of the NUTS entry points below (the services' argument; set before the call, 0 by default)
static double g_stepsize_jitter = 0.0;
void b200stan_set_stepsize_jitter(double j) { g_stepsize_jitter = j; }

// draws: [chain][warmup+sample][7 + P] (lp__, accept_stat__, stepsize__, treedepth__, n_leapfrog__, divergent__, energy__, params)
static int nuts_impl(void* h, int mode /* 0 reference service, 1 batched host driver, 2 device-side */, int num_chains, unsigned seed, unsigned init_chain_id, double init_radius,
                     int num_warmup, int num_samples, double stepsize, int max_depth, double delta, int num_threads,
                     double* draws, double* stepsize_out, double* inv_metric_out, double* warm_leapfrogs,
                     double* wall_seconds, long* batch_stats, char* err, int errlen) {
  auto& m = *static_cast<glm_model*>(h);
  const int P = static_cast<int>(m.num_params_r());
  int rc = 0;
  int g = guarded(err, errlen, [&] {
    const bool batched = mode == 1;
    if (mode == 0)
      stan::math::init_threadpool_tbb(num_threads > 0 ? num_threads : num_chains);
    std::vector<std::shared_ptr<stan::io::var_context>> inits, metrics;
    for (int c = 0; c < num_chains; ++c) {
      inits.emplace_back(std::make_shared<stan::io::empty_var_context>());
      metrics.emplace_back(std::make_shared<stan::io::array_var_context>(
          stan::services::util::create_unit_e_diag_inv_metric(P)));
    }
    stan::callbacks::interrupt interrupt;
    collecting_logger logger;
    std::vector<stan::callbacks::writer> init_w(num_chains), diag_w(num_chains);
    std::vector<draw_writer> sample_w(num_chains);
    std::vector<metric_writer> metric_w(num_chains);
    auto t0 = std::chrono::steady_clock::now();
    if (mode == 2) {
      // the C ABI of libb200glm.so as the driver's backend table
      static_assert(sizeof(b200::nuts_config) == sizeof(b200glm_nuts_config), "nuts_config mirrors b200glm_nuts_config");
      static_assert(sizeof(b200::nuts_status) == sizeof(b200glm_nuts_status), "nuts_status mirrors b200glm_nuts_status");
      b200::nuts_backend be;
      be.ctx = m.handle();
      be.reserve = [](void* c, std::int32_t n, const b200::nuts_config* cfg) {
        return b200glm_nuts_reserve(static_cast<b200glm_handle*>(c), n, reinterpret_cast<const b200glm_nuts_config*>(cfg));
      };
      be.buffers = [](void* c, double** no, double** un, b200::nuts_status** st, double** dr, double** me) {
        return b200glm_nuts_buffers(static_cast<b200glm_handle*>(c), no, un, reinterpret_cast<b200glm_nuts_status**>(st), dr, me);
      };
      be.init_chain = [](void* c, std::int32_t chain, const double* q0, const double* im, double eps) {
        return b200glm_nuts_init_chain(static_cast<b200glm_handle*>(c), chain, q0, im, eps);
      };
      be.round = [](void* c, std::int32_t n, const std::int32_t* chains) {
        return b200glm_nuts_round(static_cast<b200glm_handle*>(c), n, chains);
      };
      be.last_error = [](void* c) { return b200glm_last_error(static_cast<b200glm_handle*>(c)); };
      rc = b200::hmc_nuts_diag_e_adapt_device(m, be, num_chains, inits, metrics, seed, init_chain_id, init_radius,
                                              num_warmup, num_samples, 1, true, 0, stepsize, g_stepsize_jitter, max_depth, delta, 0.05,
                                              0.75, 10.0, 75, 50, 25, interrupt, logger, init_w, sample_w, diag_w,
                                              metric_w, batch_stats);
    } else if (batched)
      rc = b200::hmc_nuts_diag_e_adapt_batched(
          m, num_chains, inits, metrics, seed, init_chain_id, init_radius, num_warmup, num_samples, 1, true, 0,
          stepsize, g_stepsize_jitter, max_depth, delta, 0.05, 0.75, 10.0, 75, 50, 25, interrupt, logger, init_w, sample_w, diag_w,
          metric_w, batch_stats);
    else
      rc = stan::services::sample::hmc_nuts_diag_e_adapt(
          m, num_chains, inits, metrics, seed, init_chain_id, init_radius, num_warmup, num_samples, 1, true, 0,
          stepsize, g_stepsize_jitter, max_depth, delta, 0.05, 0.75, 10.0, 75, 50, 25, interrupt, logger, init_w, sample_w, diag_w,
          metric_w);
    auto t1 = std::chrono::steady_clock::now();
    if (wall_seconds)
      *wall_seconds = std::chrono::duration<double>(t1 - t0).count();
    if (rc != 0)
      throw std::runtime_error("hmc_nuts_diag_e_adapt rc=" + std::to_string(rc) + ": " + logger.errors);
    const int W = 7 + P;
    for (int c = 0; c < num_chains; ++c) {
      auto& rows = sample_w[c].rows;
      if (static_cast<int>(rows.size()) != num_warmup + num_samples)
        throw std::runtime_error("unexpected number of draws: " + logger.errors);
      double wl = 0;
      for (int i = 0; i < num_warmup; ++i)
        wl += rows[i][4];
      if (warm_leapfrogs)
        warm_leapfrogs[c] = wl;
      const int T = num_warmup + num_samples;
      for (int i = 0; i < T; ++i)
        std::memcpy(draws + (static_cast<size_t>(c) * T + i) * W, rows[i].data(), W * sizeof(double));
      if (stepsize_out)
        stepsize_out[c] = metric_w[c].stepsize;
      if (inv_metric_out)
        std::memcpy(inv_metric_out + static_cast<size_t>(c) * P, metric_w[c].inv_metric.data(), P * sizeof(double));
    }
  });
  return g ? g : rc;
}

int b200stan_nuts(void* h, int num_chains, unsigned seed, unsigned init_chain_id, double init_radius, int num_warmup,
                  int num_samples, double stepsize, int max_depth, double delta, int num_threads, double* draws,
                  double* stepsize_out, double* inv_metric_out, double* warm_leapfrogs, double* wall_seconds,
                  char* err, int errlen) {
  return nuts_impl(h, 0, num_chains, seed, init_chain_id, init_radius, num_warmup, num_samples, stepsize, max_depth,
                   delta, num_threads, draws, stepsize_out, inv_metric_out, warm_leapfrogs, wall_seconds, nullptr, err,
                   errlen);
}

// batch_stats[10] = {batched launches, lanes served, batches with <= 2, 16, 32, 64, 128, 256, 512, more lanes}
int b200stan_nuts_batched(void* h, int num_chains, unsigned seed, unsigned init_chain_id, double init_radius,
                          int num_warmup, int num_samples, double stepsize, int max_depth, double delta, double* draws,
                          double* stepsize_out, double* inv_metric_out, double* warm_leapfrogs, double* wall_seconds,
                          long* batch_stats, char* err, int errlen) {
  return nuts_impl(h, 1, num_chains, seed, init_chain_id, init_radius, num_warmup, num_samples, stepsize, max_depth,
                   delta, 0, draws, stepsize_out, inv_metric_out, warm_leapfrogs, wall_seconds, batch_stats, err,
                   errlen);
}

// stats[4] = {rounds, leapfrog lanes served, uniform variates generated, vectors of normal variates generated}
int b200stan_nuts_device(void* h, int num_chains, unsigned seed, unsigned init_chain_id, double init_radius,
                         int num_warmup, int num_samples, double stepsize, int max_depth, double delta, double* draws,
                         double* stepsize_out, double* inv_metric_out, double* warm_leapfrogs, double* wall_seconds,
                         long* stats, char* err, int errlen) {
  return nuts_impl(h, 2, num_chains, seed, init_chain_id, init_radius, num_warmup, num_samples, stepsize, max_depth,
                   delta, 0, draws, stepsize_out, inv_metric_out, warm_leapfrogs, wall_seconds, stats, err, errlen);
}

// ---- output formats: the reference's own writers and reader -------------------------
// b200stan_nuts_csv runs the same service as b200stan_nuts but hands it the writers CmdStan uses:
//   sample  -> stan::callbacks::unique_stream_writer<std::ofstream>  (ST/callbacks/unique_stream_writer.hpp:22-36)
//              "<prefix>_<chain>.csv": header, draws, "# Adaptation terminated / Step size / Diagonal elements
//              of inverse mass matrix" block (mcmc_writer.hpp, run_adaptive_sampler.hpp:87-90), timing comments
//   metric  -> stan::callbacks::json_writer<std::ofstream>            (ST/callbacks/json_writer.hpp:170-188)
//              "<prefix>_metric_<chain>.json": {"stepsize": .., "inv_metric": [..]}
// Warm-up draws are not saved: stan_csv_reader skips them only with CmdStan's "# save_warmup" metadata, which is
// CmdStan's preamble, not Stan's.  Values are written with 17 significant digits so that a round trip is exact.
int b200stan_nuts_csv(void* h, int num_chains, unsigned seed, unsigned init_chain_id, double init_radius,
                      int num_warmup, int num_samples, double stepsize, int max_depth, double delta, int num_threads,
                      const char* prefix, char* err, int errlen) {
  auto& m = *static_cast<glm_model*>(h);
  const int P = static_cast<int>(m.num_params_r());
  int rc = 0;
  int g = guarded(err, errlen, [&] {
    stan::math::init_threadpool_tbb(num_threads > 0 ? num_threads : num_chains);
    std::vector<std::shared_ptr<stan::io::var_context>> inits, metrics;
    using csv_writer = stan::callbacks::unique_stream_writer<std::ofstream>;
    using json_writer = stan::callbacks::json_writer<std::ofstream>;
    std::vector<csv_writer> sample_w;
    std::vector<json_writer> metric_w;
    for (int c = 0; c < num_chains; ++c) {
      inits.emplace_back(std::make_shared<stan::io::empty_var_context>());
      metrics.emplace_back(std::make_shared<stan::io::array_var_context>(
          stan::services::util::create_unit_e_diag_inv_metric(P)));
      const std::string id = std::to_string(init_chain_id + c);
      auto os = std::make_unique<std::ofstream>(std::string(prefix) + "_" + id + ".csv");
      auto js = std::make_unique<std::ofstream>(std::string(prefix) + "_metric_" + id + ".json");
      if (!os->good() || !js->good())
        throw std::invalid_argument(std::string("cannot write ") + prefix + "_*.csv");
      os->precision(17);
      js->precision(17);
      sample_w.emplace_back(std::move(os), "# ");
      metric_w.emplace_back(std::move(js));
    }
    stan::callbacks::interrupt interrupt;
    collecting_logger logger;
    std::vector<stan::callbacks::writer> init_w(num_chains), diag_w(num_chains);
    rc = stan::services::sample::hmc_nuts_diag_e_adapt(
        m, num_chains, inits, metrics, seed, init_chain_id, init_radius, num_warmup, num_samples, 1, false, 0,
        stepsize, 0.0, max_depth, delta, 0.05, 0.75, 10.0, 75, 50, 25, interrupt, logger, init_w, sample_w, diag_w,
        metric_w);
    if (rc != 0)
      throw std::runtime_error("hmc_nuts_diag_e_adapt rc=" + std::to_string(rc) + ": " + logger.errors);
  });
  return g ? g : rc;
}

// stan::io::stan_csv_reader::parse (ST/io/stan_csv_reader.hpp:350-383) on a file written above.
// samples: [max_rows][max_cols] row-major (may be NULL to query the shape); header: comma-joined names.
int b200stan_read_csv(const char* path, double* samples, int max_rows, int max_cols, int* n_rows, int* n_cols,
                      double* step_size, double* metric_diag, int max_metric, int* n_metric, char* header,
                      int header_len, char* err, int errlen) {
  return guarded(err, errlen, [&] {
    std::ifstream in(path);
    if (!in.good())
      throw std::invalid_argument(std::string("cannot open ") + path);
    std::stringstream msgs;
    stan::io::stan_csv csv = stan::io::stan_csv_reader::parse(in, &msgs);
    *n_rows = static_cast<int>(csv.samples.rows());
    *n_cols = static_cast<int>(csv.samples.cols());
    *step_size = csv.adaptation.step_size;
    *n_metric = static_cast<int>(csv.adaptation.metric.size());
    for (int i = 0; i < *n_metric && i < max_metric; ++i)
      metric_diag[i] = csv.adaptation.metric(i);
    std::string hd;
    for (size_t i = 0; i < csv.header.size(); ++i)
      hd += (i ? "," : "") + csv.header[i];
    set_err(header, header_len, hd.c_str());
    if (samples)
      for (int r = 0; r < *n_rows && r < max_rows; ++r)
        for (int c = 0; c < *n_cols && c < max_cols; ++c)
          samples[static_cast<size_t>(r) * max_cols + c] = csv.samples(r, c);
  });
}

// ---- function-level binding (b200/glm_functions.hpp) -------------------------------------------------
void* b200stan_func_create(int family, long long N, int K, const double* X, const void* y, const int* group, int G,
                           char* err, int errlen, const int* trials) {
  b200::glm_data* d = nullptr;
  guarded(err, errlen, [&] { d = new b200::glm_data(family, N, K, X, y, group, G, 0, 4, trials); });
  return d;
}
void b200stan_func_destroy(void* d) { delete static_cast<b200::glm_data*>(d); }

// Calls the stan::math overload with alpha/beta (and sigma iff sigma_is_var) as vars when
// operands_are_var, runs the reverse sweep on f = scale * glm(...) + 0.5 * sum(beta^2) (so the op is
// exercised as ONE node of a larger tape) and returns value and adjoints.
int b200stan_func_eval(void* dv, int propto, int operands_are_var, int sigma_is_var, const double* alpha, int n_alpha,
                       const double* beta, double sigma, double scale, double* f, double* d_alpha, double* d_beta,
                       double* d_sigma, char* err, int errlen) {
  const b200::glm_data& d = *static_cast<b200::glm_data*>(dv);
  const int K = d.K(), G = d.G();
  return guarded(err, errlen, [&] {
    using stan::math::var;
    auto run = [&](auto tag_op, auto tag_sigma) {
      using TO = decltype(tag_op);
      using TS = decltype(tag_sigma);
      stan::math::nested_rev_autodiff nested;
      Eigen::Matrix<TO, -1, 1> a(n_alpha), b(K);
      for (int g = 0; g < n_alpha; ++g) a[g] = alpha[g];
      for (int k = 0; k < K; ++k) b[k] = beta[k];
      TS sg = sigma;
      TO a0 = a[0];
      stan::return_type_t<TO, TS> lp = 0;
      auto call = [&](auto pt) {
        constexpr bool P = decltype(pt)::value;
        if (d.family() == B200GLM_BERNOULLI_LOGIT)
          lp = G > 0 ? stan::math::bernoulli_logit_glm_lpmf<P>(d.y(), d.x(), b200::by_group(a), b)
                     : stan::math::bernoulli_logit_glm_lpmf<P>(d.y(), d.x(), a0, b);
        else if (d.family() == B200GLM_POISSON_LOG)
          lp = G > 0 ? stan::math::poisson_log_glm_lpmf<P>(d.y(), d.x(), b200::by_group(a), b)
                     : stan::math::poisson_log_glm_lpmf<P>(d.y(), d.x(), a0, b);
        else if (d.family() == B200GLM_BINOMIAL_LOGIT)
          lp = G > 0 ? stan::math::binomial_logit_glm_lpmf<P>(d.y(), d.trials(), d.x(), b200::by_group(a), b)
                     : stan::math::binomial_logit_glm_lpmf<P>(d.y(), d.trials(), d.x(), a0, b);
        else if (d.family() == B200GLM_NEG_BINOMIAL_2_LOG)
          lp = G > 0 ? stan::math::neg_binomial_2_log_glm_lpmf<P>(d.y(), d.x(), b200::by_group(a), b, sg)
                     : stan::math::neg_binomial_2_log_glm_lpmf<P>(d.y(), d.x(), a0, b, sg);
        else
          lp = G > 0 ? stan::math::normal_id_glm_lpdf<P>(d.y(), d.x(), b200::by_group(a), b, sg)
                     : stan::math::normal_id_glm_lpdf<P>(d.y(), d.x(), a0, b, sg);
      };
      if (propto)
        call(std::true_type());
      else
        call(std::false_type());
      stan::return_type_t<TO, TS> total = scale * lp + 0.5 * stan::math::dot_self(b);
      *f = stan::math::value_of(total);
      for (int g = 0; g < n_alpha; ++g) d_alpha[g] = 0;
      for (int k = 0; k < K; ++k) d_beta[k] = 0;
      *d_sigma = 0;
      if constexpr (std::is_same<TO, var>::value || std::is_same<TS, var>::value) {
        total.grad();
        if constexpr (std::is_same<TO, var>::value) {
          if (G > 0)
            for (int g = 0; g < n_alpha; ++g) d_alpha[g] = a[g].adj();
          else
            d_alpha[0] = a0.adj();
          for (int k = 0; k < K; ++k) d_beta[k] = b[k].adj();
        }
        if constexpr (std::is_same<TS, var>::value) *d_sigma = sg.adj();
      }
    };
    const bool has_scale = d.family() == B200GLM_NORMAL_ID || d.family() == B200GLM_NEG_BINOMIAL_2_LOG;
    if (operands_are_var) {
      if (sigma_is_var && has_scale)
        run(var(0), var(0));
      else
        run(var(0), double(0));
    } else if (sigma_is_var && has_scale) {
      run(double(0), var(0));     // the scale / precision is the only autodiff operand
    } else {
      run(double(0), double(0));
    }
  });
}

// The reference's own diagnostics (ST/analyze/mcmc/{ess,mcse,rhat}.hpp) on draws produced through this library:
// d is column-major n_draws x n_chains for one parameter.  which: 0 ess, 1 rhat, 2 mcse_mean, 3 mcse_sd
double b200stan_diagnostic(int which, const double* d, int n_draws, int n_chains) {
  Eigen::MatrixXd m = Eigen::Map<const Eigen::MatrixXd>(d, n_draws, n_chains);
  switch (which) {
    case 0: return stan::analyze::ess(m);
    case 1: return stan::analyze::rhat(m);
    case 2: return stan::analyze::mcse_mean(m);
    case 3: return stan::analyze::mcse_sd(m);
  }
  return std::numeric_limits<double>::quiet_NaN();
}

// The per-row overloads (b200::by_row(alpha), vector sigma) as one node of a reverse-mode tape:
// f = scale * glm(y | x, by_row(a), b [, s]) + 0.5 * sum(b^2), all operands vars; sigma_rows may be NULL (scalar sigma).
int b200stan_func_eval_rows(void* dv, int propto, const double* alpha_rows, const double* beta,
                            const double* sigma_rows, double sigma, double scale, double* f, double* d_alpha_rows,
                            double* d_beta, double* d_sigma_rows, double* d_sigma, char* err, int errlen) {
  const b200::glm_data& d = *static_cast<b200::glm_data*>(dv);
  const int K = d.K();
  const long long N = d.N();
  return guarded(err, errlen, [&] {
    using stan::math::var;
    stan::math::nested_rev_autodiff nested;
    Eigen::Matrix<var, -1, 1> a(N), b(K), sv(sigma_rows ? N : 0);
    for (long long i = 0; i < N; ++i) a[i] = alpha_rows[i];
    for (int k = 0; k < K; ++k) b[k] = beta[k];
    for (long long i = 0; i < sv.size(); ++i) sv[i] = sigma_rows[i];
    var sg = sigma;
    var lp = 0;
    auto call = [&](auto pt) {
      constexpr bool P = decltype(pt)::value;
      switch (d.family()) {
        case B200GLM_BERNOULLI_LOGIT: lp = stan::math::bernoulli_logit_glm_lpmf<P>(d.y(), d.x(), b200::by_row(a), b); break;
        case B200GLM_POISSON_LOG: lp = stan::math::poisson_log_glm_lpmf<P>(d.y(), d.x(), b200::by_row(a), b); break;
        case B200GLM_BINOMIAL_LOGIT:
          lp = stan::math::binomial_logit_glm_lpmf<P>(d.y(), d.trials(), d.x(), b200::by_row(a), b);
          break;
        case B200GLM_NEG_BINOMIAL_2_LOG:
          lp = stan::math::neg_binomial_2_log_glm_lpmf<P>(d.y(), d.x(), b200::by_row(a), b, sg);
          break;
        default:
          lp = sigma_rows ? stan::math::normal_id_glm_lpdf<P>(d.y(), d.x(), b200::by_row(a), b, sv)
                          : stan::math::normal_id_glm_lpdf<P>(d.y(), d.x(), b200::by_row(a), b, sg);
      }
    };
    if (propto)
      call(std::true_type());
    else
      call(std::false_type());
    var total = scale * lp + 0.5 * stan::math::dot_self(b);
    total.grad();
    *f = total.val();
    for (long long i = 0; i < N; ++i) d_alpha_rows[i] = a[i].adj();
    for (int k = 0; k < K; ++k) d_beta[k] = b[k].adj();
    if (sigma_rows)
      for (long long i = 0; i < N; ++i) d_sigma_rows[i] = sv[i].adj();
    *d_sigma = sg.adj();
  });
}

const char* b200stan_version() {
  return "b200::glm_model behind stan::services::sample::hmc_nuts_diag_e_adapt (reference headers: stan@9048555, math@2fdd3ed)";
}

}  // extern "C"
