#ifndef B200_DEVICE_NUTS_HPP
#define B200_DEVICE_NUTS_HPP
// hmc_nuts_diag_e_adapt with the transition and the adaptation ON THE DEVICE.
//
// Same argument list and outputs as the reference's multi-chain service
// (ST/services/sample/hmc_nuts_diag_e_adapt.hpp:331-404); per chain it does what the single-chain service does
// (:58-117 -> util::run_adaptive_sampler -> generate_transitions), except that adapt_diag_e_nuts::transition -- tree
// building, U-turn checks, multinomial sampling, dual averaging, the Welford metric windows, init_stepsize -- runs in
// the backend's per-chain state machine (stan_b200/csrc/nuts_tree.cuh) right behind the batched leapfrog, for all chains
// at once.  What the host keeps is what cannot be moved without changing the results: every chain's boost engine.
// create_rng(seed, chain id), util::initialize (random inits), and then per momentum refresh P normal variates, per
// direction / multinomial decision one uniform variate, produced in the reference's order and handed to the device
// through pinned memory; the device reports how many it consumed and the engine is put back to exactly that point.  A
// chain therefore gives the reference's draws for the same seed (up to the summation order of dot products).
// Per round the host reads 40 bytes of status per chain, and 3 P + 8 doubles for a chain that finished a transition (the
// draw, and the selected state's momentum and gradient for the diagnostic writer); inside a trajectory q, p, g never leave
// the device.
//
// The backend is a table of C functions (in the product: b200glm_nuts_* of libb200glm.so; tests substitute a host build of
// the same state machine to check this driver against the reference without a GPU).
#include <stan/callbacks/interrupt.hpp>
#include <stan/callbacks/logger.hpp>
#include <stan/callbacks/structured_writer.hpp>
#include <stan/callbacks/writer.hpp>
#include <stan/mcmc/base_mcmc.hpp>
#include <stan/mcmc/sample.hpp>
#include <stan/mcmc/windowed_adaptation.hpp>
#include <stan/services/error_codes.hpp>
#include <stan/services/util/create_rng.hpp>
#include <stan/services/util/initialize.hpp>
#include <stan/services/util/inv_metric.hpp>
#include <stan/services/util/mcmc_writer.hpp>
#include <boost/random/normal_distribution.hpp>
#include <boost/random/uniform_01.hpp>
#include <boost/random/variate_generator.hpp>
#include <chrono>
#include <iomanip>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace b200 {

// mirrors b200glm_nuts_config / b200glm_nuts_status of include/b200glm.h (kept free of that header so that the host
// check can build this file without the CUDA library)
struct nuts_config {
  std::int32_t max_depth, num_warmup, num_samples;
  std::uint32_t w_num_warmup, w_init_buffer, w_term_buffer, w_base_window, w_size0, w_next0;
  double max_deltaH, delta, gamma, kappa, t0;
  double stepsize_jitter;
};
struct nuts_status {
  std::int32_t phase, need_normals, iter, fail_code;
  std::int32_t adapt_done, reserved;
  std::uint64_t n_unif;
  double eps_nom;
};
constexpr int kNutsUnifCap = 64;      // NUTS_UNIF_CAP
constexpr int kNutsUnifStride = 72;   // NUTS_UNIF_STRIDE: the ring, then the step-size jitter variate
constexpr int kNutsDrawExtra = 8;     // NUTS_DRAW_EXTRA
constexpr int kNutsPhaseDone = 5, kNutsPhaseFailed = 6;

struct nuts_backend {
  void* ctx;
  int (*reserve)(void* ctx, std::int32_t n_chains, const nuts_config* cfg);
  int (*buffers)(void* ctx, double** normals, double** uniforms, nuts_status** status, double** draws, double** metric);
  int (*init_chain)(void* ctx, std::int32_t chain, const double* q0, const double* inv_metric, double stepsize);
  int (*round)(void* ctx, std::int32_t n_lanes, const std::int32_t* chains);
  const char* (*last_error)(void* ctx);
};

namespace detail {

// the window schedule exactly as the reference computes it (and its log messages)
struct window_probe : stan::mcmc::windowed_adaptation {
  window_probe() : windowed_adaptation("variance") {}
  unsigned nw() const { return num_warmup_; }
  unsigned ib() const { return adapt_init_buffer_; }
  unsigned tb() const { return adapt_term_buffer_; }
  unsigned bw() const { return adapt_base_window_; }
  unsigned size0() const { return adapt_window_size_; }
  unsigned next0() const { return adapt_next_window_; }
};

// what mcmc_writer asks a sampler for (base_nuts::get_sampler_param_names / get_sampler_params, base_hmc::write_sampler_state,
// ps_point::get_param_names / get_params); the values are those of the chain's latest draw
class device_sampler_view : public stan::mcmc::base_mcmc {
 public:
  explicit device_sampler_view(int P) : q(P, 0.0), p(P, 0.0), g(P, 0.0), inv_metric(P, 1.0) {}
  stan::mcmc::sample transition(stan::mcmc::sample& s, stan::callbacks::logger&) override { return s; }
  void get_sampler_param_names(std::vector<std::string>& names) override {
    for (const char* n : {"stepsize__", "treedepth__", "n_leapfrog__", "divergent__", "energy__"})
      names.push_back(n);
  }
  void get_sampler_params(std::vector<double>& values) override {
    for (double v : {stepsize, treedepth, n_leapfrog, divergent, energy})
      values.push_back(v);
  }
  void write_sampler_state(stan::callbacks::writer& writer) override {
    std::stringstream ss;
    ss << "Step size = " << nom_stepsize;
    writer(ss.str());
    writer("Diagonal elements of inverse mass matrix:");
    std::stringstream ms;
    for (size_t i = 0; i < inv_metric.size(); ++i)
      ms << (i ? ", " : "") << inv_metric[i];
    writer(ms.str());
  }
  void get_sampler_diagnostic_names(std::vector<std::string>& model_names, std::vector<std::string>& names) override {
    for (auto& n : model_names)
      names.emplace_back(n);
    for (auto& n : model_names)
      names.emplace_back("p_" + n);
    for (auto& n : model_names)
      names.emplace_back("g_" + n);
  }
  void get_sampler_diagnostics(std::vector<double>& values) override {
    values.insert(values.end(), q.begin(), q.end());
    values.insert(values.end(), p.begin(), p.end());
    values.insert(values.end(), g.begin(), g.end());
  }
  double stepsize = 0, treedepth = 0, n_leapfrog = 0, divergent = 0, energy = 0, nom_stepsize = 0;
  std::vector<double> q, p, g, inv_metric;
};

// base_hmc::write_sampler_state_struct
inline void write_metric_struct(stan::callbacks::structured_writer& mw, double stepsize,
                                const std::vector<double>& inv_metric) {
  mw.begin_record();
  mw.write("stepsize", stepsize);
  mw.write("metric_type", std::string("diag_e"));
  Eigen::VectorXd im = Eigen::Map<const Eigen::VectorXd>(inv_metric.data(), inv_metric.size());
  mw.write("inv_metric", im);
  mw.end_record();
}

}  // namespace detail

// stats (optional, 4 longs): {rounds, leapfrog lanes served, uniform variates generated, normal vectors generated}
template <class Model, typename InitContextPtr, typename InitInvContextPtr, typename InitWriter, typename SampleWriter,
          typename DiagnosticWriter, typename MetricWriter>
int hmc_nuts_diag_e_adapt_device(Model& model, nuts_backend& be, size_t num_chains,
                                 const std::vector<InitContextPtr>& init,
                                 const std::vector<InitInvContextPtr>& init_inv_metric, unsigned int random_seed,
                                 unsigned int init_chain_id, double init_radius, int num_warmup, int num_samples,
                                 int num_thin, bool save_warmup, int refresh, double stepsize, double stepsize_jitter,
                                 int max_depth, double delta, double gamma, double kappa, double t0,
                                 unsigned int init_buffer, unsigned int term_buffer, unsigned int window,
                                 stan::callbacks::interrupt& interrupt, stan::callbacks::logger& logger,
                                 std::vector<InitWriter>& init_writer, std::vector<SampleWriter>& sample_writer,
                                 std::vector<DiagnosticWriter>& diagnostic_writer,
                                 std::vector<MetricWriter>& metric_writer, long* stats = nullptr) {
  namespace su = stan::services::util;
  using stan::services::error_codes;
  const int C = static_cast<int>(num_chains);
  const int P = static_cast<int>(model.num_params_r());
  if (max_depth < 1 || max_depth > 16 || P < 1 || C < 1) {
    logger.error("hmc_nuts_diag_e_adapt_device: needs 1 <= max_depth <= 16, at least one parameter and one chain");
    return error_codes::CONFIG;
  }

  // ---- per chain: engine, initial point, inverse metric (hmc_nuts_diag_e_adapt.hpp:69-84) ----
  struct chain_host {
    stan::rng_t rng, mark;        // mark = the engine right after the last vector of normal variates
    std::uint64_t base = 0, gen = 0;   // uniform variates: consumed when `mark` was taken / generated so far
    int seen_iter = 0;
    std::vector<double> cont;
    chain_host(stan::rng_t r) : rng(r), mark(r) {}
  };
  std::vector<chain_host> hc;
  hc.reserve(C);
  std::vector<Eigen::VectorXd> inv_metric0(C);
  for (int i = 0; i < C; ++i) {
    hc.emplace_back(su::create_rng(random_seed, init_chain_id + i));
    try {
      hc[i].cont = su::initialize(model, *init[i], hc[i].rng, init_radius, true, logger, init_writer[i]);
      inv_metric0[i] = su::read_diag_inv_metric(*init_inv_metric[i], model.num_params_r(), logger);
      su::validate_diag_inv_metric(inv_metric0[i], logger);
    } catch (const std::exception& e) {
      logger.error(e.what());
      return error_codes::CONFIG;
    }
    hc[i].mark = hc[i].rng;
  }

  // ---- sampler configuration (:86-103) ----
  nuts_config cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.max_depth = max_depth;
  cfg.num_warmup = num_warmup;
  cfg.num_samples = num_samples;
  cfg.max_deltaH = 1000;
  // stepsize_adaptation's setters ignore out-of-range values and keep the defaults
  cfg.delta = (delta > 0 && delta < 1) ? delta : 0.5;
  cfg.gamma = gamma > 0 ? gamma : 0.05;
  cfg.kappa = kappa > 0 ? kappa : 0.75;
  cfg.t0 = t0 > 0 ? t0 : 10;
  // base_hmc::set_stepsize_jitter ignores values outside [0, 1] (the constructor's 0 stays)
  cfg.stepsize_jitter = (stepsize_jitter > 0 && stepsize_jitter < 1) ? stepsize_jitter : 0.0;
  for (int i = 0; i < C; ++i) {   // every chain's service logs the window warnings
    detail::window_probe wp;
    wp.set_window_params(num_warmup, init_buffer, term_buffer, window, logger);
    cfg.w_num_warmup = wp.nw();
    cfg.w_init_buffer = wp.ib();
    cfg.w_term_buffer = wp.tb();
    cfg.w_base_window = wp.bw();
    cfg.w_size0 = wp.size0();   // not recomputed here: the reference's rescaled schedule keeps the constructor's cursor
    cfg.w_next0 = wp.next0();
  }
  auto fail = [&](const char* what) {
    const char* e = be.last_error ? be.last_error(be.ctx) : nullptr;
    logger.error(std::string(what) + (e ? std::string(": ") + e : std::string()));
    return error_codes::SOFTWARE;
  };
  if (be.reserve(be.ctx, C, &cfg) != 0)
    return fail("device NUTS: reserve");
  double *normals = nullptr, *uniforms = nullptr, *draws = nullptr, *metric = nullptr;
  nuts_status* status = nullptr;
  if (be.buffers(be.ctx, &normals, &uniforms, &status, &draws, &metric) != 0)
    return fail("device NUTS: buffers");
  for (int i = 0; i < C; ++i)
    if (be.init_chain(be.ctx, i, hc[i].cont.data(), inv_metric0[i].data(), stepsize) != 0)
      return fail("device NUTS: init_chain");

  // ---- writers: headers (run_adaptive_sampler.hpp:66-71) ----
  std::vector<std::unique_ptr<detail::device_sampler_view>> view;
  std::vector<std::unique_ptr<su::mcmc_writer>> writer;
  for (int i = 0; i < C; ++i) {
    view.emplace_back(new detail::device_sampler_view(P));
    view[i]->inv_metric.assign(inv_metric0[i].data(), inv_metric0[i].data() + P);
    writer.emplace_back(new su::mcmc_writer(sample_writer[i], diagnostic_writer[i], logger));
    Eigen::Map<Eigen::VectorXd> cp(hc[i].cont.data(), P);
    stan::mcmc::sample s(cp, 0, 0);
    writer[i]->write_sample_names(s, *view[i], model);
    writer[i]->write_diagnostic_names(s, *view[i], model);
  }

  const int DW = 3 * P + kNutsDrawExtra;   // nuts_draw_doubles(P): q, the 8 scalars, p, g
  const int total = num_warmup + num_samples;
  std::vector<std::int32_t> lanes;
  lanes.reserve(C);
  std::vector<char> finished(C, 0);
  std::vector<std::chrono::steady_clock::time_point> t_start(C, std::chrono::steady_clock::now()), t_warm(C);
  long n_rounds = 0, n_lanes = 0, n_unif = 0, n_norm = 0;
  auto adapt_finish = [&](int i, double eps_nom) {   // run_adaptive_sampler.hpp:83-86
    view[i]->nom_stepsize = eps_nom;
    view[i]->inv_metric.assign(metric + static_cast<size_t>(i) * P, metric + static_cast<size_t>(i + 1) * P);
    writer[i]->write_adapt_finish(*view[i]);
    view[i]->write_sampler_state(sample_writer[i]);
    detail::write_metric_struct(metric_writer[i], eps_nom, view[i]->inv_metric);
    t_warm[i] = std::chrono::steady_clock::now();
  };
  auto refresh_msg = [&](int i, int it) {   // generate_transitions.hpp:52-67; it = 0-based iteration about to run
    if (refresh <= 0 || it >= total)
      return;
    const bool warm = it < num_warmup;
    const int start = warm ? 0 : num_warmup, m = it - start;
    if (!(it + 1 == total || m == 0 || (m + 1) % refresh == 0))
      return;
    const int width = static_cast<int>(std::ceil(std::log10(static_cast<double>(total))));
    std::stringstream message;
    if (C != 1)
      message << "Chain [" << (init_chain_id + i) << "] ";
    message << "Iteration: " << std::setw(width) << it + 1 << " / " << total << " [" << std::setw(3)
            << static_cast<int>((100.0 * (it + 1)) / total) << "%] " << (warm ? " (Warmup)" : " (Sampling)");
    logger.info(message);
  };

  int n_live = C;
  while (n_live > 0) {
    interrupt();
    lanes.clear();
    for (int i = 0; i < C; ++i) {
      if (finished[i])
        continue;
      chain_host& h = hc[i];
      const nuts_status st = status[i];
      // ---- a draw came out of the last round ----
      if (st.iter > h.seen_iter) {
        const double* d = draws + static_cast<size_t>(i) * DW;
        const int it = h.seen_iter;   // 0-based index of the transition that produced it
        h.seen_iter = st.iter;
        detail::device_sampler_view& v = *view[i];
        v.q.assign(d, d + P);
        v.p.assign(d + P + kNutsDrawExtra, d + 2 * P + kNutsDrawExtra);
        v.g.assign(d + 2 * P + kNutsDrawExtra, d + 3 * P + kNutsDrawExtra);
        v.stepsize = d[P + 2];
        v.treedepth = d[P + 3];
        v.n_leapfrog = d[P + 4];
        v.divergent = d[P + 5];
        v.energy = d[P + 6];
        const bool warm = it < num_warmup;
        const int m = warm ? it : it - num_warmup;
        if ((warm ? save_warmup : true) && (m % num_thin) == 0) {
          Eigen::Map<Eigen::VectorXd> q(v.q.data(), P);
          stan::mcmc::sample s(q, d[P + 0], d[P + 1]);
          writer[i]->write_sample_params(h.rng, s, v, model);
          writer[i]->write_diagnostic_params(s, v);
        }
      }
      if (st.phase == kNutsPhaseFailed) {
        static const char* why[] = {"", "Posterior is improper. Please check your model.",
                                    "No acceptably small step size could be found. Perhaps the posterior is not continuous?",
                                    "Numerical overflow in metric adaptation."};
        logger.error(why[st.fail_code >= 0 && st.fail_code <= 3 ? st.fail_code : 0]);
        return error_codes::SOFTWARE;
      }
      // warm-up over (also when there is none): "Adaptation terminated", step size, metric
      if (st.adapt_done && t_warm[i] == std::chrono::steady_clock::time_point())
        adapt_finish(i, st.eps_nom);
      if (st.phase == kNutsPhaseDone) {
        finished[i] = 1;
        --n_live;
        const auto t1 = std::chrono::steady_clock::now();
        writer[i]->write_timing(std::chrono::duration<double>(t_warm[i] - t_start[i]).count(),
                                std::chrono::duration<double>(t1 - t_warm[i]).count());
        continue;
      }
      // ---- randomness for the next round ----
      boost::uniform_01<stan::rng_t&> unif(h.rng);
      double* ur = uniforms + static_cast<size_t>(i) * kNutsUnifStride;
      if (st.need_normals) {
        // put the engine where the reference's would be: after the normal variates of the previous refresh and the
        // uniform variates the device actually consumed since
        h.rng = h.mark;
        for (std::uint64_t k = h.base; k < st.n_unif; ++k)
          (void)unif();
        if (st.phase == 4 /* NPH_TREE */) {
          refresh_msg(i, st.iter);
          if (cfg.stepsize_jitter != 0)   // sample_stepsize() comes first in transition() (base_nuts.hpp:80, then :84 sample_p)
            ur[kNutsUnifCap] = unif();
        }
        boost::variate_generator<stan::rng_t&, boost::normal_distribution<> > gaus(h.rng, boost::normal_distribution<>());
        double* nr = normals + static_cast<size_t>(i) * P;
        for (int k = 0; k < P; ++k)
          nr[k] = gaus();
        ++n_norm;
        h.mark = h.rng;
        h.base = h.gen = st.n_unif;
      }
      while (h.gen < st.n_unif + kNutsUnifCap / 2) {   // speculative: rewound at the next refresh
        ur[h.gen % kNutsUnifCap] = unif();
        ++h.gen;
        ++n_unif;
      }
      lanes.push_back(i);
    }
    if (lanes.empty())
      break;
    if (be.round(be.ctx, static_cast<std::int32_t>(lanes.size()), lanes.data()) != 0)
      return fail("device NUTS: round");
    ++n_rounds;
    n_lanes += static_cast<long>(lanes.size());
  }
  if (stats) {
    stats[0] = n_rounds;
    stats[1] = n_lanes;
    stats[2] = n_unif;
    stats[3] = n_norm;
  }
  return error_codes::OK;
}

}  // namespace b200

#endif
