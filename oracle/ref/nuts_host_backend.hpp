// oracle/ref/nuts_host_backend.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/glm_oracle.h).
//
// A HOST backend for b200::hmc_nuts_diag_e_adapt_device (stan_b200/cpp/b200/device_nuts.hpp): the product's per-chain
// NUTS state machine (stan_b200/csrc/nuts_tree.cuh, the code the GPU kernels run with one warp per chain) compiled for
// the host with the one-lane policy, and every leapfrog step done by the REFERENCE's own integrator on the reference's
// own model (expl_leapfrog<diag_e_metric<ref_glm_model>>::evolve).  With it the product's driver + state machine can be
// run against stan::services::sample::hmc_nuts_diag_e_adapt on the same seeds without a GPU: what differs is only what
// this round moved to the device (tree building, U-turn checks, multinomial sampling, adaptation, init_stepsize).
#ifndef ORACLE_REF_NUTS_HOST_BACKEND_HPP
#define ORACLE_REF_NUTS_HOST_BACKEND_HPP

#include <b200/device_nuts.hpp>
#include <nuts_tree.cuh>

#include <stan/mcmc/hmc/hamiltonians/diag_e_metric.hpp>
#include <stan/mcmc/hmc/integrators/expl_leapfrog.hpp>

#include <string>
#include <vector>

namespace oracle_ref {

static_assert(sizeof(b200::nuts_status) == sizeof(b200glm::NutsStatus), "status mirror out of date");

template <class Model>
struct nuts_host_backend {
  using H = stan::mcmc::diag_e_metric<Model, stan::rng_t>;
  const Model& model;
  int P = 0, C = 0;
  b200glm::NutsConfig cfg;
  std::vector<b200glm::NutsChain> chains;
  std::vector<double> vec, Q, Pm, Gd, IM, V, normals, uniforms, draws, metric;
  std::vector<b200glm::NutsStatus> status;
  size_t vstride = 0;
  long n_leapfrogs = 0;
  std::string err;

  explicit nuts_host_backend(const Model& m) : model(m), P(static_cast<int>(m.num_params_r())) {}

  b200glm::NutsSlot slot(int c) {
    return b200glm::NutsSlot{Q.data(), Pm.data(), Gd.data(), IM.data(), V.data(), static_cast<size_t>(C), c};
  }

  static int reserve(void* ctx, std::int32_t n, const b200::nuts_config* c) {
    auto& b = *static_cast<nuts_host_backend*>(ctx);
    b.C = n;
    b.cfg.P = b.P;
    b.cfg.max_depth = c->max_depth;
    b.cfg.max_deltaH = c->max_deltaH;
    b.cfg.delta = c->delta;
    b.cfg.gamma = c->gamma;
    b.cfg.kappa = c->kappa;
    b.cfg.t0 = c->t0;
    b.cfg.w_num_warmup = c->w_num_warmup;
    b.cfg.w_init_buffer = c->w_init_buffer;
    b.cfg.w_term_buffer = c->w_term_buffer;
    b.cfg.w_base_window = c->w_base_window;
    b.cfg.w_size0 = c->w_size0;
    b.cfg.w_next0 = c->w_next0;
    b.cfg.num_warmup = c->num_warmup;
    b.cfg.num_samples = c->num_samples;
    b.cfg.stepsize_jitter = c->stepsize_jitter;
    b.vstride = b200glm::nuts_vec_doubles(b.P, c->max_depth);
    const size_t PC = static_cast<size_t>(b.P) * n;
    b.chains.assign(n, b200glm::NutsChain());
    b.vec.assign(b.vstride * n, 0.0);
    for (auto* a : {&b.Q, &b.Pm, &b.Gd, &b.IM, &b.normals, &b.metric})
      a->assign(PC, 0.0);
    b.V.assign(n, 0.0);
    b.uniforms.assign(static_cast<size_t>(n) * b200glm::NUTS_UNIF_STRIDE, 0.0);
    b.draws.assign(static_cast<size_t>(n) * b200glm::nuts_draw_doubles(b.P), 0.0);
    b.status.assign(n, b200glm::NutsStatus());
    return 0;
  }
  static int buffers(void* ctx, double** normals, double** uniforms, b200::nuts_status** status, double** draws,
                     double** metric) {
    auto& b = *static_cast<nuts_host_backend*>(ctx);
    *normals = b.normals.data();
    *uniforms = b.uniforms.data();
    *status = reinterpret_cast<b200::nuts_status*>(b.status.data());
    *draws = b.draws.data();
    *metric = b.metric.data();
    return 0;
  }
  static int init_chain(void* ctx, std::int32_t c, const double* q0, const double* inv_metric, double stepsize) {
    auto& b = *static_cast<nuts_host_backend*>(ctx);
    b200glm::NutsSlot s = b.slot(c);
    for (int k = 0; k < b.P; ++k) {
      s.q(k) = q0[k];
      s.p(k) = 0.0;
      s.g(k) = 0.0;
      s.im(k) = inv_metric[k];
      b.metric[static_cast<size_t>(c) * b.P + k] = inv_metric[k];
    }
    s.v() = 0.0;
    b200glm::nuts_chain_init<b200glm::NutsOneLane>(b.cfg, b.chains[c], b.vec.data() + b.vstride * c, s, stepsize);
    b200glm::nuts_publish(b.chains[c], b.status[c]);
    return 0;
  }
  static int round(void* ctx, std::int32_t n, const std::int32_t* lanes) {
    auto& b = *static_cast<nuts_host_backend*>(ctx);
    using LN = b200glm::NutsOneLane;
    try {
      H ham(b.model);
      stan::mcmc::expl_leapfrog<H> integrator;
      stan::callbacks::logger logger;
      stan::mcmc::diag_e_point z(b.P);
      for (int i = 0; i < n; ++i) {
        const int c = lanes[i];
        b200glm::NutsChain& ch = b.chains[c];
        double* v = b.vec.data() + b.vstride * c;
        b200glm::NutsSlot s = b.slot(c);
        if (ch.need_normals
            && (ch.phase == b200glm::NPH_SS_FIRST || ch.phase == b200glm::NPH_SS_LOOP || ch.phase == b200glm::NPH_TREE))
          b200glm::nuts_begin<LN>(b.cfg, ch, v, s, b.normals.data() + static_cast<size_t>(c) * b.P,
                                  b.uniforms.data() + static_cast<size_t>(c) * b200glm::NUTS_UNIF_STRIDE);
        for (int k = 0; k < b.P; ++k) {
          z.q(k) = s.q(k);
          z.p(k) = s.p(k);
          z.g(k) = s.g(k);
          z.inv_e_metric_(k) = s.im(k);
        }
        z.V = s.v();
        if (ch.phase == b200glm::NPH_INIT_GRAD)
          ham.init(z, logger);                               // the gradient at the initial point
        else
          integrator.evolve(z, ham, ch.lane_eps, logger);    // the reference's leapfrog step
        ++b.n_leapfrogs;
        for (int k = 0; k < b.P; ++k) {
          s.q(k) = z.q(k);
          s.p(k) = z.p(k);
          s.g(k) = z.g(k);
        }
        s.v() = z.V;
        b200glm::nuts_after_leapfrog<LN>(b.cfg, ch, v, s,
                                         b.uniforms.data() + static_cast<size_t>(c) * b200glm::NUTS_UNIF_STRIDE,
                                         b.draws.data() + static_cast<size_t>(c) * b200glm::nuts_draw_doubles(b.P),
                                         b.metric.data() + static_cast<size_t>(c) * b.P);
        b200glm::nuts_publish(ch, b.status[c]);
      }
    } catch (const std::exception& e) {
      b.err = e.what();
      return 3;
    }
    return 0;
  }
  static const char* last_error(void* ctx) { return static_cast<nuts_host_backend*>(ctx)->err.c_str(); }

  b200::nuts_backend table() { return b200::nuts_backend{this, &reserve, &buffers, &init_chain, &round, &last_error}; }
};

}  // namespace oracle_ref

#endif
