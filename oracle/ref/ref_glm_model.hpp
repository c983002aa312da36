// oracle/ref/ref_glm_model.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/glm_oracle.h).
//
// The reference itself: a hand-written model class in the shape stanc3 would generate
// (template: /root/reference/src/test/unit/model/model_base_crtp_test.cpp:11-112) whose
// log_prob calls the REFERENCE's own functions --
//   stan::math::bernoulli_logit_glm_lpmf  (lib/stan_math/stan/math/prim/prob/bernoulli_logit_glm_lpmf.hpp:49)
//   stan::math::poisson_log_glm_lpmf      (.../poisson_log_glm_lpmf.hpp:51)
//   stan::math::normal_id_glm_lpdf        (.../normal_id_glm_lpdf.hpp:54)
//   stan::math::binomial_logit_glm_lpmf   (.../binomial_logit_glm_lpmf.hpp:55)
//   stan::math::neg_binomial_2_log_glm_lpmf (.../neg_binomial_2_log_glm_lpmf.hpp:64)
//   stan::math::ordered_logistic_glm_lpmf (.../ordered_logistic_glm_lpmf.hpp:49) with stan::math::ordered_constrain
//   stan::math::categorical_logit_glm_lpmf (.../categorical_logit_glm_lpmf.hpp:43)
//   stan::math::normal_lpdf, stan::math::lb_constrain (prim/constraint/lb_constrain.hpp:60-67),
//   stan::model::rvalue(v, name, index_multi) (src/stan/model/indexing/rvalue.hpp:154-172)
// It is compiled against the headers where they lie under /root/reference; no reference
// source is copied here.
#ifndef ORACLE_REF_GLM_MODEL_HPP
#define ORACLE_REF_GLM_MODEL_HPP

#include <stan/model/model_header.hpp>

#include <string>
#include <vector>

#include "../glm_oracle.h"

namespace oracle_ref {

class ref_glm_model final : public stan::model::model_base_crtp<ref_glm_model> {
 public:
  int family_;
  int64_t N_;
  int K_;
  int G_;
  Eigen::MatrixXd X_;
  std::vector<int> y_int_;
  Eigen::VectorXd y_real_;
  std::vector<int> group_;
  std::vector<int> trials_;
  double prior_alpha_sd_, prior_beta_sd_, prior_sigma_loc_, prior_sigma_scale_,
      prior_sigma_a_scale_;

  int C_ = 0;   // ORDERED_LOGISTIC / CATEGORICAL_LOGIT: number of classes
  bool class_model() const {
    return family_ == GLM_ORDERED_LOGISTIC || family_ == GLM_CATEGORICAL_LOGIT;
  }

  static size_t count_params(const glm_spec& s) {
    if (s.family == GLM_ORDERED_LOGISTIC)
      return s.K + (s.n_classes > 0 ? s.n_classes - 1 : 0);
    if (s.family == GLM_CATEGORICAL_LOGIT)
      return static_cast<size_t>(s.n_classes) * (1 + s.K);
    size_t p = (s.G > 0 ? 2 + s.G : 1) + s.K;
    if (s.family == GLM_NORMAL_ID || s.family == GLM_NEG_BINOMIAL_2_LOG)
      p += 1;
    return p;
  }
  // the trailing positive scalar: sigma (normal_id) or phi (neg_binomial_2_log)
  bool has_scale() const {
    return family_ == GLM_NORMAL_ID || family_ == GLM_NEG_BINOMIAL_2_LOG;
  }
  const char* scale_name() const {
    return family_ == GLM_NEG_BINOMIAL_2_LOG ? "phi" : "sigma";
  }

  explicit ref_glm_model(const glm_spec& s)
      : model_base_crtp(count_params(s)),
        family_(s.family),
        N_(s.N),
        K_(s.K),
        G_(s.G),
        X_(s.N, s.K),
        prior_alpha_sd_(s.prior_alpha_sd),
        prior_beta_sd_(s.prior_beta_sd),
        prior_sigma_loc_(s.prior_sigma_loc),
        prior_sigma_scale_(s.prior_sigma_scale),
        prior_sigma_a_scale_(s.prior_sigma_a_scale) {
    for (int j = 0; j < K_; ++j)
      for (int64_t i = 0; i < N_; ++i)
        X_(i, j) = s.X[i + static_cast<int64_t>(j) * s.ldx];
    if (family_ == GLM_NORMAL_ID) {
      y_real_.resize(N_);
      for (int64_t i = 0; i < N_; ++i)
        y_real_[i] = s.y_real[i];
    } else {
      y_int_.assign(s.y_int, s.y_int + N_);
    }
    if (G_ > 0)
      group_.assign(s.group, s.group + N_);
    if (family_ == GLM_BINOMIAL_LOGIT)
      trials_.assign(s.trials, s.trials + N_);
    if (class_model())
      C_ = s.n_classes;
  }

  ~ref_glm_model() override {}

  std::string model_name() const override { return "ref_glm_model"; }
  std::vector<std::string> model_compile_info() const {
    return {"stanc_version = hand-written (no stanc offline)"};
  }

  void base_names(std::vector<std::string>& names) const {
    if (family_ == GLM_ORDERED_LOGISTIC) {
      for (int k = 1; k <= K_; ++k)
        names.emplace_back("beta." + std::to_string(k));
      for (int c = 1; c < C_; ++c)
        names.emplace_back("c." + std::to_string(c));
      return;
    }
    if (family_ == GLM_CATEGORICAL_LOGIT) {
      for (int c = 1; c <= C_; ++c)
        names.emplace_back("alpha." + std::to_string(c));
      for (int c = 1; c <= C_; ++c)
        for (int k = 1; k <= K_; ++k)
          names.emplace_back("beta." + std::to_string(k) + "." + std::to_string(c));
      return;
    }
    if (G_ > 0) {
      names.emplace_back("mu_a");
      names.emplace_back("sigma_a");
      for (int g = 1; g <= G_; ++g)
        names.emplace_back("a." + std::to_string(g));
    } else {
      names.emplace_back("alpha");
    }
    for (int k = 1; k <= K_; ++k)
      names.emplace_back("beta." + std::to_string(k));
    if (has_scale())
      names.emplace_back(scale_name());
  }

  void get_param_names(std::vector<std::string>& names, bool = true,
                       bool = true) const override {
    names.clear();
    if (family_ == GLM_ORDERED_LOGISTIC) {
      names = {"beta", "c"};
      return;
    }
    if (family_ == GLM_CATEGORICAL_LOGIT) {
      names = {"alpha", "beta"};
      return;
    }
    if (G_ > 0) {
      names = {"mu_a", "sigma_a", "a", "beta"};
    } else {
      names = {"alpha", "beta"};
    }
    if (has_scale())
      names.emplace_back(scale_name());
  }
  void get_dims(std::vector<std::vector<size_t>>& dimss, bool = true,
                bool = true) const override {
    dimss.clear();
    if (family_ == GLM_ORDERED_LOGISTIC) {
      dimss.push_back({static_cast<size_t>(K_)});
      dimss.push_back({static_cast<size_t>(C_ > 0 ? C_ - 1 : 0)});
      return;
    }
    if (family_ == GLM_CATEGORICAL_LOGIT) {
      dimss.push_back({static_cast<size_t>(C_)});
      dimss.push_back({static_cast<size_t>(K_), static_cast<size_t>(C_)});
      return;
    }
    if (G_ > 0) {
      dimss.push_back({});
      dimss.push_back({});
      dimss.push_back({static_cast<size_t>(G_)});
    } else {
      dimss.push_back({});
    }
    dimss.push_back({static_cast<size_t>(K_)});
    if (has_scale())
      dimss.push_back({});
  }
  // stanc-generated models APPEND here (mcmc_writer.hpp:66-77 passes a vector that already holds
  // the sample and sampler column names)
  void constrained_param_names(std::vector<std::string>& names, bool = true,
                               bool = true) const override {
    base_names(names);
  }
  void unconstrained_param_names(std::vector<std::string>& names, bool = true,
                                 bool = true) const override {
    base_names(names);
  }

  // ---- the density ------------------------------------------------------------
  template <bool propto, bool jacobian, typename VecR>
  stan::scalar_type_t<VecR> log_prob_impl(VecR& params_r,
                                          std::ostream* /*msgs*/) const {
    using T = stan::scalar_type_t<VecR>;
    using stan::math::normal_lpdf;
    using vec_t = Eigen::Matrix<T, -1, 1>;
    T lp__(0.0);
    stan::math::accumulator<T> lp_accum__;
    size_t pos = 0;

    if (family_ == GLM_ORDERED_LOGISTIC) {
      // parameters { vector[K] beta; ordered[C-1] c; }
      vec_t beta_o(K_), u(C_ > 0 ? C_ - 1 : 0);
      for (int k = 0; k < K_; ++k)
        beta_o[k] = params_r[pos++];
      for (int c = 0; c + 1 < C_; ++c)
        u[c] = params_r[pos++];
      vec_t cuts = stan::math::ordered_constrain<jacobian>(u, lp__);
      lp_accum__.add(normal_lpdf<propto>(beta_o, 0, prior_beta_sd_));
      lp_accum__.add(normal_lpdf<propto>(cuts, 0, prior_alpha_sd_));
      lp_accum__.add(stan::math::ordered_logistic_glm_lpmf<propto>(y_int_, X_, beta_o, cuts));
      lp_accum__.add(lp__);
      return lp_accum__.sum();
    }
    if (family_ == GLM_CATEGORICAL_LOGIT) {
      // parameters { vector[C] alpha; matrix[K, C] beta; }
      vec_t alpha_c(C_);
      Eigen::Matrix<T, -1, -1> beta_m(K_, C_);
      for (int c = 0; c < C_; ++c)
        alpha_c[c] = params_r[pos++];
      for (int c = 0; c < C_; ++c)
        for (int k = 0; k < K_; ++k)
          beta_m(k, c) = params_r[pos++];
      lp_accum__.add(normal_lpdf<propto>(alpha_c, 0, prior_alpha_sd_));
      lp_accum__.add(normal_lpdf<propto>(stan::math::to_vector(beta_m), 0, prior_beta_sd_));
      lp_accum__.add(stan::math::categorical_logit_glm_lpmf<propto>(y_int_, X_, alpha_c, beta_m));
      lp_accum__.add(lp__);
      return lp_accum__.sum();
    }

    T alpha(0.0), mu_a(0.0), sigma_a(0.0), sigma(0.0);
    vec_t a, beta(K_);
    if (G_ > 0) {
      mu_a = params_r[pos++];
      T u = params_r[pos++];
      sigma_a = jacobian ? stan::math::lb_constrain(u, 0, lp__)
                         : stan::math::lb_constrain(u, 0);
      a.resize(G_);
      for (int g = 0; g < G_; ++g)
        a[g] = params_r[pos++];
    } else {
      alpha = params_r[pos++];
    }
    for (int k = 0; k < K_; ++k)
      beta[k] = params_r[pos++];
    if (has_scale()) {
      T u = params_r[pos++];
      sigma = jacobian ? stan::math::lb_constrain(u, 0, lp__)
                       : stan::math::lb_constrain(u, 0);
    }

    if (G_ > 0) {
      lp_accum__.add(normal_lpdf<propto>(mu_a, 0, prior_alpha_sd_));
      lp_accum__.add(normal_lpdf<propto>(sigma_a, 0, prior_sigma_a_scale_));
      lp_accum__.add(normal_lpdf<propto>(a, mu_a, sigma_a));
    } else {
      lp_accum__.add(normal_lpdf<propto>(alpha, 0, prior_alpha_sd_));
    }
    lp_accum__.add(normal_lpdf<propto>(beta, 0, prior_beta_sd_));
    if (has_scale())
      lp_accum__.add(
          normal_lpdf<propto>(sigma, prior_sigma_loc_, prior_sigma_scale_));

    auto likelihood = [&](const auto& intercept) -> T {
      switch (family_) {
        case GLM_BERNOULLI_LOGIT:
          return stan::math::bernoulli_logit_glm_lpmf<propto>(y_int_, X_,
                                                              intercept, beta);
        case GLM_POISSON_LOG:
          return stan::math::poisson_log_glm_lpmf<propto>(y_int_, X_, intercept,
                                                          beta);
        case GLM_BINOMIAL_LOGIT:
          return stan::math::binomial_logit_glm_lpmf<propto>(
              y_int_, trials_, X_, intercept, beta);
        case GLM_NEG_BINOMIAL_2_LOG:
          return stan::math::neg_binomial_2_log_glm_lpmf<propto>(
              y_int_, X_, intercept, beta, sigma);
        default:
          return stan::math::normal_id_glm_lpdf<propto>(y_real_, X_, intercept,
                                                        beta, sigma);
      }
    };
    if (G_ > 0) {
      lp_accum__.add(likelihood(stan::model::rvalue(
          a, "a", stan::model::index_multi(group_))));
    } else {
      lp_accum__.add(likelihood(alpha));
    }
    lp_accum__.add(lp__);
    return lp_accum__.sum();
  }

  template <bool propto, bool jacobian, typename T>
  T log_prob(Eigen::Matrix<T, -1, 1>& params_r, std::ostream* msgs) const {
    return log_prob_impl<propto, jacobian>(params_r, msgs);
  }
  template <bool propto, bool jacobian, typename T>
  T log_prob(std::vector<T>& params_r, std::vector<int>& /*params_i*/,
             std::ostream* msgs) const {
    return log_prob_impl<propto, jacobian>(params_r, msgs);
  }

  // ---- constrain / unconstrain ------------------------------------------------
  template <typename VecIn, typename VecOut>
  void constrain_impl(const VecIn& u, VecOut& c) const {
    const size_t P = num_params_r();
    for (size_t i = 0; i < P; ++i)
      c[i] = u[i];
    if (family_ == GLM_ORDERED_LOGISTIC) {
      for (int k = 1; k + 1 < C_; ++k)
        c[K_ + k] = c[K_ + k - 1] + std::exp(u[K_ + k]);
      return;
    }
    if (family_ == GLM_CATEGORICAL_LOGIT)
      return;
    if (G_ > 0)
      c[1] = std::exp(u[1]);
    if (has_scale())
      c[P - 1] = std::exp(u[P - 1]);
  }
  template <typename VecIn, typename VecOut>
  void unconstrain_impl(const VecIn& c, VecOut& u) const {
    const size_t P = num_params_r();
    for (size_t i = 0; i < P; ++i)
      u[i] = c[i];
    if (family_ == GLM_ORDERED_LOGISTIC) {
      for (int k = 1; k + 1 < C_; ++k)
        u[K_ + k] = std::log(c[K_ + k] - c[K_ + k - 1]);   // ordered_free
      return;
    }
    if (family_ == GLM_CATEGORICAL_LOGIT)
      return;
    if (G_ > 0)
      u[1] = stan::math::lb_free(c[1], 0);
    if (has_scale())
      u[P - 1] = stan::math::lb_free(c[P - 1], 0);
  }

  template <typename RNG>
  void write_array(RNG& /*rng*/, Eigen::VectorXd& params_r,
                   Eigen::VectorXd& vars, bool = true, bool = true,
                   std::ostream* = nullptr) const {
    vars.resize(num_params_r());
    constrain_impl(params_r, vars);
  }
  template <typename RNG>
  void write_array(RNG& /*rng*/, std::vector<double>& params_r,
                   std::vector<int>& /*params_i*/, std::vector<double>& vars,
                   bool = true, bool = true, std::ostream* = nullptr) const {
    vars.resize(num_params_r());
    constrain_impl(params_r, vars);
  }

  void read_context(const stan::io::var_context& context,
                    std::vector<double>& constrained) const {
    std::vector<std::string> names;
    get_param_names(names);
    constrained.clear();
    for (const auto& nm : names) {
      std::vector<double> v = context.vals_r(nm);
      constrained.insert(constrained.end(), v.begin(), v.end());
    }
    if (constrained.size() != num_params_r())
      throw std::invalid_argument("init context has wrong number of values");
  }
  void transform_inits(const stan::io::var_context& context,
                       Eigen::VectorXd& params_r,
                       std::ostream* = nullptr) const override {
    std::vector<double> c;
    read_context(context, c);
    params_r.resize(num_params_r());
    unconstrain_impl(c, params_r);
  }
  void transform_inits(const stan::io::var_context& context,
                       std::vector<int>& /*params_i*/,
                       std::vector<double>& params_r,
                       std::ostream* = nullptr) const override {
    std::vector<double> c;
    read_context(context, c);
    params_r.resize(num_params_r());
    unconstrain_impl(c, params_r);
  }
  void unconstrain_array(const Eigen::VectorXd& c, Eigen::VectorXd& u,
                         std::ostream* = nullptr) const override {
    u.resize(num_params_r());
    unconstrain_impl(c, u);
  }
  void unconstrain_array(const std::vector<double>& c, std::vector<double>& u,
                         std::ostream* = nullptr) const override {
    u.resize(num_params_r());
    unconstrain_impl(c, u);
  }
};

}  // namespace oracle_ref
#endif
