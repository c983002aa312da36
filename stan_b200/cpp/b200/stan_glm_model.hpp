// b200/stan_glm_model.hpp -- reference-side binding of the B200 GLM backend.
//
// Include this header BEFORE the first use of stan::services::sample::hmc_nuts_diag_e_adapt with
// b200::glm_model.  It provides
//   (1) b200::glm_model : stan::model::model_base_crtp<glm_model>      (ST/model/model_base_crtp.hpp:74-301)
//       -- log_prob<propto,jacobian,double> -> b200glm_log_prob            (value, all constants: initialize.hpp:128)
//       -- log_prob<propto,jacobian,var>    -> b200glm_log_prob_grad + stan::math::precomputed_gradients
//          (SM/rev/core/precomputed_gradients.hpp:211-224), so log_prob_grad (ST/model/log_prob_grad.hpp:29-50)
//          and every other tape-based caller work unchanged;
//   (2) an explicit specialisation of stan::model::gradient<b200::glm_model> (ST/model/gradient.hpp:22-35).
//       base_hamiltonian::update_potential_gradient (ST/mcmc/hmc/hamiltonians/base_hamiltonian.hpp:61-70) calls
//       it qualified, so only a specialisation (not a later overload) is seen; it skips the AD tape;
//   (3) a full specialisation of stan::mcmc::expl_leapfrog<diag_e_metric<b200::glm_model, stan::rng_t>>
//       (ST/mcmc/hmc/integrators/expl_leapfrog.hpp:16-32) whose evolve() is ONE fused device launch on
//       device-resident (q,p,g); base_hmc only ever calls integrator_.evolve (base_hmc.hpp:113,132; base_nuts.hpp:254).
//       It reaches the model through the Hamiltonian it is handed (no thread-local lookup).
// With these, the UNMODIFIED adapt_diag_e_nuts / base_nuts / diag_e_metric / services drive the GPU path.
//
// Threading: chains run as TBB tasks on several host threads sharing one const model (hmc_nuts_diag_e_adapt.hpp:387-401).
// Each host thread is bound to its own device slot (stream + workspace); give the model n_slots >= number of threads.
#ifndef B200_STAN_GLM_MODEL_HPP
#define B200_STAN_GLM_MODEL_HPP

#include <stan/model/model_header.hpp>
#include <stan/model/gradient.hpp>
#include <stan/mcmc/hmc/hamiltonians/diag_e_metric.hpp>
#include <stan/mcmc/hmc/integrators/expl_leapfrog.hpp>

#include <b200glm.h>

#include <atomic>
#include <cstring>
#include <memory>
#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace b200 {

// What the var_context constructor needs to know about the data block of the Stan program the model
// stands for (a stanc-generated model has this baked in): the names of the data variables, the
// likelihood family, the prior scales and where to put the data.
struct glm_config {
  int family = B200GLM_BERNOULLI_LOGIT;
  std::string name_N = "N", name_K = "K", name_X = "X", name_y = "y", name_G = "G", name_group = "group",
              name_trials = "trials";   // binomial_logit: array[N] int of population sizes
  // brms-style `transformed data`: Xc[, k] = X[, k] - mean(X[, k]); the intercept parameter is then the
  // intercept for centred predictors and b_Intercept = Intercept - dot(means_X, b) (glm_model::means_x())
  bool center_x = false;
  double prior_alpha_sd = 2.5, prior_beta_sd = 2.5, prior_sigma_loc = 1.0, prior_sigma_scale = 2.0,
         prior_sigma_a_scale = 1.0;
  int device = 0, n_slots = 1;
};

class glm_model final : public stan::model::model_base_crtp<glm_model> {
  // data read from a var_context, alive until b200glm_create has copied it to the device
  struct loaded_data {
    b200glm_desc desc;
    std::vector<double> X, y_real, means;
    std::vector<int> y_int, group, trials;
  };
  static loaded_data load(const stan::io::var_context& context, const glm_config& cfg) {
    // same checks and messages a stanc-generated constructor performs (validate_dims / vals_i / vals_r)
    loaded_data L;
    std::memset(&L.desc, 0, sizeof(L.desc));
    const char* stage = "data initialization";
    context.validate_dims(stage, cfg.name_N, "int", std::vector<size_t>{});
    const int N = context.vals_i(cfg.name_N)[0];
    stan::math::check_greater_or_equal("model constructor", cfg.name_N.c_str(), N, 0);
    context.validate_dims(stage, cfg.name_K, "int", std::vector<size_t>{});
    const int K = context.vals_i(cfg.name_K)[0];
    stan::math::check_greater_or_equal("model constructor", cfg.name_K.c_str(), K, 0);
    context.validate_dims(stage, cfg.name_X, "double", std::vector<size_t>{static_cast<size_t>(N), static_cast<size_t>(K)});
    L.X = context.vals_r(cfg.name_X);               // matrix[N, K]: column-major, as Eigen::MatrixXd
    if (cfg.center_x) {
      L.means.assign(K, 0.0);
      for (int k = 0; k < K; ++k) {
        double* col = L.X.data() + static_cast<size_t>(k) * N;
        // stan::math::mean == Eigen's mean(): sum / N
        const double mu = N > 0 ? Eigen::Map<const Eigen::VectorXd>(col, N).mean() : 0.0;
        L.means[k] = mu;
        for (int i = 0; i < N; ++i)
          col[i] -= mu;
      }
    }
    if (cfg.family == B200GLM_NORMAL_ID) {
      context.validate_dims(stage, cfg.name_y, "double", std::vector<size_t>{static_cast<size_t>(N)});
      L.y_real = context.vals_r(cfg.name_y);
    } else {
      context.validate_dims(stage, cfg.name_y, "int", std::vector<size_t>{static_cast<size_t>(N)});
      L.y_int = context.vals_i(cfg.name_y);
    }
    if (cfg.family == B200GLM_BINOMIAL_LOGIT) {
      context.validate_dims(stage, cfg.name_trials, "int", std::vector<size_t>{static_cast<size_t>(N)});
      L.trials = context.vals_i(cfg.name_trials);
    }
    int G = 0;
    if (context.contains_i(cfg.name_G)) {
      context.validate_dims(stage, cfg.name_G, "int", std::vector<size_t>{});
      G = context.vals_i(cfg.name_G)[0];
      stan::math::check_greater_or_equal("model constructor", cfg.name_G.c_str(), G, 0);
    }
    if (G > 0) {
      context.validate_dims(stage, cfg.name_group, "int", std::vector<size_t>{static_cast<size_t>(N)});
      L.group = context.vals_i(cfg.name_group);
      stan::math::check_bounded("model constructor", cfg.name_group.c_str(), L.group, 1, G);
    }
    b200glm_desc& d = L.desc;
    d.family = cfg.family;
    d.N = N;
    d.K = K;
    d.X = L.X.data();
    d.ldx = N > 0 ? N : 1;
    d.y_int = L.y_int.empty() ? nullptr : L.y_int.data();
    d.y_real = L.y_real.empty() ? nullptr : L.y_real.data();
    d.G = G;
    d.group = L.group.empty() ? nullptr : L.group.data();
    d.trials = L.trials.empty() ? nullptr : L.trials.data();
    d.prior_alpha_sd = cfg.prior_alpha_sd;
    d.prior_beta_sd = cfg.prior_beta_sd;
    d.prior_sigma_loc = cfg.prior_sigma_loc;
    d.prior_sigma_scale = cfg.prior_sigma_scale;
    d.prior_sigma_a_scale = cfg.prior_sigma_a_scale;
    d.device = cfg.device;
    d.n_slots = cfg.n_slots;
    d.world = 1;
    return L;
  }
  explicit glm_model(loaded_data&& L) : glm_model(L.desc) { means_x_ = std::move(L.means); }

 public:
  // The constructor signature of a stanc-generated model (model(var_context&, seed, ostream*)): reads the
  // data block from any stan::io::var_context -- stan::json::json_data (ST/io/json/json_data.hpp:42),
  // stan::io::dump, array_var_context -- and uploads it.  `cfg` stands for what stanc would have compiled in.
  glm_model(const stan::io::var_context& context, const glm_config& cfg, unsigned int /*random_seed*/ = 0,
            std::ostream* /*pstream*/ = nullptr)
      : glm_model(load(context, cfg)) {}

  // column means removed from X by glm_config::center_x (empty otherwise)
  const std::vector<double>& means_x() const { return means_x_; }

  static size_t count_params(const b200glm_desc& d) {
    if (d.family == B200GLM_ORDERED_LOGISTIC)       // parameters { vector[K] beta; ordered[C-1] c; }
      return d.K + (d.n_classes > 0 ? d.n_classes - 1 : 0);
    if (d.family == B200GLM_CATEGORICAL_LOGIT)      // parameters { vector[C] alpha; matrix[K, C] beta; }
      return static_cast<size_t>(d.n_classes) * (1 + d.K);
    return (d.G > 0 ? 2 + d.G : 1) + d.K + (has_scale(d.family) ? 1 : 0);
  }
  bool is_ordered() const { return desc_.family == B200GLM_ORDERED_LOGISTIC; }
  bool is_categorical() const { return desc_.family == B200GLM_CATEGORICAL_LOGIT; }
  // families with a trailing positive scalar parameter: sigma (normal_id) or phi (neg_binomial_2_log)
  static bool has_scale(int family) { return family == B200GLM_NORMAL_ID || family == B200GLM_NEG_BINOMIAL_2_LOG; }
  const char* scale_name() const { return desc_.family == B200GLM_NEG_BINOMIAL_2_LOG ? "phi" : "sigma"; }

  explicit glm_model(const b200glm_desc& desc)
      : model_base_crtp(count_params(desc)), desc_(desc), h_(nullptr), uid_(next_uid()) {
    if (desc_.n_slots < 1)
      desc_.n_slots = 1;
    guards_.reset(new slot_guard[desc_.n_slots]);
    if (b200glm_abi_version() != B200GLM_ABI_VERSION)   // a stale libb200glm.so would misread b200glm_desc
      throw std::runtime_error("b200glm: library ABI " + std::to_string(b200glm_abi_version()) + " != header ABI "
                               + std::to_string(B200GLM_ABI_VERSION));
    const int rc = b200glm_create(&desc_, &h_);
    if (rc != B200GLM_OK) {
      std::string msg = h_ ? b200glm_last_error(h_) : "b200glm_create failed";
      if (h_)
        b200glm_destroy(h_);
      h_ = nullptr;
      raise(rc, msg);
    }
    // the host pointers in desc are not retained
    desc_.X = nullptr;
    desc_.y_int = nullptr;
    desc_.y_real = nullptr;
    desc_.group = nullptr;
    desc_.trials = nullptr;
    registry(+1, uid_);
  }
  glm_model(const glm_model&) = delete;
  glm_model& operator=(const glm_model&) = delete;
  ~glm_model() override {
    registry(-1, uid_);
    if (h_)
      b200glm_destroy(h_);
  }

  // Thread-local caches (slot binding, resident leapfrog state) are keyed by a
  // process-unique id, never by address: a new model may be allocated where a destroyed one lived.
  static unsigned long long next_uid() {
    static std::atomic<unsigned long long> n{1};
    return n.fetch_add(1);
  }
  unsigned long long uid() const { return uid_; }
  // op = +1 register, -1 unregister, 0 query
  static bool registry(int op, unsigned long long uid) {
    static std::mutex mu;
    static std::set<unsigned long long> live;
    std::lock_guard<std::mutex> lock(mu);
    if (op > 0)
      live.insert(uid);
    else if (op < 0)
      live.erase(uid);
    return live.count(uid) != 0;
  }
  static unsigned long long thread_token() {
    static std::atomic<unsigned long long> n{1};
    static thread_local unsigned long long tok = n.fetch_add(1);
    return tok;
  }

  b200glm_handle* handle() const { return h_; }
  const b200glm_desc& desc() const { return desc_; }

  // ---------------------------------------------------------------- device calls
  [[noreturn]] static void raise(int rc, const std::string& msg) {
    if (rc == B200GLM_DOMAIN)
      throw std::domain_error(msg);
    if (rc == B200GLM_INVALID)
      throw std::invalid_argument(msg);
    throw std::runtime_error("b200glm: " + msg);
  }
  void check(int rc) const {
    if (rc != B200GLM_OK)
      raise(rc, b200glm_last_error(h_));
  }

  // one device slot per host thread (round-robin over n_slots; with more threads than slots two
  // threads share one, which is safe -- see slot_guard -- but costs a state upload per switch)
  int slot() const {
    static thread_local int tls_slot = -1;
    static thread_local unsigned long long tls_owner = 0;
    if (tls_owner != uid_) {
      tls_owner = uid_;
      tls_slot = next_slot_.fetch_add(1) % desc_.n_slots;
    }
    return tls_slot;
  }
  // Serialises the {set_state, leapfrog} pair on a slot and remembers which thread's trajectory the
  // slot's device-resident state belongs to.
  struct slot_guard {
    std::mutex mu;
    unsigned long long state_token = 0, metric_token = 0;
  };

  // Batched-chains hook (b200/batched_nuts.hpp): while a host thread runs a chain inside a
  // batch_scope, its gradient and leapfrog calls are not launched one by one but handed to the
  // batcher, which serves all waiting chains with ONE pass over X (b200glm_leapfrog_batched).
  struct batch_hook {
    virtual ~batch_hook() {}
    virtual void gradient(int chain, const double* theta, double& lp, double* grad) = 0;
    virtual void leapfrog(int chain, Eigen::VectorXd& q, Eigen::VectorXd& p, Eigen::VectorXd& g, double& V,
                          const Eigen::VectorXd& inv_metric, double epsilon, stan::callbacks::logger& logger) = 0;
    virtual void leave(int chain) = 0;
  };
  static batch_hook*& tls_hook() {
    static thread_local batch_hook* hook = nullptr;
    return hook;
  }
  static int& tls_chain() {
    static thread_local int chain = -1;
    return chain;
  }
  struct batch_scope {
    batch_scope(batch_hook* hook, int chain) {
      tls_hook() = hook;
      tls_chain() = chain;
    }
    ~batch_scope() {
      batch_hook* hook = tls_hook();
      tls_hook() = nullptr;
      if (hook)
        hook->leave(tls_chain());
      tls_chain() = -1;
    }
  };

  // on_tape: the caller is log_prob<..., var> with autodiff variables of the thread's tape alive.  Such a call is
  // never handed to the batcher: the batched driver runs all chains as fibers of ONE thread, and a chain parked in
  // the middle of a tape computation would have its variables recovered by the next chain's log_prob_grad.
  void device_log_prob_grad(const double* theta, bool propto, bool jacobian,
                            double& lp, double* grad, bool on_tape = false) const {
    n_gradients_.fetch_add(1, std::memory_order_relaxed);
    if (batch_hook* hook = tls_hook(); hook && propto && jacobian && !on_tape) {
      hook->gradient(tls_chain(), theta, lp, grad);
      return;
    }
    check(b200glm_log_prob_grad(h_, slot(), theta, propto, jacobian, &lp, grad));
  }
  double device_log_prob(const double* theta, bool propto, bool jacobian) const {
    double lp = 0;
    check(b200glm_log_prob(h_, slot(), theta, propto, jacobian, &lp));
    return lp;
  }

  struct resident_state {
    unsigned long long owner = 0;
    bool valid = false, metric_valid = false;
    std::vector<double> q, p, g, inv_metric;
  };
  static resident_state& resident() {
    static thread_local resident_state rs;
    return rs;
  }

  // expl_leapfrog::evolve on the device.  (q,p,g) stay resident between consecutive steps of a
  // trajectory; they are re-uploaded only when the caller's z is not the state the device produced last.
  void device_leapfrog(Eigen::VectorXd& q, Eigen::VectorXd& p, Eigen::VectorXd& g, double& V,
                       const Eigen::VectorXd& inv_metric, double epsilon,
                       stan::callbacks::logger& logger) const {
    n_leapfrogs_.fetch_add(1, std::memory_order_relaxed);
    if (batch_hook* hook = tls_hook()) {
      hook->leapfrog(tls_chain(), q, p, g, V, inv_metric, epsilon, logger);
      return;
    }
    const size_t P = num_params_r();
    const size_t bytes = P * sizeof(double);
    resident_state& rs = resident();
    const int sl = slot();
    slot_guard& guard = guards_[sl];
    const unsigned long long me = thread_token();
    std::lock_guard<std::mutex> lock(guard.mu);
    if (rs.owner != uid_) {
      rs.owner = uid_;
      rs.valid = rs.metric_valid = false;
      rs.q.resize(P);
      rs.p.resize(P);
      rs.g.resize(P);
      rs.inv_metric.resize(P);
    }
    const bool same = rs.valid && guard.state_token == me && std::memcmp(rs.q.data(), q.data(), bytes) == 0
                      && std::memcmp(rs.p.data(), p.data(), bytes) == 0
                      && std::memcmp(rs.g.data(), g.data(), bytes) == 0;
    if (!same) {
      check(b200glm_set_state(h_, sl, q.data(), p.data(), g.data(), V));
      n_uploads_.fetch_add(1, std::memory_order_relaxed);
    }
    const double* im = nullptr;
    if (!rs.metric_valid || guard.metric_token != me
        || std::memcmp(rs.inv_metric.data(), inv_metric.data(), bytes) != 0) {
      std::memcpy(rs.inv_metric.data(), inv_metric.data(), bytes);
      rs.metric_valid = true;
      guard.metric_token = me;
      im = rs.inv_metric.data();
    }
    guard.state_token = me;
    const int rc = b200glm_leapfrog(h_, sl, epsilon, im, q.data(), p.data(), g.data(), &V);
    if (rc == B200GLM_DOMAIN) {
      // data-level domain error (y out of range): same outcome as base_hamiltonian.hpp:65-69
      V = std::numeric_limits<double>::infinity();
      rs.valid = false;
      reject_message(b200glm_last_error(h_), logger);
      return;
    }
    check(rc);
    if (V == std::numeric_limits<double>::infinity())
      reject_message("non-finite log density or gradient", logger);
    std::memcpy(rs.q.data(), q.data(), bytes);
    std::memcpy(rs.p.data(), p.data(), bytes);
    std::memcpy(rs.g.data(), g.data(), bytes);
    rs.valid = true;
  }

  static void reject_message(const std::string& what, stan::callbacks::logger& logger) {
    // wording of base_hamiltonian::write_error_msg_ (base_hamiltonian.hpp:83-96)
    logger.error("Informational Message: The current Metropolis proposal is about to be rejected because of "
                 "the following issue:");
    logger.error(what);
    logger.error("If this warning occurs sporadically, such as for highly constrained variable types like "
                 "covariance matrices, then the sampler is fine,");
    logger.error("but if this warning occurs often then your model may be either severely ill-conditioned or "
                 "misspecified.");
    logger.error("");
  }

  long n_gradients() const { return n_gradients_.load(); }
  long n_leapfrogs() const { return n_leapfrogs_.load(); }
  long n_uploads() const { return n_uploads_.load(); }
  void count_upload(long n = 1) const { n_uploads_.fetch_add(n, std::memory_order_relaxed); }

  // ---------------------------------------------------------------- model_base interface
  std::string model_name() const override { return "b200_glm_model"; }
  std::vector<std::string> model_compile_info() const {
    return {std::string("backend = ") + b200glm_version()};
  }

  // appends, as stanc-generated models do: mcmc_writer::write_sample_names (services/util/mcmc_writer.hpp:66-77)
  // passes a vector that already holds the sample and sampler column names
  void flat_names(std::vector<std::string>& names) const {
    if (is_ordered()) {
      for (int k = 1; k <= desc_.K; ++k)
        names.emplace_back("beta." + std::to_string(k));
      for (int j = 1; j < desc_.n_classes; ++j)
        names.emplace_back("c." + std::to_string(j));
      return;
    }
    if (is_categorical()) {
      for (int c = 1; c <= desc_.n_classes; ++c)
        names.emplace_back("alpha." + std::to_string(c));
      for (int c = 1; c <= desc_.n_classes; ++c)     // matrix[K, C] in column-major order, as stanc writes it
        for (int k = 1; k <= desc_.K; ++k)
          names.emplace_back("beta." + std::to_string(k) + "." + std::to_string(c));
      return;
    }
    if (desc_.G > 0) {
      names.emplace_back("mu_a");
      names.emplace_back("sigma_a");
      for (int g = 1; g <= desc_.G; ++g)
        names.emplace_back("a." + std::to_string(g));
    } else {
      names.emplace_back("alpha");
    }
    for (int k = 1; k <= desc_.K; ++k)
      names.emplace_back("beta." + std::to_string(k));
    if (has_scale(desc_.family))
      names.emplace_back(scale_name());
  }
  void get_param_names(std::vector<std::string>& names, bool = true, bool = true) const override {
    if (is_ordered()) {
      names = {"beta", "c"};
      return;
    }
    if (is_categorical()) {
      names = {"alpha", "beta"};
      return;
    }
    if (desc_.G > 0)
      names = {"mu_a", "sigma_a", "a", "beta"};
    else
      names = {"alpha", "beta"};
    if (has_scale(desc_.family))
      names.emplace_back(scale_name());
  }
  void get_dims(std::vector<std::vector<size_t>>& dimss, bool = true, bool = true) const override {
    dimss.clear();
    if (is_ordered()) {
      dimss.push_back({static_cast<size_t>(desc_.K)});
      dimss.push_back({static_cast<size_t>(desc_.n_classes > 0 ? desc_.n_classes - 1 : 0)});
      return;
    }
    if (is_categorical()) {
      dimss.push_back({static_cast<size_t>(desc_.n_classes)});
      dimss.push_back({static_cast<size_t>(desc_.K), static_cast<size_t>(desc_.n_classes)});
      return;
    }
    if (desc_.G > 0) {
      dimss.push_back({});
      dimss.push_back({});
      dimss.push_back({static_cast<size_t>(desc_.G)});
    } else {
      dimss.push_back({});
    }
    dimss.push_back({static_cast<size_t>(desc_.K)});
    if (has_scale(desc_.family))
      dimss.push_back({});
  }
  void constrained_param_names(std::vector<std::string>& names, bool = true, bool = true) const override {
    flat_names(names);
  }
  void unconstrained_param_names(std::vector<std::string>& names, bool = true, bool = true) const override {
    flat_names(names);
  }

  template <bool propto, bool jacobian, typename T, typename Vec>
  T log_prob_any(Vec& params_r) const {
    const size_t P = num_params_r();
    if (static_cast<size_t>(params_r.size()) != P)
      throw std::invalid_argument("b200::glm_model::log_prob: wrong number of parameters");
    if constexpr (std::is_same<T, double>::value) {
      return device_log_prob(params_r.data(), propto, jacobian);
    } else {
      static_assert(std::is_same<T, stan::math::var>::value,
                    "b200::glm_model supports double and reverse-mode var only (NUTS needs first order)");
      std::vector<double> th(P), grad(P);
      std::vector<stan::math::var> ops(P);
      for (size_t i = 0; i < P; ++i) {
        ops[i] = params_r[i];
        th[i] = ops[i].val();
      }
      double lp = 0;
      device_log_prob_grad(th.data(), propto, jacobian, lp, grad.data(), /*on_tape=*/true);
      return stan::math::precomputed_gradients(lp, ops, grad);
    }
  }
  template <bool propto, bool jacobian, typename T>
  T log_prob(Eigen::Matrix<T, -1, 1>& params_r, std::ostream* = nullptr) const {
    return log_prob_any<propto, jacobian, T>(params_r);
  }
  template <bool propto, bool jacobian, typename T>
  T log_prob(std::vector<T>& params_r, std::vector<int>&, std::ostream* = nullptr) const {
    return log_prob_any<propto, jacobian, T>(params_r);
  }

  template <typename VecIn, typename VecOut>
  void constrain(const VecIn& u, VecOut& c) const {
    const size_t P = num_params_r();
    for (size_t i = 0; i < P; ++i)
      c[i] = u[i];
    if (is_ordered()) {   // ordered_constrain.hpp:34-37
      for (int k = 1; k < desc_.n_classes - 1; ++k)
        c[desc_.K + k] = c[desc_.K + k - 1] + std::exp(u[desc_.K + k]);
      return;
    }
    if (is_categorical())
      return;
    if (desc_.G > 0)
      c[1] = std::exp(u[1]);
    if (has_scale(desc_.family))
      c[P - 1] = std::exp(u[P - 1]);
  }
  template <typename VecIn, typename VecOut>
  void unconstrain(const VecIn& c, VecOut& u) const {
    const size_t P = num_params_r();
    for (size_t i = 0; i < P; ++i)
      u[i] = c[i];
    if (is_ordered()) {   // ordered_free
      for (int k = 1; k < desc_.n_classes - 1; ++k)
        u[desc_.K + k] = std::log(c[desc_.K + k] - c[desc_.K + k - 1]);
      return;
    }
    if (is_categorical())
      return;
    if (desc_.G > 0)
      u[1] = stan::math::lb_free(c[1], 0);
    if (has_scale(desc_.family))
      u[P - 1] = stan::math::lb_free(c[P - 1], 0);
  }
  template <typename RNG>
  void write_array(RNG&, Eigen::VectorXd& params_r, Eigen::VectorXd& vars, bool = true, bool = true,
                   std::ostream* = nullptr) const {
    vars.resize(num_params_r());
    constrain(params_r, vars);
  }
  template <typename RNG>
  void write_array(RNG&, std::vector<double>& params_r, std::vector<int>&, std::vector<double>& vars,
                   bool = true, bool = true, std::ostream* = nullptr) const {
    vars.resize(num_params_r());
    constrain(params_r, vars);
  }
  void read_inits(const stan::io::var_context& context, std::vector<double>& c) const {
    std::vector<std::string> names;
    get_param_names(names);
    c.clear();
    for (const auto& nm : names) {
      std::vector<double> v = context.vals_r(nm);
      c.insert(c.end(), v.begin(), v.end());
    }
    if (c.size() != num_params_r())
      throw std::invalid_argument("init context has the wrong number of values");
  }
  void transform_inits(const stan::io::var_context& context, Eigen::VectorXd& params_r,
                       std::ostream* = nullptr) const override {
    std::vector<double> c;
    read_inits(context, c);
    params_r.resize(num_params_r());
    unconstrain(c, params_r);
  }
  void transform_inits(const stan::io::var_context& context, std::vector<int>&, std::vector<double>& params_r,
                       std::ostream* = nullptr) const override {
    std::vector<double> c;
    read_inits(context, c);
    params_r.resize(num_params_r());
    unconstrain(c, params_r);
  }
  void unconstrain_array(const Eigen::VectorXd& c, Eigen::VectorXd& u, std::ostream* = nullptr) const override {
    u.resize(num_params_r());
    unconstrain(c, u);
  }
  void unconstrain_array(const std::vector<double>& c, std::vector<double>& u,
                         std::ostream* = nullptr) const override {
    u.resize(num_params_r());
    unconstrain(c, u);
  }

 private:
  b200glm_desc desc_;
  b200glm_handle* h_;
  unsigned long long uid_;
  std::vector<double> means_x_;
  mutable std::unique_ptr<slot_guard[]> guards_;
  mutable std::atomic<int> next_slot_{0};
  mutable std::atomic<long> n_gradients_{0}, n_leapfrogs_{0}, n_uploads_{0};
};

}  // namespace b200

// ---------------------------------------------------------------------------------------------
// (2) stan::model::gradient -- explicit specialisations (both signatures of gradient.hpp:14-35)
// ---------------------------------------------------------------------------------------------
namespace stan {
namespace model {

template <>
inline void gradient<b200::glm_model>(const b200::glm_model& model, const Eigen::Matrix<double, Eigen::Dynamic, 1>& x,
                                      double& f, Eigen::Matrix<double, Eigen::Dynamic, 1>& grad_f,
                                      std::ostream* /*msgs*/) {
  Eigen::VectorXd g(x.size());
  model.device_log_prob_grad(x.data(), true, true, f, g.data());  // throws before grad_f is touched
  grad_f = std::move(g);
}

template <>
inline void gradient<b200::glm_model>(const b200::glm_model& model, const Eigen::Matrix<double, Eigen::Dynamic, 1>& x,
                                      double& f, Eigen::Matrix<double, Eigen::Dynamic, 1>& grad_f,
                                      callbacks::logger& /*logger*/) {
  gradient<b200::glm_model>(model, x, f, grad_f, static_cast<std::ostream*>(nullptr));
}

}  // namespace model

// ---------------------------------------------------------------------------------------------
// (3) the integrator slot: device-resident leapfrog
// ---------------------------------------------------------------------------------------------
namespace mcmc {

template <>
class expl_leapfrog<diag_e_metric<b200::glm_model, stan::rng_t>>
    : public base_leapfrog<diag_e_metric<b200::glm_model, stan::rng_t>> {
 public:
  using hamiltonian_t = diag_e_metric<b200::glm_model, stan::rng_t>;
  using point_t = typename hamiltonian_t::PointType;

  expl_leapfrog() : base_leapfrog<hamiltonian_t>() {}

  // The model the Hamiltonian was built on.  base_hamiltonian keeps it as a protected reference (model_,
  // base_hamiltonian.hpp:81) and offers no accessor; a derived type with no members of its own reads it.  This
  // replaces round 1's thread-local "last model evaluated on this thread", which depended on hamiltonian.init
  // having run on the same thread before the first evolve.
  struct model_peek : hamiltonian_t {
    static const b200::glm_model& get(hamiltonian_t& h) { return static_cast<model_peek&>(h).model_; }
  };

  // one launch: p -= eps/2 g; q += eps M^-1 p; (V,g) = -(lp, grad lp)(q); p -= eps/2 g
  void evolve(point_t& z, hamiltonian_t& hamiltonian, const double epsilon, callbacks::logger& logger) {
    model_peek::get(hamiltonian).device_leapfrog(z.q, z.p, z.g, z.V, z.inv_e_metric_, epsilon, logger);
  }

  // host sub-steps, kept so the class still satisfies base_leapfrog's interface
  void begin_update_p(point_t& z, hamiltonian_t& hamiltonian, double epsilon, callbacks::logger& logger) {
    z.p -= epsilon * hamiltonian.dphi_dq(z, logger);
  }
  void update_q(point_t& z, hamiltonian_t& hamiltonian, double epsilon, callbacks::logger& logger) {
    z.q += epsilon * hamiltonian.dtau_dp(z);
    hamiltonian.update_potential_gradient(z, logger);
  }
  void end_update_p(point_t& z, hamiltonian_t& hamiltonian, double epsilon, callbacks::logger& logger) {
    z.p -= epsilon * hamiltonian.dphi_dq(z, logger);
  }
};

}  // namespace mcmc
}  // namespace stan

#endif
