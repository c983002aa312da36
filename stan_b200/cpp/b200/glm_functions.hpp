// b200/glm_functions.hpp -- function-level binding: the GLM densities as stan::math overloads.
//
// This is the slot the reference's own GPU backend uses: an overload of the density selected by the
// argument TYPE (matrix_cl there, b200 device views here) that returns a var carrying the
// device-computed partials -- SM/opencl/prim/bernoulli_logit_glm_lpmf.hpp:52-58 (overload),
// :133-138 (partials -> ops_partials.build(logp)); poisson_log_glm_lpmf.hpp, normal_id_glm_lpdf.hpp likewise.
// A hand-written or stanc-generated model keeps its own priors / transforms / other likelihood terms on
// the CPU tape and only swaps the data arguments of the GLM call:
//
//   b200::glm_data d(B200GLM_BERNOULLI_LOGIT, N, K, X.data(), y.data());      // uploads once
//   lp += stan::math::bernoulli_logit_glm_lpmf<propto>(d.y(), d.x(), alpha, beta);   // alpha, beta: var or double
//   lp += stan::math::poisson_log_glm_lpmf<propto>(d.y(), d.x(), b200::by_group(a), beta);   // alpha = a[group]
//   lp += stan::math::binomial_logit_glm_lpmf<propto>(d.y(), d.trials(), d.x(), alpha, beta);
//   lp += stan::math::neg_binomial_2_log_glm_lpmf<propto>(d.y(), d.x(), alpha, beta, phi);   // phi: var or double
//
// Each call is one fused single-pass launch (b200glm_glm_lpmf); the result enters the tape through
// stan::math::precomputed_gradients.  Errors map as in stan_glm_model.hpp (DOMAIN -> std::domain_error ...).
#ifndef B200_GLM_FUNCTIONS_HPP
#define B200_GLM_FUNCTIONS_HPP

#include <stan/math.hpp>

#include <b200glm.h>

#include <atomic>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace b200 {

class glm_data;
struct y_view {
  const glm_data* d;
};
struct x_view {
  const glm_data* d;
};
struct trials_view {   // binomial_logit: the population sizes held next to y
  const glm_data* d;
};
// intercept given per group: alpha_i = a[group_i] with the group index the glm_data holds
// (the model-level construct stan::model::rvalue(a, index_multi(group)), ST/model/indexing/rvalue.hpp:154-172)
template <typename Vec>
struct grouped_intercept {
  const Vec& a;
};
template <typename Vec>
grouped_intercept<Vec> by_group(const Vec& a) {
  return {a};
}

// intercept given PER ROW: alpha is an N-vector (the vector-alpha form of the reference's overloads,
// SM/opencl/prim/normal_id_glm_lpdf.hpp:68-84 is_alpha_vector)
template <typename Vec>
struct row_intercept {
  const Vec& a;
};
template <typename Vec>
row_intercept<Vec> by_row(const Vec& a) {
  return {a};
}

// (y, X [, group]) resident on the device in the panel format the kernels stream
class glm_data {
 public:
  glm_data(int family, long long N, int K, const double* X, const void* y, const int* group = nullptr, int G = 0,
           int device = 0, int n_slots = 4, const int* trials = nullptr) {
    b200glm_desc d = {};
    d.family = family;
    d.N = N;
    d.K = K;
    d.X = X;
    d.ldx = N > 0 ? N : 1;
    if (family == B200GLM_NORMAL_ID)
      d.y_real = static_cast<const double*>(y);
    else
      d.y_int = static_cast<const int32_t*>(y);
    d.group = group;
    d.G = G;
    d.trials = trials;
    d.prior_alpha_sd = d.prior_beta_sd = d.prior_sigma_scale = d.prior_sigma_a_scale = 1.0;  // unused: no priors here
    d.device = device;
    d.n_slots = n_slots;
    d.world = 1;
    family_ = family;
    N_ = N;
    K_ = K;
    G_ = G;
    n_slots_ = n_slots;
    const int rc = b200glm_create(&d, &h_);
    if (rc != B200GLM_OK) {
      const std::string msg = h_ ? b200glm_last_error(h_) : "b200glm_create failed";
      if (h_)
        b200glm_destroy(h_);
      h_ = nullptr;
      raise(rc, msg);
    }
  }
  glm_data(const glm_data&) = delete;
  glm_data& operator=(const glm_data&) = delete;
  ~glm_data() {
    if (h_)
      b200glm_destroy(h_);
  }
  long long N() const { return N_; }
  y_view y() const { return {this}; }
  x_view x() const { return {this}; }
  trials_view trials() const { return {this}; }
  int family() const { return family_; }
  int K() const { return K_; }
  int G() const { return G_; }
  b200glm_handle* handle() const { return h_; }

  [[noreturn]] static void raise(int rc, const std::string& msg) {
    if (rc == B200GLM_DOMAIN)
      throw std::domain_error(msg);
    if (rc == B200GLM_INVALID)
      throw std::invalid_argument(msg);
    throw std::runtime_error("b200glm: " + msg);
  }

  // value + partials; alpha has 1 (G == 0) or G entries
  void evaluate(bool propto, bool operands_are_var, int sigma_is_var, const double* alpha, const double* beta,
                double sigma, double& logp, double* d_alpha, double* d_beta, double* d_sigma) const {
    // evaluations are stateless and serialised per slot inside the library, so any mapping of
    // threads to slots is safe; a per-thread ordinal spreads concurrent callers over the slots
    static std::atomic<int> next_thread{0};
    static thread_local const int thread_ordinal = next_thread.fetch_add(1);
    const int slot = thread_ordinal % (n_slots_ > 0 ? n_slots_ : 1);
    const int rc = b200glm_glm_lpmf(h_, slot, propto, operands_are_var, sigma_is_var, alpha, beta, sigma, &logp,
                                    d_alpha, d_beta, d_sigma);
    if (rc != B200GLM_OK)
      raise(rc, b200glm_last_error(h_));
  }

 private:
  b200glm_handle* h_ = nullptr;
  int family_ = 0, K_ = 0, G_ = 0, n_slots_ = 1;
  long long N_ = 0;

 public:
  // per-row operands (b200glm_glm_lpmf_rows): alpha_rows / sigma_rows may be NULL
  void evaluate_rows(bool propto, bool operands_are_var, int sigma_is_var, const double* alpha_rows, double alpha,
                     const double* beta, const double* sigma_rows, double sigma, double& logp, double* d_alpha_rows,
                     double* d_alpha, double* d_beta, double* d_sigma_rows, double* d_sigma) const {
    static std::atomic<int> next_thread{0};
    static thread_local const int thread_ordinal = next_thread.fetch_add(1);
    const int slot = thread_ordinal % (n_slots_ > 0 ? n_slots_ : 1);
    const int rc = b200glm_glm_lpmf_rows(h_, slot, propto, operands_are_var, sigma_is_var, alpha_rows, alpha, beta,
                                         sigma_rows, sigma, &logp, d_alpha_rows, d_alpha, d_beta, d_sigma_rows, d_sigma);
    if (rc != B200GLM_OK)
      raise(rc, b200glm_last_error(h_));
  }
};

namespace internal {

template <typename T>
inline void push_scalar(const T& x, std::vector<double>& vals, std::vector<stan::math::var>& ops) {
  vals.push_back(stan::math::value_of(x));
  if constexpr (std::is_same<T, stan::math::var>::value)
    ops.push_back(x);
}
template <typename Vec>
inline void push_vector(const Vec& v, std::vector<double>& vals, std::vector<stan::math::var>& ops) {
  for (Eigen::Index i = 0; i < v.size(); ++i)
    push_scalar(v.coeff(i), vals, ops);
}

// common body of all overloads
template <bool propto, typename AlphaPush, typename T_alpha, typename T_beta, typename T_sigma>
stan::return_type_t<T_alpha, T_beta, T_sigma> glm_call(const char* function, int family, const y_view& y,
                                                         const x_view& x, int n_alpha, AlphaPush&& push_alpha,
                                                         const T_beta& beta, const T_sigma& sigma, T_alpha*) {
  using stan::math::var;
  using T_ret = stan::return_type_t<T_alpha, T_beta, T_sigma>;
  if (y.d != x.d || y.d == nullptr)
    throw std::invalid_argument(std::string(function) + ": y and x must be views of the same b200::glm_data");
  const glm_data& d = *y.d;
  if (d.family() != family)
    throw std::invalid_argument(std::string(function) + ": the glm_data was built for another family");
  if (static_cast<int>(beta.size()) != d.K())
    throw std::invalid_argument(std::string(function) + ": Weight vector has the wrong size");   // check_consistent_size
  if (n_alpha != (d.G() > 0 ? d.G() : 1))
    throw std::invalid_argument(std::string(function) + ": Vector of intercepts has the wrong size");
  constexpr bool alpha_var = std::is_same<stan::scalar_type_t<T_alpha>, var>::value;
  constexpr bool beta_var = std::is_same<stan::scalar_type_t<T_beta>, var>::value;
  constexpr bool sigma_var = std::is_same<T_sigma, var>::value;
  constexpr bool any_var = alpha_var || beta_var || sigma_var;
  std::vector<double> va, vb, vs;
  std::vector<var> ops;
  push_alpha(va, ops);
  push_vector(beta, vb, ops);
  push_scalar(sigma, vs, ops);
  double logp = 0, ds = 0;
  std::vector<double> da(va.size()), db(vb.size() + 1);
  // sigma_is_var = 2: the scale / precision is the only autodiff operand (the terms that involve only alpha and
  // beta drop under propto: neg_binomial_2_log_glm_lpmf.hpp:188-190)
  d.evaluate(propto, any_var, sigma_var ? ((alpha_var || beta_var) ? 1 : 2) : 0, va.data(), vb.data(), vs[0], logp,
             da.data(), db.data(), &ds);
  if constexpr (!any_var) {
    return logp;
  } else {
    std::vector<double> grads;
    if (alpha_var)
      grads.insert(grads.end(), da.begin(), da.end());
    if (beta_var)
      grads.insert(grads.end(), db.begin(), db.begin() + vb.size());
    if (sigma_var)
      grads.push_back(ds);
    return T_ret(stan::math::precomputed_gradients(logp, ops, grads));
  }
}

template <typename T>
using is_scalar_operand = std::integral_constant<bool, std::is_arithmetic<T>::value
                                                         || std::is_same<T, stan::math::var>::value>;

// common body of the overloads with per-row operands: T_alpha / T_sigma are scalars or Eigen vectors of size N
template <bool propto, typename T_alpha, typename T_beta, typename T_sigma>
stan::return_type_t<T_alpha, T_beta, T_sigma> glm_call_rows(const char* function, int family, const y_view& y,
                                                              const x_view& x, const T_alpha& alpha, const T_beta& beta,
                                                              const T_sigma& sigma) {
  using stan::math::var;
  using T_ret = stan::return_type_t<T_alpha, T_beta, T_sigma>;
  if (y.d != x.d || y.d == nullptr)
    throw std::invalid_argument(std::string(function) + ": y and x must be views of the same b200::glm_data");
  const glm_data& d = *y.d;
  if (d.family() != family)
    throw std::invalid_argument(std::string(function) + ": the glm_data was built for another family");
  if (static_cast<int>(beta.size()) != d.K())
    throw std::invalid_argument(std::string(function) + ": Weight vector has the wrong size");
  constexpr bool alpha_vec = !is_scalar_operand<T_alpha>::value, sigma_vec = !is_scalar_operand<T_sigma>::value;
  constexpr bool alpha_var = std::is_same<stan::scalar_type_t<T_alpha>, var>::value;
  constexpr bool beta_var = std::is_same<stan::scalar_type_t<T_beta>, var>::value;
  constexpr bool sigma_var = std::is_same<stan::scalar_type_t<T_sigma>, var>::value;
  constexpr bool any_var = alpha_var || beta_var || sigma_var;
  std::vector<double> va, vb, vs;
  std::vector<var> ops;
  if constexpr (alpha_vec) {
    if (static_cast<long long>(alpha.size()) != d.N())   // check_size_match("Rows of x", N, "size of alpha")
      throw std::invalid_argument(std::string(function) + ": Vector of intercepts has the wrong size");
    push_vector(alpha, va, ops);
  } else {
    push_scalar(alpha, va, ops);
  }
  push_vector(beta, vb, ops);
  if constexpr (sigma_vec) {
    if (static_cast<long long>(sigma.size()) != d.N())
      throw std::invalid_argument(std::string(function) + ": Scale vector has the wrong size");
    push_vector(sigma, vs, ops);
  } else {
    push_scalar(sigma, vs, ops);
  }
  double logp = 0, da0 = 0, ds0 = 0;
  std::vector<double> da(alpha_vec ? va.size() : 0), db(vb.size() + 1), ds(sigma_vec ? vs.size() : 0);
  d.evaluate_rows(propto, any_var, sigma_var ? ((alpha_var || beta_var) ? 1 : 2) : 0, alpha_vec ? va.data() : nullptr,
                  va.empty() ? 0.0 : va[0], vb.data(), sigma_vec ? vs.data() : nullptr, vs.empty() ? 1.0 : vs[0], logp,
                  alpha_vec ? da.data() : nullptr, &da0, db.data(), sigma_vec ? ds.data() : nullptr, &ds0);
  if constexpr (!any_var) {
    return logp;
  } else {
    std::vector<double> grads;
    if (alpha_var) {
      if (alpha_vec)
        grads.insert(grads.end(), da.begin(), da.end());
      else
        grads.push_back(da0);
    }
    if (beta_var)
      grads.insert(grads.end(), db.begin(), db.begin() + vb.size());
    if (sigma_var) {
      if (sigma_vec)
        grads.insert(grads.end(), ds.begin(), ds.end());
      else
        grads.push_back(ds0);
    }
    return T_ret(stan::math::precomputed_gradients(logp, ops, grads));
  }
}

}  // namespace internal
}  // namespace b200

namespace stan {
namespace math {

// ---- scalar intercept -------------------------------------------------------------------------
template <bool propto = false, typename T_alpha, typename T_beta,
          std::enable_if_t<b200::internal::is_scalar_operand<T_alpha>::value>* = nullptr>
return_type_t<T_alpha, T_beta> bernoulli_logit_glm_lpmf(const b200::y_view& y, const b200::x_view& x,
                                                         const T_alpha& alpha, const T_beta& beta) {
  return b200::internal::glm_call<propto>(
      "bernoulli_logit_glm_lpmf", B200GLM_BERNOULLI_LOGIT, y, x, 1,
      [&](std::vector<double>& v, std::vector<var>& o) { b200::internal::push_scalar(alpha, v, o); }, beta, 1.0,
      static_cast<T_alpha*>(nullptr));
}
template <bool propto = false, typename T_alpha, typename T_beta,
          std::enable_if_t<b200::internal::is_scalar_operand<T_alpha>::value>* = nullptr>
return_type_t<T_alpha, T_beta> poisson_log_glm_lpmf(const b200::y_view& y, const b200::x_view& x, const T_alpha& alpha,
                                                     const T_beta& beta) {
  return b200::internal::glm_call<propto>(
      "poisson_log_glm_lpmf", B200GLM_POISSON_LOG, y, x, 1,
      [&](std::vector<double>& v, std::vector<var>& o) { b200::internal::push_scalar(alpha, v, o); }, beta, 1.0,
      static_cast<T_alpha*>(nullptr));
}
template <bool propto = false, typename T_alpha, typename T_beta, typename T_sigma,
          std::enable_if_t<b200::internal::is_scalar_operand<T_alpha>::value
                           && b200::internal::is_scalar_operand<T_sigma>::value>* = nullptr>
return_type_t<T_alpha, T_beta, T_sigma> normal_id_glm_lpdf(const b200::y_view& y, const b200::x_view& x,
                                                            const T_alpha& alpha, const T_beta& beta,
                                                            const T_sigma& sigma) {
  return b200::internal::glm_call<propto>(
      "normal_id_glm_lpdf", B200GLM_NORMAL_ID, y, x, 1,
      [&](std::vector<double>& v, std::vector<var>& o) { b200::internal::push_scalar(alpha, v, o); }, beta, sigma,
      static_cast<T_alpha*>(nullptr));
}

// binomial_logit_glm_lpmf(n | N, x, alpha, beta)  (SM/prim/prob/binomial_logit_glm_lpmf.hpp:55-161,
// OpenCL counterpart SM/opencl/prim/binomial_logit_glm_lpmf.hpp)
template <bool propto = false, typename T_alpha, typename T_beta,
          std::enable_if_t<b200::internal::is_scalar_operand<T_alpha>::value>* = nullptr>
return_type_t<T_alpha, T_beta> binomial_logit_glm_lpmf(const b200::y_view& n, const b200::trials_view& N,
                                                        const b200::x_view& x, const T_alpha& alpha,
                                                        const T_beta& beta) {
  if (N.d != n.d)
    throw std::invalid_argument("binomial_logit_glm_lpmf: n and N must be views of the same b200::glm_data");
  return b200::internal::glm_call<propto>(
      "binomial_logit_glm_lpmf", B200GLM_BINOMIAL_LOGIT, n, x, 1,
      [&](std::vector<double>& v, std::vector<var>& o) { b200::internal::push_scalar(alpha, v, o); }, beta, 1.0,
      static_cast<T_alpha*>(nullptr));
}
// neg_binomial_2_log_glm_lpmf(y | x, alpha, beta, phi)  (SM/prim/prob/neg_binomial_2_log_glm_lpmf.hpp:64-249)
template <bool propto = false, typename T_alpha, typename T_beta, typename T_phi,
          std::enable_if_t<b200::internal::is_scalar_operand<T_alpha>::value>* = nullptr>
return_type_t<T_alpha, T_beta, T_phi> neg_binomial_2_log_glm_lpmf(const b200::y_view& y, const b200::x_view& x,
                                                                   const T_alpha& alpha, const T_beta& beta,
                                                                   const T_phi& phi) {
  return b200::internal::glm_call<propto>(
      "neg_binomial_2_log_glm_lpmf", B200GLM_NEG_BINOMIAL_2_LOG, y, x, 1,
      [&](std::vector<double>& v, std::vector<var>& o) { b200::internal::push_scalar(alpha, v, o); }, beta, phi,
      static_cast<T_alpha*>(nullptr));
}

// ---- intercept by group: alpha = a[group] ---------------------------------------------------------
template <bool propto = false, typename Vec, typename T_beta>
return_type_t<Vec, T_beta> binomial_logit_glm_lpmf(const b200::y_view& n, const b200::trials_view& N,
                                                    const b200::x_view& x, const b200::grouped_intercept<Vec>& alpha,
                                                    const T_beta& beta) {
  if (N.d != n.d)
    throw std::invalid_argument("binomial_logit_glm_lpmf: n and N must be views of the same b200::glm_data");
  return b200::internal::glm_call<propto>(
      "binomial_logit_glm_lpmf", B200GLM_BINOMIAL_LOGIT, n, x, static_cast<int>(alpha.a.size()),
      [&](std::vector<double>& v, std::vector<var>& o) { b200::internal::push_vector(alpha.a, v, o); }, beta, 1.0,
      static_cast<Vec*>(nullptr));
}
template <bool propto = false, typename Vec, typename T_beta, typename T_phi>
return_type_t<Vec, T_beta, T_phi> neg_binomial_2_log_glm_lpmf(const b200::y_view& y, const b200::x_view& x,
                                                               const b200::grouped_intercept<Vec>& alpha,
                                                               const T_beta& beta, const T_phi& phi) {
  return b200::internal::glm_call<propto>(
      "neg_binomial_2_log_glm_lpmf", B200GLM_NEG_BINOMIAL_2_LOG, y, x, static_cast<int>(alpha.a.size()),
      [&](std::vector<double>& v, std::vector<var>& o) { b200::internal::push_vector(alpha.a, v, o); }, beta, phi,
      static_cast<Vec*>(nullptr));
}
template <bool propto = false, typename Vec, typename T_beta>
return_type_t<Vec, T_beta> bernoulli_logit_glm_lpmf(const b200::y_view& y, const b200::x_view& x,
                                                     const b200::grouped_intercept<Vec>& alpha, const T_beta& beta) {
  return b200::internal::glm_call<propto>(
      "bernoulli_logit_glm_lpmf", B200GLM_BERNOULLI_LOGIT, y, x, static_cast<int>(alpha.a.size()),
      [&](std::vector<double>& v, std::vector<var>& o) { b200::internal::push_vector(alpha.a, v, o); }, beta, 1.0,
      static_cast<Vec*>(nullptr));
}
template <bool propto = false, typename Vec, typename T_beta>
return_type_t<Vec, T_beta> poisson_log_glm_lpmf(const b200::y_view& y, const b200::x_view& x,
                                                 const b200::grouped_intercept<Vec>& alpha, const T_beta& beta) {
  return b200::internal::glm_call<propto>(
      "poisson_log_glm_lpmf", B200GLM_POISSON_LOG, y, x, static_cast<int>(alpha.a.size()),
      [&](std::vector<double>& v, std::vector<var>& o) { b200::internal::push_vector(alpha.a, v, o); }, beta, 1.0,
      static_cast<Vec*>(nullptr));
}
template <bool propto = false, typename Vec, typename T_beta, typename T_sigma>
return_type_t<Vec, T_beta, T_sigma> normal_id_glm_lpdf(const b200::y_view& y, const b200::x_view& x,
                                                        const b200::grouped_intercept<Vec>& alpha, const T_beta& beta,
                                                        const T_sigma& sigma) {
  return b200::internal::glm_call<propto>(
      "normal_id_glm_lpdf", B200GLM_NORMAL_ID, y, x, static_cast<int>(alpha.a.size()),
      [&](std::vector<double>& v, std::vector<var>& o) { b200::internal::push_vector(alpha.a, v, o); }, beta, sigma,
      static_cast<Vec*>(nullptr));
}

// ---- intercept per row: alpha = b200::by_row(a), a an N-vector; normal_id also with a vector scale ----------------
template <bool propto = false, typename Vec, typename T_beta>
return_type_t<Vec, T_beta> bernoulli_logit_glm_lpmf(const b200::y_view& y, const b200::x_view& x,
                                                     const b200::row_intercept<Vec>& alpha, const T_beta& beta) {
  return b200::internal::glm_call_rows<propto>("bernoulli_logit_glm_lpmf", B200GLM_BERNOULLI_LOGIT, y, x, alpha.a, beta, 1.0);
}
template <bool propto = false, typename Vec, typename T_beta>
return_type_t<Vec, T_beta> poisson_log_glm_lpmf(const b200::y_view& y, const b200::x_view& x,
                                                 const b200::row_intercept<Vec>& alpha, const T_beta& beta) {
  return b200::internal::glm_call_rows<propto>("poisson_log_glm_lpmf", B200GLM_POISSON_LOG, y, x, alpha.a, beta, 1.0);
}
template <bool propto = false, typename Vec, typename T_beta>
return_type_t<Vec, T_beta> binomial_logit_glm_lpmf(const b200::y_view& n, const b200::trials_view& N,
                                                    const b200::x_view& x, const b200::row_intercept<Vec>& alpha,
                                                    const T_beta& beta) {
  if (N.d != n.d)
    throw std::invalid_argument("binomial_logit_glm_lpmf: n and N must be views of the same b200::glm_data");
  return b200::internal::glm_call_rows<propto>("binomial_logit_glm_lpmf", B200GLM_BINOMIAL_LOGIT, n, x, alpha.a, beta, 1.0);
}
template <bool propto = false, typename Vec, typename T_beta, typename T_phi>
return_type_t<Vec, T_beta, T_phi> neg_binomial_2_log_glm_lpmf(const b200::y_view& y, const b200::x_view& x,
                                                               const b200::row_intercept<Vec>& alpha,
                                                               const T_beta& beta, const T_phi& phi) {
  return b200::internal::glm_call_rows<propto>("neg_binomial_2_log_glm_lpmf", B200GLM_NEG_BINOMIAL_2_LOG, y, x, alpha.a,
                                               beta, phi);
}
// sigma: a scalar or an N-vector (Eigen)
template <bool propto = false, typename Vec, typename T_beta, typename T_sigma>
return_type_t<Vec, T_beta, T_sigma> normal_id_glm_lpdf(const b200::y_view& y, const b200::x_view& x,
                                                        const b200::row_intercept<Vec>& alpha, const T_beta& beta,
                                                        const T_sigma& sigma) {
  return b200::internal::glm_call_rows<propto>("normal_id_glm_lpdf", B200GLM_NORMAL_ID, y, x, alpha.a, beta, sigma);
}
// scalar intercept with a vector scale
template <bool propto = false, typename T_alpha, typename T_beta, typename VecS,
          std::enable_if_t<b200::internal::is_scalar_operand<T_alpha>::value
                           && !b200::internal::is_scalar_operand<VecS>::value>* = nullptr>
return_type_t<T_alpha, T_beta, VecS> normal_id_glm_lpdf(const b200::y_view& y, const b200::x_view& x,
                                                         const T_alpha& alpha, const T_beta& beta, const VecS& sigma) {
  return b200::internal::glm_call_rows<propto>("normal_id_glm_lpdf", B200GLM_NORMAL_ID, y, x, alpha, beta, sigma);
}

}  // namespace math
}  // namespace stan

#endif
