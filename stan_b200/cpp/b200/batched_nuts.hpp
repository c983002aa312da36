// b200/batched_nuts.hpp -- multi-chain batched driver for b200::glm_model.
//
// The reference runs chains as independent tasks, each calling the model gradient on its own
// (ST/services/sample/hmc_nuts_diag_e_adapt.hpp:364-401: tbb::parallel_for over chains, one
// util::run_adaptive_sampler per chain).  Here every chain still runs the reference's UNMODIFIED
// single-chain service (hmc_nuts_diag_e_adapt.hpp:58-117 -> adapt_diag_e_nuts -> base_nuts::transition,
// same RNG stream create_rng(seed, init_chain_id + i) as the multi-chain overload :352), but as a FIBER:
// a re-entrant computation with its own small stack, the chains dealt out to a handful of worker threads (one per
// core, at most 16 -- not one OS thread per chain).  Whenever a chain needs a
// leapfrog step (expl_leapfrog::evolve, base_nuts.hpp:254) or a gradient (hamiltonian.init, base_nuts.hpp:85)
// it records the request and switches back to the scheduler -- in the middle of the reference's recursive
// build_tree (base_nuts.hpp:247-352), whose frames simply stay on the fiber's stack; once every live chain has
// a request pending, ONE b200glm_leapfrog_batched call -- one pass over X, the fp64 DMMA GEMM pair -- serves
// them all, and the workers resume their chains one after the other.  (Round 1 gave every chain an OS thread and
// met at a mutex + condition variable: 1024 threads, 1024 wake-ups per batch, an AD tape per thread.)  Chains
// advance in lock-step by leapfrog call; trees of different depth simply make a chain take part in
// more or fewer batches per transition.  A gradient request is served as a leapfrog lane with eps = 0
// (q stays, V and g are refreshed), so a batch never needs two passes.
#ifndef B200_BATCHED_NUTS_HPP
#define B200_BATCHED_NUTS_HPP

#include <b200/fiber.hpp>
#include <b200/stan_glm_model.hpp>

#include <stan/services/sample/hmc_nuts_diag_e_adapt.hpp>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <limits>
#include <memory>
#include <mutex>
#include <thread>

namespace b200 {

class chain_batcher final : public glm_model::batch_hook {
 public:
  chain_batcher(const glm_model& m, int n_chains)
      : m_(m), P_(m.num_params_r()), n_(n_chains), n_active_(n_chains), req_(n_chains), res_(n_chains) {
    // shapes the DMMA kernel does not cover (K > 208, group intercepts) are still served in lock-step,
    // lane by lane, by the single-chain kernels
    const int rc = b200glm_batch_reserve(m_.handle(), n_chains);
    if (rc == B200GLM_INVALID)
      dmma_ok_ = false;
    else
      m_.check(rc);
    const size_t np = static_cast<size_t>(n_) * P_;
    for (auto* v : {&uq_, &up_, &ug_, &uim_, &oq_, &op_, &og_})
      v->resize(np);
    uV_.resize(n_);
    oV_.resize(n_);
    eps_.resize(n_);
    lanes_.resize(n_);
    up_lanes_.resize(n_);
    status_.resize(n_);
    for (auto& r : res_) {
      r.q.assign(P_, 0.0);
      r.p.assign(P_, 0.0);
      r.g.assign(P_, 0.0);
      r.im.assign(P_, 1.0);
    }
  }

  void gradient(int chain, const double* theta, double& lp, double* grad) override {
    request& r = req_[chain];
    r = request();
    r.kind = GRAD;
    r.theta = theta;
    submit(chain);
    if (r.status != B200GLM_OK)
      throw std::domain_error("non-finite log density or gradient (parameters, intercept or X*beta not finite)");
    lp = -r.V_out;
    const double* g = og_.data() + static_cast<size_t>(r.lane) * P_;
    for (size_t k = 0; k < P_; ++k)
      grad[k] = -g[k];
  }

  void leapfrog(int chain, Eigen::VectorXd& q, Eigen::VectorXd& p, Eigen::VectorXd& g, double& V,
                const Eigen::VectorXd& inv_metric, double epsilon, stan::callbacks::logger& logger) override {
    request& r = req_[chain];
    r = request();
    r.kind = LEAPFROG;
    r.q = q.data();
    r.p = p.data();
    r.g = g.data();
    r.V_in = V;
    r.im = inv_metric.data();
    r.eps = epsilon;
    submit(chain);
    const size_t o = static_cast<size_t>(r.lane) * P_, bytes = P_ * sizeof(double);
    std::memcpy(q.data(), oq_.data() + o, bytes);
    std::memcpy(p.data(), op_.data() + o, bytes);
    std::memcpy(g.data(), og_.data() + o, bytes);
    V = r.V_out;
    if (r.status != B200GLM_OK || V == std::numeric_limits<double>::infinity())
      glm_model::reject_message("non-finite log density or gradient", logger);
  }

  // a chain has finished (or failed): the others no longer wait for it
  void leave(int) override { --n_active_; }

  // scheduler side (hmc_nuts_diag_e_adapt_batched): every live chain has a request pending -> one batched launch
  int n_waiting() const { return n_waiting_; }
  void serve() {
    try {
      dispatch();
    } catch (const std::exception& e) {
      fatal_ = e.what();
    }
    for (auto& r : req_)
      r.pending = false;
    n_waiting_ = 0;
  }

  static constexpr int kSingleLaneThreshold = 2;
  long n_batches() const { return n_batches_; }
  long n_single() const { return n_single_; }
  // batches by size: lanes <= 2, 16, 32, 64, 128, 256, 512, more
  static constexpr int kHistBuckets = 8;
  long hist(int b) const { return hist_[b]; }
  long n_lanes() const { return n_lanes_; }

 private:
  enum { GRAD = 0, LEAPFROG = 1 };
  struct request {
    int kind = GRAD;
    bool pending = false;
    const double* theta = nullptr;
    double *q = nullptr, *p = nullptr, *g = nullptr;
    const double* im = nullptr;
    double V_in = 0, eps = 0, V_out = 0;
    int lane = -1, status = 0;
  };
  struct resident {  // what the device holds for a chain slot (as last written / read back)
    bool valid = false;
    std::vector<double> q, p, g, im;
  };

  // chain side: record the request and hand control back to the scheduler; when the chain is resumed its
  // lane of the batch has been computed
  void submit(int chain) {
    req_[chain].pending = true;
    ++n_waiting_;
    fiber::yield();
    if (!fatal_.empty())
      throw std::runtime_error(fatal_);
  }

  void dispatch() {
    const size_t bytes = P_ * sizeof(double);
    int n = 0, n_up = 0;
    for (int c = 0; c < n_; ++c) {
      request& r = req_[c];
      if (!r.pending)
        continue;
      resident& rs = res_[c];
      r.lane = n;
      lanes_[n] = c;
      bool upload;
      if (r.kind == GRAD) {
        eps_[n] = 0.0;
        upload = true;
      } else {
        eps_[n] = r.eps;
        upload = !(rs.valid && std::memcmp(rs.q.data(), r.q, bytes) == 0 && std::memcmp(rs.p.data(), r.p, bytes) == 0
                   && std::memcmp(rs.g.data(), r.g, bytes) == 0 && std::memcmp(rs.im.data(), r.im, bytes) == 0);
      }
      if (upload) {
        const size_t o = static_cast<size_t>(n_up) * P_;
        if (r.kind == GRAD) {
          std::memcpy(uq_.data() + o, r.theta, bytes);
          std::memset(up_.data() + o, 0, bytes);
          std::memset(ug_.data() + o, 0, bytes);
          uV_[n_up] = 0.0;
        } else {
          std::memcpy(uq_.data() + o, r.q, bytes);
          std::memcpy(up_.data() + o, r.p, bytes);
          std::memcpy(ug_.data() + o, r.g, bytes);
          uV_[n_up] = r.V_in;
          std::memcpy(rs.im.data(), r.im, bytes);
        }
        std::memcpy(uim_.data() + o, rs.im.data(), bytes);
        up_lanes_[n_up++] = c;
      }
      ++n;
    }
    if (n == 0)
      return;
    {
      static constexpr int edge[kHistBuckets - 1] = {2, 16, 32, 64, 128, 256, 512};
      int b = 0;
      while (b < kHistBuckets - 1 && n > edge[b])
        ++b;
      ++hist_[b];
    }
    b200glm_handle* h = m_.handle();
    if (!dmma_ok_ || n <= kSingleLaneThreshold) {
      // a straggler or two: the single-chain kernel (one HBM-bound launch each) beats a 64-chain DMMA block
      for (int i = 0; i < n; ++i) {
        const int c = lanes_[i];
        request& r = req_[c];
        resident& rs = res_[c];
        const size_t o = static_cast<size_t>(i) * P_;
        rs.valid = false;   // the batch slot of this chain is stale from now on
        if (r.kind == GRAD) {
          double lp = 0;
          const int rc = b200glm_log_prob_grad(h, 0, r.theta, 1, 1, &lp, og_.data() + o);
          if (rc != B200GLM_OK && rc != B200GLM_DOMAIN)
            m_.check(rc);
          r.status = rc;
          r.V_out = -lp;
          for (size_t k = 0; k < P_; ++k)
            og_[o + k] = -og_[o + k];
        } else {
          m_.check(b200glm_set_state(h, 0, r.q, r.p, r.g, r.V_in));
          const int rc = b200glm_leapfrog(h, 0, r.eps, r.im, oq_.data() + o, op_.data() + o, og_.data() + o, &oV_[i]);
          if (rc != B200GLM_OK && rc != B200GLM_DOMAIN)
            m_.check(rc);
          r.status = (rc == B200GLM_DOMAIN || oV_[i] == std::numeric_limits<double>::infinity()) ? B200GLM_DOMAIN
                                                                                                  : B200GLM_OK;
          r.V_out = rc == B200GLM_DOMAIN ? std::numeric_limits<double>::infinity() : oV_[i];
        }
      }
      ++n_batches_;
      n_lanes_ += n;
      n_single_ += n;
      return;
    }
    if (n_up > 0) {
      m_.check(b200glm_set_state_batched(h, n_up, up_lanes_.data(), uq_.data(), up_.data(), ug_.data(), uV_.data(),
                                         uim_.data()));
      m_.count_upload(n_up);
    }
    m_.check(b200glm_leapfrog_batched(h, n, lanes_.data(), eps_.data(), oq_.data(), op_.data(), og_.data(),
                                      oV_.data(), status_.data()));
    ++n_batches_;
    n_lanes_ += n;
    for (int i = 0; i < n; ++i) {
      const int c = lanes_[i];
      request& r = req_[c];
      resident& rs = res_[c];
      r.V_out = oV_[i];
      r.status = status_[i];
      const size_t o = static_cast<size_t>(i) * P_;
      std::memcpy(rs.q.data(), oq_.data() + o, bytes);
      std::memcpy(rs.p.data(), op_.data() + o, bytes);
      std::memcpy(rs.g.data(), og_.data() + o, bytes);
      rs.valid = true;
    }
  }

  const glm_model& m_;
  const size_t P_;
  const int n_;
  std::atomic<int> n_active_, n_waiting_{0};   // chains of different workers park / leave concurrently
  std::string fatal_;
  std::vector<request> req_;
  std::vector<resident> res_;
  std::vector<double> uq_, up_, ug_, uim_, uV_, oq_, op_, og_, oV_, eps_;
  std::vector<int32_t> lanes_, up_lanes_, status_;
  long n_batches_ = 0, n_lanes_ = 0, n_single_ = 0;
  long hist_[kHistBuckets] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool dmma_ok_ = true;
};

// Same contract and argument list as the reference's multi-chain
// stan::services::sample::hmc_nuts_diag_e_adapt (hmc_nuts_diag_e_adapt.hpp:331-404); chain i uses
// create_rng(random_seed, init_chain_id + i) exactly as there.  stats (optional, 10 longs) receives
// {batched launches, lanes served}.
template <typename InitContextPtr, typename InitInvContextPtr, typename InitWriter, typename SampleWriter,
          typename DiagnosticWriter, typename MetricWriter>
int hmc_nuts_diag_e_adapt_batched(glm_model& model, size_t num_chains, const std::vector<InitContextPtr>& init,
                                  const std::vector<InitInvContextPtr>& init_inv_metric, unsigned int random_seed,
                                  unsigned int init_chain_id, double init_radius, int num_warmup, int num_samples,
                                  int num_thin, bool save_warmup, int refresh, double stepsize,
                                  double stepsize_jitter, int max_depth, double delta, double gamma, double kappa,
                                  double t0, unsigned int init_buffer, unsigned int term_buffer, unsigned int window,
                                  stan::callbacks::interrupt& interrupt, stan::callbacks::logger& logger,
                                  std::vector<InitWriter>& init_writer, std::vector<SampleWriter>& sample_writer,
                                  std::vector<DiagnosticWriter>& diagnostic_writer,
                                  std::vector<MetricWriter>& metric_writer, long* stats = nullptr) {
  chain_batcher batcher(model, static_cast<int>(num_chains));
  std::vector<int> rc(num_chains, 0);
  std::vector<std::unique_ptr<fiber>> chains;
  chains.reserve(num_chains);
  for (size_t i = 0; i < num_chains; ++i) {
    chains.emplace_back(new fiber([&, i] {
      glm_model::batch_scope scope(&batcher, static_cast<int>(i));
      try {
        rc[i] = stan::services::sample::hmc_nuts_diag_e_adapt(
            model, *init[i], *init_inv_metric[i], random_seed, init_chain_id + i, init_radius, num_warmup,
            num_samples, num_thin, save_warmup, refresh, stepsize, stepsize_jitter, max_depth, delta, gamma, kappa,
            t0, init_buffer, term_buffer, window, interrupt, logger, init_writer[i], sample_writer[i],
            diagnostic_writer[i], metric_writer[i]);
      } catch (const std::exception& e) {
        logger.error(e.what());
        rc[i] = stan::services::error_codes::SOFTWARE;
      }
    }));
  }
  // The scheduler.  The chains are dealt out to a few worker threads (not one thread per chain: as many as the host
  // has cores, at most 16); a worker resumes each of its unfinished chains in turn -- each runs, through the
  // reference's transition / build_tree code, until its next leapfrog or gradient request, or to its end -- then
  // meets the other workers; the last one to arrive serves ALL parked chains with one batched launch, and the next
  // sweep starts.  A fiber is only ever resumed by its own worker (thread-local AD tape, slot binding).  With one
  // worker this is a plain loop; the workers only spread the host-side tree arithmetic (a few P-vector operations and
  // copies per leapfrog and chain) over cores -- at 1024 chains that is ~8 ms per full batch on a single core.
  const int n_workers = static_cast<int>(std::max<size_t>(
      1, std::min<size_t>({num_chains, std::max(1u, std::thread::hardware_concurrency()), size_t(16)})));
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  size_t live_total = 0;
  std::uint64_t gen = 0;
  bool stop = false;
  auto worker = [&](int t) {
    stan::math::ChainableStack tape;   // STAN_THREADS: every thread owns an AD tape (init_chainablestack.hpp)
    for (;;) {
      size_t live = 0;
      for (size_t i = t; i < num_chains; i += n_workers) {
        if (chains[i]->done())
          continue;
        glm_model::tls_hook() = &batcher;          // the fibers of a worker share its thread-locals
        glm_model::tls_chain() = static_cast<int>(i);
        chains[i]->resume();
        if (!chains[i]->done())
          ++live;
      }
      glm_model::tls_hook() = nullptr;
      glm_model::tls_chain() = -1;
      std::unique_lock<std::mutex> lk(mu);
      live_total += live;
      if (++arrived == n_workers) {
        if (live_total > 0)
          batcher.serve();
        stop = live_total == 0;
        live_total = 0;
        arrived = 0;
        ++gen;
        cv.notify_all();
      } else {
        const std::uint64_t g0 = gen;
        cv.wait(lk, [&] { return gen != g0; });
      }
      if (stop)
        break;
    }
  };
  {
    std::vector<std::thread> pool;
    for (int t = 1; t < n_workers; ++t)
      pool.emplace_back(worker, t);
    worker(0);
    for (auto& th : pool)
      th.join();
  }
  if (stats) {
    stats[0] = batcher.n_batches();
    stats[1] = batcher.n_lanes();
    for (int b = 0; b < chain_batcher::kHistBuckets; ++b)
      stats[2 + b] = batcher.hist(b);   // batches with <= 2, 16, 32, 64, 128, 256, 512, more lanes
  }
  for (int r : rc)
    if (r != 0)
      return r;
  return stan::services::error_codes::OK;
}

}  // namespace b200

#endif
