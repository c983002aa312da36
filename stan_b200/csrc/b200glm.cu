// b200glm.cu -- C-ABI implementation over the sm_100a kernels (include/b200glm.h).
//
// Host side of the boundary: owns device buffers (X in row-panel format, per-slot workspaces and
// streams), launches, the NCCL all-reduce of the likelihood partials, and status mapping.
// No CPU fallback: without a usable CUDA device every compute entry point returns B200GLM_CUDA.
#include "../../include/b200glm.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "glm_kernels.cuh"
#include "glm_wide_kernel.cuh"
#include "glm_multi_kernel.cuh"
#include "nuts_kernels.cuh"
#include "glm_batched_kernel.cuh"
#include "glm_class_kernel.cuh"
#include "measure.cuh"

using namespace b200glm;

namespace {

// ---------------------------------------------------------------------------- NCCL (dlopen)
// NCCL is resolved at run time so the library loads on a box without it and shares the copy a
// host process (e.g. torch) already mapped.  Only ncclAllReduce/ncclCommInitRank are used.
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8, ncclSum = 0 };
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) return;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
    api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy;
  });
  return api;
}

struct Slot {
  cudaStream_t stream = nullptr;
  double* theta = nullptr;       // P
  double* state[2] = {nullptr, nullptr};  // 3P+1 each
  int cur = 0;
  double* inv_metric = nullptr;  // P
  double* partials = nullptr;    // grid * pstride
  unsigned int* ticket = nullptr;
  double* r_out = nullptr;       // n_panels*32 (G > 0, unfused group path)
  double* gpart = nullptr;       // grid * Gcs (fused group path)
  double* cuts = nullptr;        // ordered_logistic: 2 (C - 1) doubles of epilogue scratch
  // b200glm_glm_lpmf_rows (allocated on first use): per-row operands in, per-row partials out, N doubles each
  double *alpha_rows = nullptr, *sigma_rows = nullptr, *s_out = nullptr;
  bool use_alpha_rows = false, use_sigma_rows = false;   // set for the duration of one rows evaluation
  double* lik = nullptr;         // P + 2
  double* result = nullptr;      // P + 2
  double* theta_used = nullptr;  // P
  double* h_pinned = nullptr;    // pinned staging: max(3P+1, P+2) * 2
  unsigned long long peer_seq = 0;  // evaluations exchanged through the peer mailboxes so far
  unsigned long long host_seq = 0;  // launches that mirrored their outputs into h_out so far
  double* h_out_dev = nullptr;      // device-side address of h_out
  double* h_out = nullptr;          // pinned: [result P+2][state 3P+1][sequence word], written by the epilogue itself
  unsigned long long* tl = nullptr; // (grid + 1) x 16 time stamps of the last launch (b200glm_timeline_enable)
  std::recursive_mutex mu;          // recursive: b200glm_glm_lpmf_rows holds it around eval_host
};

// Workspace of the batched (many-chain) path: chain state is feature-major [P][ld] on the device.
struct Batch {
  int max_chains = 0, ld = 0;   // ld = max_chains rounded up to 64 (state and lane leading dimension)
  int S = 0, mbh = 0, sms = 0;
  size_t smem = 0;
  int rs_pairs = 0, rs_S = 0;   // row-split variant (<= 16 lanes): active warp pairs, ring stages per pair
  size_t rs_smem = 0;
  int mu_cpl = 0, mu_S = 0;     // few-chain FMA kernel (<= 4 lanes, K <= 128; glm_multi_kernel.cuh): 0 = not available
  size_t mu_smem = 0;
  cudaStream_t stream = nullptr;
  double *Q = nullptr, *Pm = nullptr, *Gd = nullptr, *V = nullptr, *IM = nullptr;
  double *theta_c = nullptr, *p_half = nullptr, *partials = nullptr, *reduced = nullptr, *result = nullptr,
         *state_out = nullptr;
  double *theta_in = nullptr, *eps_d = nullptr;
  int32_t* chains_d = nullptr;
  double* h_pin = nullptr;      // pinned staging, max_chains * (3P + 1) doubles (+ ints)
  int32_t* h_pin_i = nullptr;
  std::mutex mu;
  // device-side NUTS (b200glm_nuts_*): per-chain state machine behind the batched leapfrog
  struct Nuts {
    NutsConfig cfg;
    int n_chains = 0;
    size_t vstride = 0;
    NutsChain* chains = nullptr;     // device
    double* vec = nullptr;           // device [n][vstride]
    double* eps_c = nullptr;         // device [ld]
    double *normals = nullptr, *uniforms = nullptr, *draws = nullptr, *metric = nullptr;   // pinned host
    NutsStatus* status = nullptr;    // pinned host
    int32_t* lanes_d = nullptr;      // device: the lane list of the last round (re-uploaded only when it changes)
    std::vector<int32_t> lanes_h;
    // a round is 6-7 small launches; for small models their launch gaps are the round.  One CUDA graph per lane count
    // (the launches' parameters depend on nothing else: lane list, steps and state are read from device memory)
    std::map<int, std::pair<cudaGraphExec_t, int>> graphs;   // n -> (executable, launches inside)
    bool use_graphs = true;
  };
  Nuts* nuts = nullptr;
};

void free_nuts(Batch* b) {
  if (!b || !b->nuts) return;
  Batch::Nuts* u = b->nuts;
  cudaFree(u->chains);
  cudaFree(u->vec);
  cudaFree(u->eps_c);
  cudaFree(u->lanes_d);
  for (auto& g : u->graphs) cudaGraphExecDestroy(g.second.first);
  for (double* q : {u->normals, u->uniforms, u->draws, u->metric})
    if (q) cudaFreeHost(q);
  if (u->status) cudaFreeHost(u->status);
  delete u;
  b->nuts = nullptr;
}

}  // namespace

struct b200glm_handle {
  b200glm_desc d;
  int P = 0, off_beta = 0, C = 0;
  long long n_panels = 0;
  double* panels = nullptr;
  long long* seg_ptr = nullptr;  // G+1 (device)
  int grid = 0, n_stages = 0, stage_a = 0;
  int group_fused = 0, Gcs = 0;   // fused group path (narrow kernel, G > 0): see KernelParams
  int4* gmeta = nullptr;          // G entries (device)
  int* cta_g0 = nullptr;          // grid entries (device)
  int state_smem = 0;  // the kernels keep the chain state + likelihood sums in shared memory for the epilogue
  size_t smem_bytes = 0;
  int cpl = 0, cmax = 0;   // cmax: class-outcome models, classes kept in registers
  int pstride = 0;         // doubles per CTA partial row
  // wide kernel (K > 256 or B200GLM_FLAG_FORCE_WIDE): 16-row panels streamed as J sub-panels
  bool wide = false;
  int panel_rows = PANEL_ROWS, Cpad = 0, Kc = 0, J = 0, spw = 0, spc = 0;
  double lgamma_sum = 0.0;  // local shard
  double lgamma_sum_total = 0.0;
  bool bad_y = false;        // any shard holds an out-of-range y (after b200glm_comm_init / set_shard_constants_total)
  bool bad_y_local = false;  // this shard does
  // streamed construction (B200GLM_FLAG_STREAMED): rows arrive through b200glm_append_rows and are re-laid out into the
  // panels chunk by chunk, so X never has to be resident a second time next to its panel copy
  bool streamed = false, ready = true;
  long long rows_appended = 0;
  double bad_count = 0.0;
  int pdl_prefetch = 2;     // B200GLM_PDL_PREFETCH=<stages> (A/B runs); see the TMA producer in glm_kernels.cuh
  int wide_producer = -1;   // -1 = chosen at create from the ring depth; B200GLM_WIDE_PRODUCER=single|lanes (A/B runs); see the TMA producer in glm_wide_kernel.cuh
  bool tl_repeat = false;   // B200GLM_TL_REPEAT=1 (timeline runs only): the last CTA sums the partial rows twice
  bool inline_theta = true; // B200GLM_NO_INLINE_THETA=1: always upload theta with a host-to-device copy (A/B runs)
  bool host_mirror = true;  // B200GLM_NO_HOST_MIRROR=1: fetch results with a device-to-host copy + stream sync (A/B runs)
  bool pdl = true;          // B200GLM_NO_PDL=1 in the environment turns programmatic dependent launch off (A/B runs)
  std::vector<Slot*> slots;
  Batch* batch = nullptr;
  ncclComm_t comm = nullptr;
  // peer mailboxes (CUDA IPC): [n_slots][PEER_BUFS][world][peer_stride] doubles on every rank
  double* mbox = nullptr;
  double* peer_mbox[MAX_PEERS] = {nullptr};
  int peer_stride = 0;
  bool peer_on = false;
  std::atomic<long long> launches{0};
  std::mutex err_mu;
  std::string last_error;
  void set_error(const std::string& s) {
    std::lock_guard<std::mutex> g(err_mu);
    last_error = s;
  }
};

namespace {

#define CUDA_TRY(h, expr)                                                                  \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      (h)->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                  \
      return B200GLM_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

const int kCplChoices[] = {2, 4, 7, 13, 16, 25, 32};

typedef void (*kernel_fn)(const KernelParams);

template <int FAMILY>
kernel_fn pick_cpl(int cpl) {
  switch (cpl) {
    case 2: return glm_fused_kernel<FAMILY, 2>;
    case 4: return glm_fused_kernel<FAMILY, 4>;
    case 7: return glm_fused_kernel<FAMILY, 7>;
    case 13: return glm_fused_kernel<FAMILY, 13>;
    case 16: return glm_fused_kernel<FAMILY, 16>;
    case 25: return glm_fused_kernel<FAMILY, 25>;
    case 32: return glm_fused_kernel<FAMILY, 32>;
  }
  return nullptr;
}
kernel_fn pick_kernel(int family, int cpl) {
  switch (family) {
    case FAM_BERNOULLI_LOGIT: return pick_cpl<FAM_BERNOULLI_LOGIT>(cpl);
    case FAM_POISSON_LOG: return pick_cpl<FAM_POISSON_LOG>(cpl);
    case FAM_NORMAL_ID: return pick_cpl<FAM_NORMAL_ID>(cpl);
    case FAM_BINOMIAL_LOGIT: return pick_cpl<FAM_BINOMIAL_LOGIT>(cpl);
    case FAM_NEG_BINOMIAL_2_LOG: return pick_cpl<FAM_NEG_BINOMIAL_2_LOG>(cpl);
  }
  return nullptr;
}

template <int FAMILY>
kernel_fn pick_wide_shape(int wr, int spw, int spc) {
  if (wr == 16 && spw == 1 && spc == 4) return glm_wide_kernel<FAMILY, 16, 1, 4>;
  if (wr == 16 && spw == 1 && spc == 8) return glm_wide_kernel<FAMILY, 16, 1, 8>;
  if (wr == 8 && spw == 1 && spc == 8) return glm_wide_kernel<FAMILY, 8, 1, 8>;
  if (wr == 4 && spw == 1 && spc == 4) return glm_wide_kernel<FAMILY, 4, 1, 4>;
  if (wr == 4 && spw == 1 && spc == 8) return glm_wide_kernel<FAMILY, 4, 1, 8>;
  if (wr == 4 && spw == 2 && spc == 8) return glm_wide_kernel<FAMILY, 4, 2, 8>;
  return nullptr;
}
kernel_fn pick_wide_kernel(int family, int wr, int spw, int spc) {
  switch (family) {
    case FAM_BERNOULLI_LOGIT: return pick_wide_shape<FAM_BERNOULLI_LOGIT>(wr, spw, spc);
    case FAM_POISSON_LOG: return pick_wide_shape<FAM_POISSON_LOG>(wr, spw, spc);
    case FAM_NORMAL_ID: return pick_wide_shape<FAM_NORMAL_ID>(wr, spw, spc);
    case FAM_BINOMIAL_LOGIT: return pick_wide_shape<FAM_BINOMIAL_LOGIT>(wr, spw, spc);
    case FAM_NEG_BINOMIAL_2_LOG: return pick_wide_shape<FAM_NEG_BINOMIAL_2_LOG>(wr, spw, spc);
  }
  return nullptr;
}
// class-outcome models: (columns per lane, classes kept in registers) shapes of glm_class_kernel
struct ClassShape {
  int cpl, cmax;
};
const ClassShape kOrderedShapes[] = {{4, 1}, {13, 1}, {32, 1}};                 // K <= 32, 104, 256
const ClassShape kCategoricalShapes[] = {{25, 2}, {13, 4}, {7, 8}, {2, 16}};    // K <= 200 / 104 / 56 / 16 with <= 2 / 4 / 8 / 16 classes
kernel_fn pick_class_kernel(int family, int cpl, int cmax) {
  if (family == FAM_ORDERED_LOGISTIC) {
    if (cpl == 4) return glm_class_kernel<true, 4, 1>;
    if (cpl == 13) return glm_class_kernel<true, 13, 1>;
    if (cpl == 32) return glm_class_kernel<true, 32, 1>;
  } else {
    if (cpl == 25 && cmax == 2) return glm_class_kernel<false, 25, 2>;
    if (cpl == 13 && cmax == 4) return glm_class_kernel<false, 13, 4>;
    if (cpl == 7 && cmax == 8) return glm_class_kernel<false, 7, 8>;
    if (cpl == 2 && cmax == 16) return glm_class_kernel<false, 2, 16>;
  }
  return nullptr;
}
kernel_fn handle_kernel(const b200glm_handle* h) {
  if (fam_is_class(h->d.family)) return pick_class_kernel(h->d.family, h->cpl, h->cmax);
  return h->wide ? pick_wide_kernel(h->d.family, h->panel_rows, h->spw, h->spc) : pick_kernel(h->d.family, h->cpl);
}

size_t fixed_smem_bytes(int K, int G, int stage_a, int S, int P_state = 0, int Gcs = 0) {
  const int Kpad = (K + 3) & ~3;
  size_t b = 0;
  if (P_state) b += (size_t)state_smem_doubles(P_state) * 8;   // on-chip chain state
  b += (size_t)NUM_CONSUMER_WARPS * Gcs * 8;                   // fused group path: per-warp group sums
  b += (size_t)Kpad * 8;                                 // sbeta
  b += (size_t)NUM_CONSUMER_WARPS * 32 * 8;              // sr
  b += (size_t)NUM_CONSUMER_WARPS * (Kpad + 4) * 8;      // red
  if (stage_a) b += (size_t)((G + 1) & ~1) * 8;          // sa
  b += (size_t)2 * S * 8;                                // barriers
  return b;
}

int validate_slot(b200glm_handle* h, int slot) {
  if (!h) return B200GLM_INVALID;
  if (!h->ready) {
    h->set_error("streamed handle: b200glm_finalize has not been called (or not all rows were appended)");
    return B200GLM_INVALID;
  }
  if (slot < 0 || slot >= (int)h->slots.size()) {
    h->set_error("slot out of range");
    return B200GLM_INVALID;
  }
  return B200GLM_OK;
}

// Rows the likelihood term runs over (all shards).  Reference quirk kept: binomial_logit_glm_lpmf returns 0 for
// an empty weight vector (size_zero(n, N, alpha, beta, x), binomial_logit_glm_lpmf.hpp:77-79).
long long lik_rows_total(const b200glm_handle* h) {
  if (h->d.family == B200GLM_BINOMIAL_LOGIT && h->d.K == 0) return 0;
  // one class: categorical_logit_glm_lpmf.hpp:70-72 returns 0; no cut-points: ordered_logistic_glm_lpmf.hpp:93-95
  if (fam_is_class(h->d.family) && h->d.n_classes <= 1) return 0;
  return h->d.N_total > 0 ? h->d.N_total : h->d.N;
}

void fill_params(b200glm_handle* h, Slot* s, KernelParams& p, int mode, int propto, int jacobian, int is_var,
                 double eps, int lik_only = 0, int sigma_is_var = 0) {
  std::memset(&p, 0, sizeof(p));
  p.panels = h->panels;
  p.n_rows = h->d.N;
  p.n_panels = h->n_panels;
  p.K = h->d.K;
  p.C = h->C;
  p.G = h->d.G;
  p.family = h->d.family;
  p.P = h->P;
  p.off_beta = h->off_beta;
  p.n_stages = h->n_stages;
  p.Cpad = h->Cpad;
  p.Kc = h->Kc;
  p.J = h->J;
  p.mode = mode;
  p.fuse_finish = ((h->d.world <= 1 || h->peer_on) && (h->d.G == 0 || h->group_fused)) ? 1 : 0;
  if (fam_is_class(h->d.family)) p.state_in_smem = 0;
  if (h->peer_on) {
    int slot_idx = 0;
    for (size_t i = 0; i < h->slots.size(); ++i)
      if (h->slots[i] == s) slot_idx = (int)i;
    p.peer.enabled = 1;
    p.peer.world = h->d.world;
    p.peer.rank = h->d.rank;
    p.peer.stride = h->peer_stride;
    p.peer.seq = s->peer_seq;
    p.peer.timeout_ns = 20ull * 1000000000ull;
    for (int r = 0; r < h->d.world; ++r)
      p.peer.mbox[r] = h->peer_mbox[r] + (size_t)slot_idx * PEER_BUFS * h->d.world * h->peer_stride;
  }
  p.stage_a_in_smem = h->stage_a;
  p.state_in_smem = h->state_smem;
  p.theta_in = s->theta;
  p.st_in = s->state[s->cur];
  p.st_out = s->state[s->cur ^ 1];
  p.inv_metric = s->inv_metric;
  p.eps = eps;
  p.partials = s->partials;
  p.pstride = h->pstride;
  p.n_classes = h->d.n_classes;
  p.cuts = s->cuts;
  p.alpha_rows = s->use_alpha_rows ? s->alpha_rows : nullptr;
  p.sigma_rows = s->use_sigma_rows ? s->sigma_rows : nullptr;
  p.s_out = s->s_out;
  p.ticket = s->ticket;
  p.r_out = s->r_out;
  p.group_fused = h->group_fused;
  p.Gcs = h->Gcs;
  p.gpart = s->gpart;
  p.gmeta = h->gmeta;
  p.cta_g0 = h->cta_g0;
  p.lik = s->lik;
  p.result = s->result;
  p.theta_used = s->theta_used;
  p.tl = s->tl;
  p.pdl_prefetch = h->pdl_prefetch;
  p.tl_repeat = (s->tl && h->tl_repeat) ? 1 : 0;
  p.wide_producer = h->wide_producer;
  ModelConst& mc = p.mc;
  mc.family = h->d.family;
  mc.K = h->d.K;
  mc.G = h->d.G;
  mc.P = h->P;
  mc.off_beta = h->off_beta;
  mc.propto = propto;
  mc.jacobian = jacobian;
  mc.is_var = is_var;
  mc.lik_only = lik_only;
  mc.sigma_is_var = sigma_is_var;
  mc.sigma_rows = s->use_sigma_rows ? 1 : 0;
  mc.N_total = (double)lik_rows_total(h);
  mc.lgamma_sum = h->lgamma_sum_total;
  mc.prior_alpha_sd = h->d.prior_alpha_sd;
  mc.prior_beta_sd = h->d.prior_beta_sd;
  mc.prior_sigma_loc = h->d.prior_sigma_loc;
  mc.prior_sigma_scale = h->d.prior_sigma_scale;
  mc.prior_sigma_a_scale = h->d.prior_sigma_a_scale;
}

// Enqueue one evaluation (all launches + the optional all-reduce) on the slot's stream.
// host_theta != NULL (MODE_THETA): theta is still on the host; it travels in the kernel's parameter block when the
// model is small enough and the main kernel runs, else through the slot's pinned staging + a host-to-device copy.
// mirror_to_host: the epilogue also writes result / state into s->h_out and releases its sequence word (wait_host_out).
int enqueue_eval(b200glm_handle* h, Slot* s, int mode, int propto, int jacobian, int is_var, double eps,
                 int lik_only = 0, int sigma_is_var = 0, const double* host_theta = nullptr,
                 bool mirror_to_host = false) {
  // the parameter block is the last thing cudaLaunchKernelEx copies: static storage keeps 2.5 KB off the stack
  static thread_local KernelParams p;
  const bool need_likelihood = ((!propto) || is_var) && h->d.N_total != -1;
  const bool rows_anywhere = lik_rows_total(h) > 0;
  const bool exchange = need_likelihood && rows_anywhere && h->peer_on;
  if (exchange) ++s->peer_seq;
  fill_params(h, s, p, mode, propto, jacobian, is_var, eps, lik_only, sigma_is_var);
  if (host_theta) {
    if (need_likelihood && rows_anywhere && h->P <= THETA_INLINE_MAX && h->inline_theta) {
      p.theta_inline_n = h->P;
      std::memcpy(p.theta_inline, host_theta, sizeof(double) * h->P);
    } else {
      std::memcpy(s->h_pinned, host_theta, sizeof(double) * h->P);
      CUDA_TRY(h, cudaMemcpyAsync(s->theta, s->h_pinned, sizeof(double) * h->P, cudaMemcpyHostToDevice, s->stream));
    }
  }
  if (mirror_to_host) {
    p.host_out = s->h_out_dev;
    p.host_seq = ++s->host_seq;
  }
  const bool sums_in_main = h->d.G == 0 || h->group_fused;   // the likelihood sums are complete when the main kernel ends
  p.peer_in_main = (exchange && sums_in_main) ? 1 : 0;
  p.peer_in_finish = (exchange && !sums_in_main) ? 1 : 0;
  if (need_likelihood && rows_anywhere) {
    // Programmatic dependent launch: this launch's CTAs may take SMs (barrier set-up, first TMA loads of X) while
    // the previous launch on the stream is still in its one-CTA epilogue; the kernels order every read of that
    // launch's results behind griddepcontrol.wait (glm_kernels.cuh).
    kernel_fn fn = handle_kernel(h);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(h->grid);
    cfg.blockDim = dim3(h->wide ? WIDE_THREADS : NUM_THREADS);
    cfg.dynamicSmemBytes = h->smem_bytes;
    cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = h->pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_TRY(h, cudaLaunchKernelEx(&cfg, fn, p));
    h->launches++;
    if (h->d.G > 0 && !h->group_fused) {
      const int gb = std::min(h->d.G, 8 * 148);   // 8 CTAs of 256 threads are resident per SM
      group_reduce_kernel<<<gb, 256, 0, s->stream>>>(s->r_out, h->seg_ptr, h->d.G, s->lik + 2);
      h->launches++;
    }
    if (fam_is_class(h->d.family) && h->d.world > 1 && !h->peer_on) {
      h->set_error("row-sharded class-outcome models need the peer-mailbox exchange (b200glm_peer_connect)");
      return B200GLM_INVALID;
    }
    if (h->d.world > 1 && !h->peer_on) {
      if (!h->comm) {
        h->set_error("world > 1 but neither b200glm_peer_connect nor b200glm_comm_init was called");
        return B200GLM_INVALID;
      }
      ncclResult_t r = nccl().AllReduce(s->lik, s->lik, (size_t)h->P + 2, ncclFloat64, ncclSum, h->comm, s->stream);
      if (r != 0) {
        h->set_error(std::string("ncclAllReduce: ") + (nccl().GetErrorString ? nccl().GetErrorString(r) : "?"));
        return B200GLM_CUDA;
      }
    }
    if (!p.fuse_finish && !fam_is_class(h->d.family)) {
      finish_kernel<<<1, NUM_THREADS, 0, s->stream>>>(p);
      h->launches++;
    }
  } else {
    // nothing data-dependent left (double semantics with propto, or N == 0): epilogue only
    CUDA_TRY(h, cudaMemsetAsync(s->lik, 0, sizeof(double) * (h->P + 2), s->stream));
    if (mode == MODE_LEAPFROG) {
      theta_from_state_kernel<<<1, NUM_THREADS, 0, s->stream>>>(p);
      h->launches++;
    } else {
      CUDA_TRY(h, cudaMemcpyAsync(s->theta_used, s->theta, sizeof(double) * h->P, cudaMemcpyDeviceToDevice, s->stream));
    }
    if (fam_is_class(h->d.family))
      class_finish_kernel<<<1, NUM_THREADS, 0, s->stream>>>(p);
    else
      finish_kernel<<<1, NUM_THREADS, 0, s->stream>>>(p);
    h->launches++;
  }
  CUDA_TRY(h, cudaGetLastError());
  return B200GLM_OK;
}

// Wait for the launch that mirrors into s->h_out: poll its sequence word (written by the epilogue after a
// system-scope fence); every so often ask the stream, so that a failed launch or a path that did not reach the
// mirrored epilogue ends the wait.  Returns true if the mirror is valid, false if the caller must copy from the device.
int wait_host_out(b200glm_handle* h, Slot* s, bool* mirrored) {
  const int P = h->P;
  volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(s->h_out + (P + 2) + (3 * P + 1));
  *mirrored = false;
  for (unsigned spins = 0;; ++spins) {
    if (*flag == s->host_seq) {
      std::atomic_thread_fence(std::memory_order_acquire);
      *mirrored = true;
      return B200GLM_OK;
    }
    if ((spins & 1023u) == 1023u) {
      const cudaError_t q = cudaStreamQuery(s->stream);
      if (q == cudaSuccess) {
        *mirrored = (*flag == s->host_seq);
        return B200GLM_OK;
      }
      if (q != cudaErrorNotReady) {
        h->set_error(std::string("cudaStreamQuery: ") + cudaGetErrorString(q));
        return B200GLM_CUDA;
      }
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
}

int eval_host(b200glm_handle* h, int slot, const double* theta, int propto, int jacobian, int is_var, double* lp,
              double* grad, int lik_only = 0, int sigma_is_var = 0) {
  int rc = validate_slot(h, slot);
  if (rc) return rc;
  if (!theta || !lp) {
    h->set_error("null pointer argument");
    return B200GLM_INVALID;
  }
  Slot* s = h->slots[slot];
  std::lock_guard<std::recursive_mutex> g(s->mu);
  CUDA_TRY(h, cudaSetDevice(h->d.device));
  const int P = h->P;
  rc = enqueue_eval(h, s, MODE_THETA, propto, jacobian, is_var, 0.0, lik_only, sigma_is_var, theta, h->host_mirror);
  if (rc) return rc;
  double* hres = s->h_out;
  bool mirrored = false;
  if (h->host_mirror) {
    rc = wait_host_out(h, s, &mirrored);
    if (rc) return rc;
  }
  if (!mirrored) {
    hres = s->h_pinned + (3 * P + 1);
    CUDA_TRY(h, cudaMemcpyAsync(hres, s->result, sizeof(double) * (P + 2), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(h, cudaStreamSynchronize(s->stream));
  }
  // neg_binomial_2_log checks y before its include_summand early return (neg_binomial_2_log_glm_lpmf.hpp:130-135)
  // neg_binomial_2_log, ordered_logistic and categorical_logit check y before their include_summand early return
  // (ordered_logistic_glm_lpmf.hpp:84, categorical_logit_glm_lpmf.hpp:74); ordered_logistic even before size_zero
  const bool y_checked_always = h->d.family == B200GLM_NEG_BINOMIAL_2_LOG || fam_is_class(h->d.family);
  const bool lik_on = lik_rows_total(h) > 0 || (h->d.family == B200GLM_ORDERED_LOGISTIC && h->d.N > 0);
  if (h->bad_y && (is_var || !propto || y_checked_always) && lik_on) {
    static const char* const msg[] = {
        "bernoulli_logit_glm_lpmf: Vector of dependent variables is out of range [0, 1]",
        "poisson_log_glm_lpmf: Vector of dependent variables is negative", "",
        "binomial_logit_glm_lpmf: Successes variable is out of range [0, Population size parameter]",
        "neg_binomial_2_log_glm_lpmf: Failures variables is negative",
        "ordered_logistic_glm_lpmf: Vector of dependent variables is out of range [1, N_classes]",
        "categorical_logit_glm_lpmf: categorical outcome out of support"};
    h->set_error(msg[h->d.family]);
    return B200GLM_DOMAIN;
  }
  if (hres[P + 1] == (double)ST_PEER_TIMEOUT) {
    h->set_error("peer exchange timed out: a rank did not launch the matching evaluation");
    return B200GLM_CUDA;
  }
  if (hres[P + 1] != 0.0) {
    h->set_error("non-finite log density or gradient (parameters, intercept or X*beta not finite)");
    return B200GLM_DOMAIN;
  }
  *lp = hres[0];
  if (grad) std::memcpy(grad, hres + 1, sizeof(double) * P);
  return B200GLM_OK;
}

}  // namespace

extern "C" {

#define B200GLM_STR2(x) #x
#define B200GLM_STR(x) B200GLM_STR2(x)
const char* b200glm_version(void) { return "b200glm 0.2 (sm_100a, abi " B200GLM_STR(B200GLM_ABI_VERSION) ")"; }
int32_t b200glm_abi_version(void) { return B200GLM_ABI_VERSION; }

int32_t b200glm_num_params(const b200glm_handle* h) { return h ? h->P : -1; }

const char* b200glm_last_error(const b200glm_handle* h) {
  static thread_local std::string copy;
  if (!h) return "null handle";
  auto* hh = const_cast<b200glm_handle*>(h);
  std::lock_guard<std::mutex> g(hh->err_mu);
  copy = hh->last_error;
  return copy.c_str();
}

int64_t b200glm_launch_count(const b200glm_handle* h) { return h ? (int64_t)h->launches.load() : 0; }

int64_t b200glm_bytes_per_gradient(const b200glm_handle* h) {
  if (!h) return 0;
  // 8*N*K (X once) + 4*N (y int32; 8*N for normal's fp64 y) [+ 4*N group index]
  const int64_t N = h->d.N, K = h->d.K;
  int64_t b = 8 * N * K + (h->d.family == B200GLM_NORMAL_ID ? 8 : 4) * N;
  if (h->d.family == B200GLM_BINOMIAL_LOGIT) b += 4 * N;   // population sizes
  if (h->d.G > 0) b += 4 * N;
  return b;
}

void b200glm_destroy(b200glm_handle* h) {
  if (!h) return;
  cudaSetDevice(h->d.device);
  for (Slot* s : h->slots) {
    if (!s) continue;
    if (s->stream) cudaStreamSynchronize(s->stream);
    cudaFree(s->theta);
    cudaFree(s->state[0]);
    cudaFree(s->state[1]);
    cudaFree(s->inv_metric);
    cudaFree(s->partials);
    cudaFree(s->ticket);
    cudaFree(s->r_out);
    cudaFree(s->gpart);
    cudaFree(s->cuts);
    cudaFree(s->alpha_rows);
    cudaFree(s->sigma_rows);
    cudaFree(s->s_out);
    cudaFree(s->lik);
    cudaFree(s->result);
    cudaFree(s->theta_used);
    cudaFree(s->tl);
    if (s->h_pinned) cudaFreeHost(s->h_pinned);
    if (s->h_out) cudaFreeHost(s->h_out);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
  }
  if (Batch* b = h->batch) {
    if (b->stream) cudaStreamSynchronize(b->stream);
    free_nuts(b);
    for (double* q : {b->Q, b->Pm, b->Gd, b->V, b->IM, b->theta_c, b->p_half, b->partials, b->reduced, b->result, b->state_out,
                      b->theta_in, b->eps_d})
      cudaFree(q);
    cudaFree(b->chains_d);
    if (b->h_pin) cudaFreeHost(b->h_pin);
    if (b->h_pin_i) cudaFreeHost(b->h_pin_i);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
  }
  if (h->comm && nccl().ok) nccl().CommDestroy(h->comm);
  for (int r = 0; r < MAX_PEERS; ++r)
    if (h->peer_mbox[r] && h->peer_mbox[r] != h->mbox) cudaIpcCloseMemHandle(h->peer_mbox[r]);
  cudaFree(h->mbox);
  cudaFree(h->panels);
  cudaFree(h->seg_ptr);
  cudaFree(h->gmeta);
  cudaFree(h->cta_g0);
  delete h;
}

int b200glm_create(const b200glm_desc* desc, b200glm_handle** out) {
  if (!desc || !out) return B200GLM_INVALID;
  *out = nullptr;
  b200glm_handle* h = new b200glm_handle();
  h->d = *desc;
  auto fail = [&](int code, const std::string& msg) {
    // keep the handle alive so the caller can read last_error (and must b200glm_destroy it), as documented;
    // whatever the handle owns by now is freed there, the upload temporaries by `tmp` below
    h->set_error(msg);
    *out = h;
    return code;
  };
#define CREATE_TRY(expr)                                                                         \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) return fail(B200GLM_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)
  // device temporaries of the upload: freed on every exit path
  struct DevTmp {
    std::vector<void**> v;
    cudaStream_t* st = nullptr;
    void track(void** p) { v.push_back(p); }
    ~DevTmp() {
      for (void** p : v)
        if (*p) cudaFree(*p);
      if (st && *st) cudaStreamDestroy(*st);
    }
  } tmp;
  const b200glm_desc& d = h->d;
  if (d.family < 0 || d.family > 6) return fail(B200GLM_INVALID, "unknown family");
  const bool cls = fam_is_class(d.family);
  if (cls && (d.n_classes < 1 || d.n_classes > CLASS_MAX_CLASSES))
    return fail(B200GLM_INVALID, "class-outcome models need 1 <= n_classes <= 16");
  if (cls && (d.G > 0 || (d.flags & (B200GLM_FLAG_FORCE_WIDE | B200GLM_FLAG_STREAMED))))
    return fail(B200GLM_INVALID, "class-outcome models: no group intercepts, wide-matrix kernel or streamed construction");
  h->streamed = (d.flags & B200GLM_FLAG_STREAMED) != 0;
  if (d.N < 0 || d.K < 0 || d.G < 0) return fail(B200GLM_INVALID, "negative size");
  if (h->streamed) {
    if (d.G > 0) return fail(B200GLM_INVALID, "streamed construction needs a scalar intercept (G == 0): the group sort needs all rows");
    h->ready = d.N == 0;
  } else {
    if (d.N > 0 && d.family == B200GLM_BINOMIAL_LOGIT && !d.trials) return fail(B200GLM_INVALID, "trials is null");
    if (d.N > 0 && d.K > 0 && (!d.X || d.ldx < d.N)) return fail(B200GLM_INVALID, "X null or ldx < N");
    if (d.N > 0 && d.family == B200GLM_NORMAL_ID && !d.y_real) return fail(B200GLM_INVALID, "y_real is null");
    if (d.N > 0 && d.family != B200GLM_NORMAL_ID && !d.y_int) return fail(B200GLM_INVALID, "y_int is null");
    if (d.G > 0 && d.N > 0 && !d.group) return fail(B200GLM_INVALID, "group is null");
  }
  if (!(d.prior_alpha_sd > 0) || !(d.prior_beta_sd > 0)) return fail(B200GLM_INVALID, "prior scales must be > 0");
  if (d.n_slots < 1) h->d.n_slots = 1;
  if (const char* e = std::getenv("B200GLM_NO_PDL")) h->pdl = !(e[0] == '1');
  if (const char* e = std::getenv("B200GLM_PDL_PREFETCH")) h->pdl_prefetch = std::max(0, std::atoi(e));
  if (const char* e = std::getenv("B200GLM_TL_REPEAT")) h->tl_repeat = (e[0] == '1');
  if (const char* e = std::getenv("B200GLM_WIDE_PRODUCER")) h->wide_producer = e[0] == 's' ? 0 : 1;
  if (const char* e = std::getenv("B200GLM_NO_INLINE_THETA")) h->inline_theta = !(e[0] == '1');
  if (const char* e = std::getenv("B200GLM_NO_HOST_MIRROR")) h->host_mirror = !(e[0] == '1');
  h->P = (d.G > 0 ? 2 + d.G : 1) + d.K + (fam_has_scale(d.family) ? 1 : 0);
  h->off_beta = d.G > 0 ? 2 + d.G : 1;
  if (d.family == B200GLM_ORDERED_LOGISTIC) h->P = d.K + d.n_classes - 1;        // [beta, unconstrained cut-points]
  if (d.family == B200GLM_CATEGORICAL_LOGIT) h->P = d.n_classes * (1 + d.K);      // [alpha, beta column-major]
  if (cls) h->off_beta = d.family == B200GLM_ORDERED_LOGISTIC ? 0 : d.n_classes;
  h->C = fam_group_col(d.family, d.K) + (d.G > 0 ? 1 : 0);   // class-outcome models: K + 1 (the class in column K)
  h->wide = !cls && (d.K > 256 || (d.flags & B200GLM_FLAG_FORCE_WIDE));
  h->panel_rows = h->wide ? wide_rows_for(h->C) : PANEL_ROWS;
  if (h->wide)
    if (const char* e = std::getenv("B200GLM_WIDE_ROWS")) {   // A/B runs: 16, 8 or 4 rows per panel
      const int wr = std::atoi(e);
      if (wr == 16 || wr == 8 || wr == 4) h->panel_rows = wr;
    }
  h->n_panels = (d.N + h->panel_rows - 1) / h->panel_rows;
  const int P = h->P;

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(B200GLM_CUDA, "no CUDA device available (this backend has no CPU fallback)");
  if (d.device < 0 || d.device >= ndev) return fail(B200GLM_INVALID, "device ordinal out of range");
  CREATE_TRY(cudaSetDevice(d.device));
  cudaDeviceProp prop;
  CREATE_TRY(cudaGetDeviceProperties(&prop, d.device));
  if (prop.major < 10) return fail(B200GLM_CUDA, "device is not sm_100 or newer");

  // launch geometry
  h->grid = d.grid_ctas > 0 ? d.grid_ctas : prop.multiProcessorCount;
  h->stage_a = (d.G > 0 && d.G <= SMEM_A_MAX_GROUPS) ? 1 : 0;
  // on-chip chain state for the fused epilogue (theta, half-updated momentum, old gradient, likelihood sums)
  h->state_smem = (size_t)state_smem_doubles(P) * 8 <= 49152 ? 1 : 0;
  if (std::getenv("B200GLM_NO_STATE_SMEM")) h->state_smem = 0;   // A/B runs
  int P_state = h->state_smem ? P : 0;
  const size_t max_dyn = (size_t)prop.sharedMemPerBlockOptin - 1024;  // static scratch + slack
  // narrow kernel: ring stages for the handle's current (stage_a, state_smem, group_fused / Gcs); called again after the
  // group sort has decided whether the group path is fused
  auto config_narrow = [&]() -> bool {
    const size_t tile_bytes = (size_t)h->C * PANEL_ROWS * 8;
    auto stages_for = [&](int p_state) {
      int S = MAX_STAGES;
      while (S > 0 && fixed_smem_bytes(d.K, d.G, h->stage_a, S, p_state, h->group_fused ? h->Gcs : 0)
                              + (size_t)S * tile_bytes > max_dyn)
        --S;
      // Any S works: a stage belongs to one consumer warp for the whole launch (glm_kernels.cuh).  Round 1 rounded S
      // down to a multiple of 8: at K = 50 that left ONE stage per warp, and each warp sat out a full TMA latency
      // between panels -- 77 % of DRAM peak instead of ~90.
      if (const char* e = std::getenv("B200GLM_STAGES_MULT8"))
        if (e[0] == '1' && S > NUM_CONSUMER_WARPS) S = (S / NUM_CONSUMER_WARPS) * NUM_CONSUMER_WARPS;
      if (const char* e = std::getenv("B200GLM_STAGES")) S = std::max(1, std::min(S, std::atoi(e)));   // A/B runs
      return S;
    };
    int S = stages_for(h->state_smem ? P : 0);
    if (h->state_smem && S < NUM_CONSUMER_WARPS && stages_for(0) > S) {   // the ring comes first
      h->state_smem = 0;
      S = stages_for(0);
    }
    if (S < 1) return false;
    h->n_stages = S;
    h->smem_bytes = fixed_smem_bytes(d.K, d.G, h->stage_a, S, h->state_smem ? P : 0, h->group_fused ? h->Gcs : 0)
                    + (size_t)S * tile_bytes;
    return true;
  };
  h->pstride = cls ? class_partial_stride(P) : partial_stride(d.K);
  if (cls) {
    h->Cpad = h->C;
    h->state_smem = 0;
    const int need_cpl = std::max(1, (d.K + 7) / 8);
    h->cpl = 0;
    if (d.family == B200GLM_ORDERED_LOGISTIC) {
      for (const ClassShape& sh : kOrderedShapes)
        if (!h->cpl && sh.cpl >= need_cpl) h->cpl = sh.cpl, h->cmax = sh.cmax;
    } else {
      for (const ClassShape& sh : kCategoricalShapes)   // fewest classes in registers first (widest K)
        if (!h->cpl && sh.cmax >= std::max(2, (int)d.n_classes) && sh.cpl >= need_cpl) h->cpl = sh.cpl, h->cmax = sh.cmax;
    }
    if (!h->cpl)
      return fail(B200GLM_INVALID, d.family == B200GLM_ORDERED_LOGISTIC
                                       ? "ordered_logistic: K <= 256"
                                       : "categorical_logit: K <= 200 with 2 classes, <= 104 with <= 4, <= 56 with <= 8, "
                                         "<= 16 with <= 16");
    const size_t fixed = class_fixed_doubles(d.family == B200GLM_ORDERED_LOGISTIC, d.K, d.n_classes, P) * 8;
    const size_t tile_bytes = (size_t)h->C * PANEL_ROWS * 8;
    int S = MAX_STAGES;
    while (S > 0 && fixed + (size_t)S * (tile_bytes + 16) > max_dyn) --S;
    if (S < 1) return fail(B200GLM_INVALID, "panel does not fit in shared memory");
    h->n_stages = S;
    h->smem_bytes = fixed + (size_t)S * (tile_bytes + 16);
  } else if (!h->wide) {
    h->Cpad = h->C;
    const int need_cpl = std::max(1, (d.K + 7) / 8);
    h->cpl = 0;
    for (int c : kCplChoices)
      if (c >= need_cpl) {
        h->cpl = c;
        break;
      }
    if (!h->cpl) return fail(B200GLM_INVALID, "K too large");
    if (!config_narrow()) return fail(B200GLM_INVALID, "panel does not fit in shared memory");
  } else {
    // sub-panels: 8 warp steps wide unless 4 steps already give every warp at most one sub-panel
    const int WR = h->panel_rows, cps = wide_cps(WR);
    h->Cpad = ((h->C + cps - 1) / cps) * cps;
    h->spc = h->Cpad <= WIDE_CONSUMER_WARPS * 4 * cps ? 4 : 8;
    if (WR == 8) h->spc = 8;
    h->Kc = cps * h->spc;
    h->J = (h->Cpad + h->Kc - 1) / h->Kc;
    h->spw = (h->J + WIDE_CONSUMER_WARPS - 1) / WIDE_CONSUMER_WARPS;
    if (!pick_wide_kernel(d.family, WR, h->spw, h->spc))
      return fail(B200GLM_INVALID, "K too large for the wide kernel (K <= 3000)");
    const size_t slot_bytes = (size_t)h->Kc * WR * 8;
    size_t fixed = 0;
    int T = 0;
    for (;;) {   // the ring comes first: drop the on-chip state if THREE row panels would not fit with it (the two
                 // resident panels + a whole panel of look-ahead: what the lane-parallel TMA producer needs; K = 1000:
                 // T = 23 -> 27 slots, 1.35 -> 1.14 ms per gradient at N = 1M)
      fixed = wide_fixed_doubles(WR, h->J, h->Kc, d.G, h->stage_a, P_state) * 8;
      T = fixed < max_dyn ? (int)((max_dyn - fixed) / (slot_bytes + 16)) : 0;
      if (T > WIDE_MAX_SLOTS) T = WIDE_MAX_SLOTS;
      if (T >= 3 * h->J || !P_state) break;
      P_state = 0;
      h->state_smem = 0;
    }
    if (T < 2 * h->J) return fail(B200GLM_INVALID, "two row panels do not fit in shared memory (K too large)");
    h->n_stages = T;
    h->smem_bytes = fixed + (size_t)T * (slot_bytes + 16);
    // TMA producer form (glm_wide_kernel.cuh): the lanes' joint wait needs a whole panel of free slots beyond the two
    // resident panels; with a shallower ring the single-lane loop is the faster one (measured, profiles/r2_wide_variants.txt)
    if (h->wide_producer < 0) h->wide_producer = T >= 3 * h->J ? 1 : 0;
  }
  // the attribute belongs to the kernel, not to the handle: always the device maximum, so that handles of
  // different shapes sharing one instantiation cannot lower it for each other
  kernel_fn fn = handle_kernel(h);
  CREATE_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin - 1024));

  // ---- data upload + re-layout ----
  cudaStream_t st = nullptr;
  tmp.st = &st;
  CREATE_TRY(cudaStreamCreate(&st));
  const size_t panel_doubles = (size_t)h->n_panels * h->Cpad * h->panel_rows;
  if (panel_doubles) CREATE_TRY(cudaMalloc(&h->panels, panel_doubles * 8));

  // y / group on host and device
  std::vector<int32_t> h_group;
  int32_t* d_y = nullptr;
  double* d_yr = nullptr;
  int32_t* d_group = nullptr;
  int32_t* d_trials = nullptr;
  long long* d_perm = nullptr;
  const bool upload_now = d.N > 0 && !h->streamed;
  bool own_y = false, own_group = false;
  void *t_y = nullptr, *t_yr = nullptr, *t_trials = nullptr, *t_group = nullptr, *t_perm = nullptr, *t_X = nullptr,
       *t_stats = nullptr;   // what `tmp` frees: the OWNED temporaries only (never the caller's device pointers)
  for (void** q : {&t_y, &t_yr, &t_trials, &t_group, &t_perm, &t_X, &t_stats}) tmp.track(q);
  if (upload_now) {
    if (d.data_on_device) {
      d_y = const_cast<int32_t*>(d.y_int);
      d_yr = const_cast<double*>(d.y_real);
      d_group = const_cast<int32_t*>(d.group);
      d_trials = const_cast<int32_t*>(d.trials);
    } else {
      if (d.family == B200GLM_NORMAL_ID) {
        CREATE_TRY(cudaMalloc(&d_yr, sizeof(double) * d.N));
        t_yr = d_yr;
        CREATE_TRY(cudaMemcpy(d_yr, d.y_real, sizeof(double) * d.N, cudaMemcpyHostToDevice));
      } else {
        CREATE_TRY(cudaMalloc(&d_y, sizeof(int32_t) * d.N));
        t_y = d_y;
        CREATE_TRY(cudaMemcpy(d_y, d.y_int, sizeof(int32_t) * d.N, cudaMemcpyHostToDevice));
        if (d.family == B200GLM_BINOMIAL_LOGIT) {
          CREATE_TRY(cudaMalloc(&d_trials, sizeof(int32_t) * d.N));
          t_trials = d_trials;
          CREATE_TRY(cudaMemcpy(d_trials, d.trials, sizeof(int32_t) * d.N, cudaMemcpyHostToDevice));
        }
      }
      own_y = true;
      if (d.G > 0) {
        CREATE_TRY(cudaMalloc(&d_group, sizeof(int32_t) * d.N));
        t_group = d_group;
        CREATE_TRY(cudaMemcpy(d_group, d.group, sizeof(int32_t) * d.N, cudaMemcpyHostToDevice));
        own_group = true;
      }
    }
  }
  // data checks + lgamma constant
  if (upload_now && d.family != B200GLM_NORMAL_ID) {
    const int nb = 296;
    double* d_stats;
    CREATE_TRY(cudaMalloc(&d_stats, sizeof(double) * 2 * nb));
    t_stats = d_stats;
    y_stats_kernel<<<nb, 256, 0, st>>>(d_y, d_trials, d.N, d.family, d_stats, d.n_classes);
    std::vector<double> hs(2 * nb);
    CREATE_TRY(cudaMemcpyAsync(hs.data(), d_stats, sizeof(double) * 2 * nb, cudaMemcpyDeviceToHost, st));
    CREATE_TRY(cudaStreamSynchronize(st));
    cudaFree(d_stats);
    t_stats = nullptr;
    double bad = 0, lg = 0;
    for (int i = 0; i < nb; ++i) {
      bad += hs[2 * i];
      lg += hs[2 * i + 1];
    }
    h->bad_y = h->bad_y_local = bad > 0;
    if (d.family == B200GLM_CATEGORICAL_LOGIT && d.n_classes == 1)
      h->bad_y = h->bad_y_local = false;   // N_classes == 1 returns 0 BEFORE the bounds check (categorical...:70-75)
    h->lgamma_sum = lg;
    h->lgamma_sum_total = lg;
  }
  // group sort (stable counting sort on the host; rows of a group become contiguous)
  if (d.G > 0) {
    h_group.resize(d.N);
    if (d.N > 0) {
      if (d.data_on_device)
        CREATE_TRY(cudaMemcpy(h_group.data(), d_group, sizeof(int32_t) * d.N, cudaMemcpyDeviceToHost));
      else
        std::memcpy(h_group.data(), d.group, sizeof(int32_t) * d.N);
    }
    std::vector<long long> seg(d.G + 1, 0);
    for (long long i = 0; i < d.N; ++i) {
      const int g = h_group[i];
      if (g < 1 || g > d.G) return fail(B200GLM_INVALID, "group index out of range [1, G]");
      seg[g]++;
    }
    for (int g = 0; g < d.G; ++g) seg[g + 1] += seg[g];
    std::vector<long long> perm(std::max<long long>(d.N, 1));
    {
      std::vector<long long> cursor(seg.begin(), seg.end() - 1);
      for (long long i = 0; i < d.N; ++i) perm[cursor[h_group[i] - 1]++] = i;
    }
    // Fused group path (narrow kernel): CTA c owns the contiguous panels [c n_panels / grid, (c + 1) n_panels / grid)
    // of the SORTED rows; it is taken when no CTA meets more than 512 groups (8 warps x Gcs doubles of smem).
    if (!h->wide && d.N > 0 && !std::getenv("B200GLM_NO_GROUP_FUSION")) {
      std::vector<int> g0(h->grid, 0), gn(h->grid, 0);
      auto group_of_row = [&](long long r) {   // 0-based group of sorted row r
        return (int)(std::upper_bound(seg.begin(), seg.end(), r) - seg.begin()) - 1;
      };
      int Gc = 1;
      for (int c = 0; c < h->grid; ++c) {
        const long long pb = ((long long)c * h->n_panels) / h->grid, pe = ((long long)(c + 1) * h->n_panels) / h->grid;
        if (pe <= pb) continue;
        const long long rb = pb * PANEL_ROWS, re = std::min<long long>(pe * PANEL_ROWS, d.N);
        g0[c] = group_of_row(rb);
        gn[c] = group_of_row(re - 1) - g0[c] + 1;
        Gc = std::max(Gc, gn[c]);
      }
      if (Gc <= 512) {
        h->group_fused = 1;
        h->Gcs = (Gc + 1) & ~1;
        std::vector<int4> meta(d.G);
        for (int g = 0; g < d.G; ++g) meta[g] = make_int4(1, 0, 0, 0);   // empty: lo > hi
        for (int c = 0; c < h->grid; ++c)
          for (int j = 0; j < gn[c]; ++j) {
            const int g = g0[c] + j;
            if (seg[g + 1] == seg[g]) continue;                            // no rows: stays empty
            if (meta[g].x > meta[g].y) meta[g] = make_int4(c, c, c * h->Gcs + j, 0);
            else meta[g].y = c;    // later CTAs: the group is their first one (entry 0)
          }
        CREATE_TRY(cudaMalloc(&h->gmeta, sizeof(int4) * d.G));
        CREATE_TRY(cudaMemcpy(h->gmeta, meta.data(), sizeof(int4) * d.G, cudaMemcpyHostToDevice));
        CREATE_TRY(cudaMalloc(&h->cta_g0, sizeof(int) * h->grid));
        CREATE_TRY(cudaMemcpy(h->cta_g0, g0.data(), sizeof(int) * h->grid, cudaMemcpyHostToDevice));
        if (!config_narrow()) return fail(B200GLM_INVALID, "panel does not fit in shared memory");
      }
    }
    CREATE_TRY(cudaMalloc(&h->seg_ptr, sizeof(long long) * (d.G + 1)));
    CREATE_TRY(cudaMemcpy(h->seg_ptr, seg.data(), sizeof(long long) * (d.G + 1), cudaMemcpyHostToDevice));
    if (d.N > 0) {
      CREATE_TRY(cudaMalloc(&d_perm, sizeof(long long) * d.N));
      t_perm = d_perm;
      CREATE_TRY(cudaMemcpy(d_perm, perm.data(), sizeof(long long) * d.N, cudaMemcpyHostToDevice));
    }
  }
  // X: device-resident or staged from the host in row chunks
  const int PR = h->panel_rows, SWZ = h->wide ? 0 : 1;
  if (upload_now) {
    if (d.data_on_device || d.K == 0) {
      relayout_kernel<<<4 * prop.multiProcessorCount, 256, 0, st>>>(d.X, d.ldx, 0, d_y, d_yr, d_group, d_trials, d_perm, 0, d.N,
                                                                    d.N, d.K, h->C, 0, h->Cpad, h->panels, PR, h->Cpad, SWZ);
      CREATE_TRY(cudaGetLastError());
    } else if (d_perm) {
      // permuted gather needs all of X on the device at once
      double* dX;
      CREATE_TRY(cudaMalloc(&dX, sizeof(double) * (size_t)d.N * d.K));
      t_X = dX;
      CREATE_TRY(cudaMemcpy2D(dX, sizeof(double) * d.N, d.X, sizeof(double) * d.ldx, sizeof(double) * d.N, d.K,
                               cudaMemcpyHostToDevice));
      relayout_kernel<<<4 * prop.multiProcessorCount, 256, 0, st>>>(dX, d.N, 0, d_y, d_yr, d_group, d_trials, d_perm, 0, d.N,
                                                                    d.N, d.K, h->C, 0, h->Cpad, h->panels, PR, h->Cpad, SWZ);
      CREATE_TRY(cudaStreamSynchronize(st));
      cudaFree(dX);
      t_X = nullptr;
    } else {
      const long long chunk_rows = std::max<long long>(32, ((long long)(256u << 20) / (8LL * std::max(d.K, 1))) & ~31LL);
      double* dX;
      CREATE_TRY(cudaMalloc(&dX, sizeof(double) * (size_t)std::min(chunk_rows, (long long)d.N) * d.K));
      t_X = dX;
      for (long long r0 = 0; r0 < d.N; r0 += chunk_rows) {
        const long long nr = std::min(chunk_rows, (long long)d.N - r0);
        CREATE_TRY(cudaMemcpy2DAsync(dX, sizeof(double) * nr, d.X + r0, sizeof(double) * d.ldx, sizeof(double) * nr,
                                      d.K, cudaMemcpyHostToDevice, st));
        relayout_kernel<<<4 * prop.multiProcessorCount, 256, 0, st>>>(dX, nr, r0, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                                      r0, nr, d.N, d.K, h->C, 0, d.K, h->panels, PR, h->Cpad, SWZ);
        CREATE_TRY(cudaStreamSynchronize(st));
      }
      cudaFree(dX);
      t_X = nullptr;
      // aux columns (y, group) in one more pass
      relayout_kernel<<<4 * prop.multiProcessorCount, 256, 0, st>>>(nullptr, 0, 0, d_y, d_yr, d_group, d_trials, nullptr, 0, d.N,
                                                                    d.N, d.K, h->C, d.K, h->Cpad, h->panels, PR, h->Cpad, SWZ);
    }
    CREATE_TRY(cudaGetLastError());
    CREATE_TRY(cudaStreamSynchronize(st));
  }
  (void)own_y;
  (void)own_group;   // the upload temporaries and the stream are released by `tmp`

  // ---- slots ----
  for (int i = 0; i < h->d.n_slots; ++i) {
    Slot* s = new Slot();
    h->slots.push_back(s);
    CREATE_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaMalloc(&s->theta, sizeof(double) * std::max(P, 1)));
    for (int b = 0; b < 2; ++b) {
      CREATE_TRY(cudaMalloc(&s->state[b], sizeof(double) * (3 * P + 1)));
      CREATE_TRY(cudaMemset(s->state[b], 0, sizeof(double) * (3 * P + 1)));
    }
    CREATE_TRY(cudaMalloc(&s->inv_metric, sizeof(double) * std::max(P, 1)));
    std::vector<double> ones(std::max(P, 1), 1.0);
    CREATE_TRY(cudaMemcpy(s->inv_metric, ones.data(), sizeof(double) * P, cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMalloc(&s->partials, sizeof(double) * (size_t)h->grid * h->pstride));
    if (d.family == B200GLM_ORDERED_LOGISTIC)
      CREATE_TRY(cudaMalloc(&s->cuts, sizeof(double) * 2 * std::max(1, (int)d.n_classes)));
    CREATE_TRY(cudaMalloc(&s->ticket, sizeof(unsigned int)));
    CREATE_TRY(cudaMemset(s->ticket, 0, sizeof(unsigned int)));
    if (d.G > 0 && h->n_panels > 0 && !h->group_fused)
      CREATE_TRY(cudaMalloc(&s->r_out, sizeof(double) * h->n_panels * h->panel_rows));
    if (h->group_fused) CREATE_TRY(cudaMalloc(&s->gpart, sizeof(double) * (size_t)h->grid * h->Gcs));
    CREATE_TRY(cudaMalloc(&s->lik, sizeof(double) * (P + 2)));
    CREATE_TRY(cudaMemset(s->lik, 0, sizeof(double) * (P + 2)));
    CREATE_TRY(cudaMalloc(&s->result, sizeof(double) * (P + 2)));
    CREATE_TRY(cudaMalloc(&s->theta_used, sizeof(double) * std::max(P, 1)));
    CREATE_TRY(cudaMallocHost(&s->h_pinned, sizeof(double) * (2 * (3 * P + 1) + P + 2)));
    CREATE_TRY(cudaHostAlloc(&s->h_out, sizeof(double) * ((P + 2) + (3 * P + 1) + 1), cudaHostAllocMapped));
    std::memset(s->h_out, 0, sizeof(double) * ((P + 2) + (3 * P + 1) + 1));
    CREATE_TRY(cudaHostGetDevicePointer((void**)&s->h_out_dev, s->h_out, 0));
  }
  *out = h;
  return B200GLM_OK;
#undef CREATE_TRY
}

int b200glm_log_prob_grad(b200glm_handle* h, int32_t slot, const double* theta, int32_t propto, int32_t jacobian,
                          double* lp, double* grad) {
  return eval_host(h, slot, theta, propto ? 1 : 0, jacobian ? 1 : 0, 1, lp, grad);
}

int b200glm_log_prob(b200glm_handle* h, int32_t slot, const double* theta, int32_t propto, int32_t jacobian,
                     double* lp) {
  return eval_host(h, slot, theta, propto ? 1 : 0, jacobian ? 1 : 0, 0, lp, nullptr);
}

int b200glm_glm_lpmf(b200glm_handle* h, int32_t slot, int32_t propto, int32_t operands_are_var,
                     int32_t sigma_is_var, const double* alpha, const double* beta, double sigma, double* logp,
                     double* d_alpha, double* d_beta, double* d_sigma) {
  if (!h) return B200GLM_INVALID;
  if (!logp || (h->d.K > 0 && !beta) || !alpha) {
    h->set_error("null pointer argument");
    return B200GLM_INVALID;
  }
  if (fam_is_class(h->d.family)) {
    h->set_error("b200glm_glm_lpmf serves families 0-4; the class-outcome models are evaluated as whole models");
    return B200GLM_INVALID;
  }
  const int P = h->P, G = h->d.G, K = h->d.K;
  const bool normal = fam_has_scale(h->d.family);   // a trailing positive scalar: sigma | phi
  if (normal && !(sigma > 0.0 && std::isfinite(sigma))) {
    h->set_error(h->d.family == B200GLM_NORMAL_ID
                     ? "normal_id_glm_lpdf: Scale vector is not positive finite"                 // normal_id_glm_lpdf.hpp:93
                     : "neg_binomial_2_log_glm_lpmf: Precision parameter is not positive finite");  // neg_binomial_2_log_glm_lpmf.hpp:131
    return B200GLM_DOMAIN;
  }
  std::vector<double> th(P, 0.0), g(P, 0.0);
  if (G > 0)
    std::memcpy(th.data() + 2, alpha, sizeof(double) * G);
  else
    th[0] = alpha[0];
  if (K > 0) std::memcpy(th.data() + h->off_beta, beta, sizeof(double) * K);
  if (normal) th[P - 1] = std::log(sigma);
  const int rc = eval_host(h, slot, th.data(), propto ? 1 : 0, 0, operands_are_var ? 1 : 0, logp, g.data(), 1,
                           sigma_is_var == 2 ? 2 : (sigma_is_var ? 1 : 0));
  if (rc) return rc;
  if (d_alpha) std::memcpy(d_alpha, G > 0 ? g.data() + 2 : g.data(), sizeof(double) * (G > 0 ? G : 1));
  if (d_beta && K > 0) std::memcpy(d_beta, g.data() + h->off_beta, sizeof(double) * K);
  if (d_sigma) *d_sigma = normal ? g[P - 1] : 0.0;
  return B200GLM_OK;
}

int b200glm_glm_lpmf_rows(b200glm_handle* h, int32_t slot, int32_t propto, int32_t operands_are_var,
                          int32_t sigma_is_var, const double* alpha_rows, double alpha, const double* beta,
                          const double* sigma_rows, double sigma, double* logp, double* d_alpha_rows, double* d_alpha,
                          double* d_beta, double* d_sigma_rows, double* d_sigma) {
  int rc = validate_slot(h, slot);
  if (rc) return rc;
  const b200glm_desc& d = h->d;
  if (!logp || (d.K > 0 && !beta)) {
    h->set_error("null pointer argument");
    return B200GLM_INVALID;
  }
  if (fam_is_class(d.family) || d.G > 0 || h->wide || d.world > 1) {
    h->set_error("b200glm_glm_lpmf_rows: per-row operands are served for families 0-4 with a scalar-intercept handle "
                 "(G == 0), K <= 256, unsharded");
    return B200GLM_INVALID;
  }
  if (sigma_rows && d.family != B200GLM_NORMAL_ID) {
    h->set_error("b200glm_glm_lpmf_rows: a per-row scale is an operand of normal_id_glm_lpdf only");
    return B200GLM_INVALID;
  }
  const long long N = d.N;
  const bool has_scale = fam_has_scale(d.family);
  if (has_scale && !sigma_rows && !(sigma > 0.0 && std::isfinite(sigma))) {
    h->set_error(d.family == B200GLM_NORMAL_ID ? "normal_id_glm_lpdf: Scale vector is not positive finite"
                                                 : "neg_binomial_2_log_glm_lpmf: Precision parameter is not positive finite");
    return B200GLM_DOMAIN;
  }
  if (sigma_rows)
    for (long long i = 0; i < N; ++i)
      if (!(sigma_rows[i] > 0.0 && std::isfinite(sigma_rows[i]))) {
        h->set_error("normal_id_glm_lpdf: Scale vector is not positive finite");   // normal_id_glm_lpdf.hpp:93
        return B200GLM_DOMAIN;
      }
  if (alpha_rows)
    for (long long i = 0; i < N; ++i)
      if (!std::isfinite(alpha_rows[i])) {
        h->set_error("Intercept is not finite");
        return B200GLM_DOMAIN;
      }
  Slot* s = h->slots[slot];
  const int P = h->P, K = d.K;
  std::vector<double> th(P, 0.0), g(P, 0.0);
  std::lock_guard<std::recursive_mutex> lk_all(s->mu);   // the per-row flags of the slot belong to this call
  {
    CUDA_TRY(h, cudaSetDevice(d.device));
    const size_t nb = sizeof(double) * (size_t)std::max<long long>(N, 1);
    if (alpha_rows) {
      if (!s->alpha_rows) CUDA_TRY(h, cudaMalloc(&s->alpha_rows, nb));
      if (!s->r_out) CUDA_TRY(h, cudaMalloc(&s->r_out, sizeof(double) * std::max<long long>(h->n_panels * h->panel_rows, 1)));
      CUDA_TRY(h, cudaMemcpyAsync(s->alpha_rows, alpha_rows, sizeof(double) * N, cudaMemcpyHostToDevice, s->stream));
    }
    if (sigma_rows) {
      if (!s->sigma_rows) CUDA_TRY(h, cudaMalloc(&s->sigma_rows, nb));
      if (!s->s_out) CUDA_TRY(h, cudaMalloc(&s->s_out, nb));
      CUDA_TRY(h, cudaMemcpyAsync(s->sigma_rows, sigma_rows, sizeof(double) * N, cudaMemcpyHostToDevice, s->stream));
    }
    s->use_alpha_rows = alpha_rows != nullptr;
    s->use_sigma_rows = sigma_rows != nullptr;
  }
  th[0] = alpha_rows ? 0.0 : alpha;
  if (K > 0) std::memcpy(th.data() + h->off_beta, beta, sizeof(double) * K);
  if (has_scale) th[P - 1] = sigma_rows ? 0.0 : std::log(sigma);
  rc = eval_host(h, slot, th.data(), propto ? 1 : 0, 0, operands_are_var ? 1 : 0, logp, g.data(), 1,
                 sigma_is_var == 2 ? 2 : (sigma_is_var ? 1 : 0));
  const bool evaluated = ((!propto) || operands_are_var) && N > 0;   // else the kernel did not run: partials are zero
  if (rc == B200GLM_OK) {
    if (d_alpha_rows && alpha_rows) {
      if (evaluated) {
        if (cudaMemcpy(d_alpha_rows, s->r_out, sizeof(double) * N, cudaMemcpyDeviceToHost) != cudaSuccess) rc = B200GLM_CUDA;
      } else {
        std::memset(d_alpha_rows, 0, sizeof(double) * N);
      }
    }
    if (d_sigma_rows && sigma_rows) {
      if (evaluated) {
        if (cudaMemcpy(d_sigma_rows, s->s_out, sizeof(double) * N, cudaMemcpyDeviceToHost) != cudaSuccess) rc = B200GLM_CUDA;
      } else {
        std::memset(d_sigma_rows, 0, sizeof(double) * N);
      }
    }
  }
  s->use_alpha_rows = s->use_sigma_rows = false;
  if (rc) return rc;
  if (d_alpha) *d_alpha = alpha_rows ? 0.0 : g[0];
  if (d_beta && K > 0) std::memcpy(d_beta, g.data() + h->off_beta, sizeof(double) * K);
  if (d_sigma) *d_sigma = (has_scale && !sigma_rows) ? g[P - 1] : 0.0;
  return B200GLM_OK;
}

int b200glm_set_state(b200glm_handle* h, int32_t slot, const double* q, const double* p, const double* g, double V) {
  int rc = validate_slot(h, slot);
  if (rc) return rc;
  if (!q || !p || !g) {
    h->set_error("null pointer argument");
    return B200GLM_INVALID;
  }
  Slot* s = h->slots[slot];
  std::lock_guard<std::recursive_mutex> lk(s->mu);
  CUDA_TRY(h, cudaSetDevice(h->d.device));
  const int P = h->P;
  double* hp = s->h_pinned;
  std::memcpy(hp, q, sizeof(double) * P);
  std::memcpy(hp + P, p, sizeof(double) * P);
  std::memcpy(hp + 2 * P, g, sizeof(double) * P);
  hp[3 * P] = V;
  CUDA_TRY(h, cudaMemcpyAsync(s->state[s->cur], hp, sizeof(double) * (3 * P + 1), cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(h, cudaStreamSynchronize(s->stream));
  return B200GLM_OK;
}

int b200glm_leapfrog_async(b200glm_handle* h, int32_t slot, double eps) {
  int rc = validate_slot(h, slot);
  if (rc) return rc;
  Slot* s = h->slots[slot];
  CUDA_TRY(h, cudaSetDevice(h->d.device));
  rc = enqueue_eval(h, s, MODE_LEAPFROG, 1, 1, 1, eps);
  if (rc) return rc;
  s->cur ^= 1;
  return B200GLM_OK;
}

int b200glm_leapfrog(b200glm_handle* h, int32_t slot, double eps, const double* inv_metric, double* q, double* p,
                     double* g, double* V) {
  int rc = validate_slot(h, slot);
  if (rc) return rc;
  Slot* s = h->slots[slot];
  std::lock_guard<std::recursive_mutex> lk(s->mu);
  CUDA_TRY(h, cudaSetDevice(h->d.device));
  const int P = h->P;
  if (inv_metric) {
    double* hm = s->h_pinned + (3 * P + 1) + (P + 2);
    std::memcpy(hm, inv_metric, sizeof(double) * P);
    CUDA_TRY(h, cudaMemcpyAsync(s->inv_metric, hm, sizeof(double) * P, cudaMemcpyHostToDevice, s->stream));
  }
  rc = enqueue_eval(h, s, MODE_LEAPFROG, 1, 1, 1, eps, 0, 0, nullptr, h->host_mirror);
  if (rc) return rc;
  s->cur ^= 1;
  double* hp = s->h_out + (P + 2);
  bool mirrored = false;
  if (h->host_mirror) {
    rc = wait_host_out(h, s, &mirrored);
    if (rc) return rc;
  }
  if (!mirrored) {
    hp = s->h_pinned;
    CUDA_TRY(h, cudaMemcpyAsync(hp, s->state[s->cur], sizeof(double) * (3 * P + 1), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(h, cudaStreamSynchronize(s->stream));
  }
  if (q) std::memcpy(q, hp, sizeof(double) * P);
  if (p) std::memcpy(p, hp + P, sizeof(double) * P);
  if (g) std::memcpy(g, hp + 2 * P, sizeof(double) * P);
  if (V) *V = hp[3 * P];
  if (h->peer_on && std::isnan(hp[3 * P])) {
    h->set_error("peer exchange timed out: a rank did not launch the matching leapfrog step");
    return B200GLM_CUDA;
  }
  if (h->bad_y) {
    h->set_error("dependent variable out of range");
    return B200GLM_DOMAIN;
  }
  return B200GLM_OK;
}

int b200glm_grad_async(b200glm_handle* h, int32_t slot, const double* theta_device) {
  int rc = validate_slot(h, slot);
  if (rc) return rc;
  Slot* s = h->slots[slot];
  CUDA_TRY(h, cudaSetDevice(h->d.device));
  if (theta_device)
    CUDA_TRY(h, cudaMemcpyAsync(s->theta, theta_device, sizeof(double) * h->P, cudaMemcpyDeviceToDevice, s->stream));
  return enqueue_eval(h, s, MODE_THETA, 1, 1, 1, 0.0);
}

int b200glm_sync(b200glm_handle* h, int32_t slot) {
  int rc = validate_slot(h, slot);
  if (rc) return rc;
  CUDA_TRY(h, cudaStreamSynchronize(h->slots[slot]->stream));
  return B200GLM_OK;
}

void* b200glm_stream(b200glm_handle* h, int32_t slot) {
  if (validate_slot(h, slot)) return nullptr;
  return (void*)h->slots[slot]->stream;
}

const double* b200glm_result_device(b200glm_handle* h, int32_t slot) {
  if (validate_slot(h, slot)) return nullptr;
  return h->slots[slot]->result;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Batched chains (fp64 DMMA path)
// ---------------------------------------------------------------------------------------------
namespace {

typedef void (*batched_fn)(const BatchedParams);
template <int FAMILY, bool RS>
batched_fn pick_batched_mbh(int mbh) {
  switch (mbh) {
    case 2: return glm_batched_kernel<FAMILY, 2, RS>;
    case 4: return glm_batched_kernel<FAMILY, 4, RS>;
    case 7: return glm_batched_kernel<FAMILY, 7, RS>;
    case 13: return glm_batched_kernel<FAMILY, 13, RS>;
  }
  return nullptr;
}
template <bool RS>
batched_fn pick_batched_rs(int family, int mbh) {
  switch (family) {
    case FAM_BERNOULLI_LOGIT: return pick_batched_mbh<FAM_BERNOULLI_LOGIT, RS>(mbh);
    case FAM_POISSON_LOG: return pick_batched_mbh<FAM_POISSON_LOG, RS>(mbh);
    case FAM_NORMAL_ID: return pick_batched_mbh<FAM_NORMAL_ID, RS>(mbh);
    case FAM_BINOMIAL_LOGIT: return pick_batched_mbh<FAM_BINOMIAL_LOGIT, RS>(mbh);
  }
  return nullptr;
}
// row_split: the variant for batches of <= 16 lanes (glm_batched_kernel.cuh)
batched_fn pick_batched(int family, int mbh, bool row_split = false) {
  return row_split ? pick_batched_rs<true>(family, mbh) : pick_batched_rs<false>(family, mbh);
}

// few-chain FMA kernel (glm_multi_kernel.cuh): 4 chains per pass, column slots CPL in {4, 8, 13, 16}
template <int FAMILY>
batched_fn pick_multi_cpl(int cpl) {
  switch (cpl) {
    case 4: return glm_multi_kernel<FAMILY, 4, 4>;
    case 8: return glm_multi_kernel<FAMILY, 8, 4>;
    case 13: return glm_multi_kernel<FAMILY, 13, 4>;
    case 16: return glm_multi_kernel<FAMILY, 16, 4>;
  }
  return nullptr;
}
batched_fn pick_multi(int family, int cpl) {
  switch (family) {
    case FAM_BERNOULLI_LOGIT: return pick_multi_cpl<FAM_BERNOULLI_LOGIT>(cpl);
    case FAM_POISSON_LOG: return pick_multi_cpl<FAM_POISSON_LOG>(cpl);
    case FAM_NORMAL_ID: return pick_multi_cpl<FAM_NORMAL_ID>(cpl);
    case FAM_BINOMIAL_LOGIT: return pick_multi_cpl<FAM_BINOMIAL_LOGIT>(cpl);
  }
  return nullptr;
}

int batch_check(b200glm_handle* h, int n) {
  if (!h) return B200GLM_INVALID;
  if (!h->batch) {
    h->set_error("b200glm_batch_reserve was not called");
    return B200GLM_INVALID;
  }
  if (n < 1 || n > h->batch->max_chains) {
    h->set_error("number of chains outside [1, max_chains]");
    return B200GLM_INVALID;
  }
  return B200GLM_OK;
}

// begin -> main -> finish for n lanes on the batch stream.  chains_d / eps_d may be NULL.
int enqueue_batched(b200glm_handle* h, int n, int mode, int propto, int jacobian, const int32_t* chains_d,
                    const double* eps_d, double eps_scalar, bool mirror_state, bool eps_by_chain = false) {
  Batch* b = h->batch;
  const int NCB = (n + BATCH_CB - 1) / BATCH_CB;
  int NS = std::max(1, b->sms / NCB);
  if ((long long)NS > std::max<long long>(h->n_panels, 1)) NS = (int)std::max<long long>(h->n_panels, 1);
  const bool multi = n <= MULTI_MAX_LANES && b->mu_cpl > 0;   // a few chains: the FMA kernel, four per pass (one CTA
                                                              // per SM, every CTA a row slice)
  BatchedStepParams sp;
  std::memset(&sp, 0, sizeof(sp));
  sp.n = n;
  sp.P = h->P;
  sp.K = h->d.K;
  sp.off_beta = h->off_beta;
  sp.family = h->d.family;
  sp.ldc = NCB * BATCH_CB;
  sp.ld_state = b->ld;
  sp.NCB = NCB;
  sp.NS = NS;
  sp.mode = mode;
  sp.chains = chains_d;
  sp.eps = eps_d;
  sp.eps_by_chain = eps_by_chain ? 1 : 0;
  sp.eps_scalar = eps_scalar;
  sp.theta_in = b->theta_in;
  sp.Q = b->Q;
  sp.Pm = b->Pm;
  sp.Gd = b->Gd;
  sp.V = b->V;
  sp.IM = b->IM;
  sp.theta_c = b->theta_c;
  sp.p_half = b->p_half;
  sp.partials = b->partials;
  sp.reduced = b->reduced;
  sp.result = b->result;
  sp.state_out = mirror_state ? b->state_out : nullptr;
  KernelParams kp;
  fill_params(h, h->slots[0], kp, mode, propto, jacobian, 1, 0.0);
  sp.mc = kp.mc;
  const long long tot = (long long)h->P * sp.ldc;
  batched_begin_kernel<<<(int)std::min<long long>((tot + 255) / 256, 4 * b->sms), 256, 0, b->stream>>>(sp);
  BatchedParams bp;
  std::memset(&bp, 0, sizeof(bp));
  bp.panels = h->panels;
  bp.n_rows = h->d.N;
  bp.n_panels = h->n_panels;
  bp.K = h->d.K;
  bp.C = h->C;
  bp.P = h->P;
  bp.off_beta = h->off_beta;
  bp.family = h->d.family;
  const bool row_split = !multi && n <= 16 && b->rs_pairs >= 2;
  bp.n_stages = multi ? b->mu_S : (row_split ? b->rs_S : b->S);
  bp.pairs = row_split ? b->rs_pairs : 4;
  sp.fold = row_split ? b->rs_pairs : 0;
  bp.NCB = NCB;
  bp.NS = NS;
  bp.ldc = sp.ldc;
  bp.n_lanes = n;
  bp.theta_c = b->theta_c;
  bp.partials = b->partials;
  if (multi) {
    for (int l0 = 0; l0 < n; l0 += 4) {   // 5-8 lanes: a second pass (2 x 1.3 ms at cfg2's shape against 4.95 ms DMMA)
      bp.lane0 = l0;
      pick_multi(h->d.family, b->mu_cpl)<<<NS, MULTI_THREADS, b->mu_smem, b->stream>>>(bp);
      if (l0) h->launches++;
    }
  } else
    pick_batched(h->d.family, b->mbh, row_split)<<<NCB * NS, BATCH_THREADS, row_split ? b->rs_smem : b->smem, b->stream>>>(bp);
  batched_reduce_kernel<<<(h->d.K + 2) * NCB, BATCH_CB, 0, b->stream>>>(sp);
  if (h->d.world > 1) {
    // row shards: every rank holds the sums over ITS rows for all chains; one all-reduce of the [K + 2][ldc] block makes
    // them the sums over all rows, identical on every rank (so the replicated epilogue, the device-side state machines
    // and the replicated host drivers stay in step).  The analogue of the P + 2 doubles of the single-chain path.
    ncclResult_t r = nccl().AllReduce(b->reduced, b->reduced, (size_t)(h->d.K + 2) * sp.ldc, ncclFloat64, ncclSum,
                                      h->comm, b->stream);
    if (r != 0) {
      h->set_error(std::string("ncclAllReduce (batched partial sums): ") +
                   (nccl().GetErrorString ? nccl().GetErrorString(r) : "?"));
      return B200GLM_CUDA;
    }
  }
  batched_finish_kernel<<<(n + 31) / 32, 256, 0, b->stream>>>(sp);
  h->launches += 4;
  CUDA_TRY(h, cudaGetLastError());
  return B200GLM_OK;
}

}  // namespace

extern "C" {

int b200glm_batch_reserve(b200glm_handle* h, int32_t max_chains) {
  if (!h) return B200GLM_INVALID;
  if (max_chains < 1) {
    h->set_error("max_chains must be >= 1");
    return B200GLM_INVALID;
  }
  if (h->batch) {
    if (max_chains <= h->batch->max_chains) return B200GLM_OK;  // idempotent for a smaller or equal request
    // a larger request: release the workspace and reserve again (chain state is not preserved; callers re-upload it)
    Batch* b = h->batch;
    cudaSetDevice(h->d.device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    free_nuts(b);
    for (double* q : {b->Q, b->Pm, b->Gd, b->V, b->IM, b->theta_c, b->p_half, b->partials, b->reduced, b->result, b->state_out,
                      b->theta_in, b->eps_d})
      cudaFree(q);
    cudaFree(b->chains_d);
    if (b->h_pin) cudaFreeHost(b->h_pin);
    if (b->h_pin_i) cudaFreeHost(b->h_pin_i);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
    h->batch = nullptr;
  }
  if (h->wide || h->d.G > 0 || h->d.K > BATCH_MAX_K || h->d.family > B200GLM_BINOMIAL_LOGIT) {
    h->set_error("batched chains need K <= 208, a scalar intercept (G == 0) and one of the bernoulli_logit / poisson_log "
                 "/ normal_id / binomial_logit families");
    return B200GLM_INVALID;
  }
  if (h->d.world > 1 && !h->comm) {   // row shards: the slice sums of all chains are combined by ONE all-reduce per round
    h->set_error("batched chains on a row-sharded handle need b200glm_comm_init first (one NCCL all-reduce of the "
                 "(K + 2) x chains partial sums per batched evaluation)");
    return B200GLM_INVALID;
  }
  CUDA_TRY(h, cudaSetDevice(h->d.device));
  cudaDeviceProp prop;
  CUDA_TRY(h, cudaGetDeviceProperties(&prop, h->d.device));
  Batch* b = new Batch();
  b->max_chains = max_chains;
  b->ld = ((max_chains + BATCH_CB - 1) / BATCH_CB) * BATCH_CB;
  b->sms = h->d.grid_ctas > 0 ? h->d.grid_ctas : prop.multiProcessorCount;
  const int MB0 = (((h->d.K + 7) >> 3) + 1) >> 1;
  for (int c : {2, 4, 7, 13})
    if (c >= MB0) {
      b->mbh = c;
      break;
    }
  const size_t max_dyn = (size_t)prop.sharedMemPerBlockOptin - 1024;
  int S = 4;
  while (S > 0 && batched_smem_bytes(h->d.K, h->C, S) > max_dyn) --S;
  if (S < 1) {
    delete b;
    h->set_error("panel + beta block do not fit in shared memory");
    return B200GLM_INVALID;
  }
  b->S = S;
  b->smem = batched_smem_bytes(h->d.K, h->C, S);
  CUDA_TRY(h, cudaFuncSetAttribute(pick_batched(h->d.family, b->mbh), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)b->smem));
  // row-split variant: beta block is K x 16, every pair has a private ring; as many pairs as panels fit, then
  // as many stages per pair (up to 3) as still fit
  {
    int fit = 0;
    while (fit < 12 && batched_smem_bytes(h->d.K, h->C, fit + 1, 16) <= max_dyn) ++fit;
    b->rs_pairs = std::min(4, fit);
    if (b->rs_pairs >= 2 && !std::getenv("B200GLM_NO_ROWSPLIT")) {
      b->rs_S = std::min(3, fit / b->rs_pairs);
      b->rs_smem = batched_smem_bytes(h->d.K, h->C, b->rs_pairs * b->rs_S, 16);
      CUDA_TRY(h, cudaFuncSetAttribute(pick_batched(h->d.family, b->mbh, true),
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->rs_smem));
    } else {
      b->rs_pairs = 0;
    }
  }
  // few-chain FMA kernel: K <= 128, as many ring stages as fit (at least 8: one per consumer warp)
  if (h->d.K >= 1 && h->d.K <= MULTI_MAX_K && !std::getenv("B200GLM_NO_MULTI")) {
    const int need = (h->d.K + 7) / 8;
    for (int c : {4, 8, 13, 16})
      if (c >= need) {
        b->mu_cpl = c;
        break;
      }
    int S2 = 16;
    while (S2 > 0 && multi_smem_bytes(h->d.K, h->C, S2, 4, b->mu_cpl) > max_dyn) --S2;
    if (S2 >= 4 && b->mu_cpl && (size_t)S2 * multi_stage_cols(h->C, b->mu_cpl) * 32 >= (size_t)NUM_CONSUMER_WARPS * (((h->d.K + 7) & ~7) + 2) * 4) {
      b->mu_S = S2;
      b->mu_smem = multi_smem_bytes(h->d.K, h->C, S2, 4, b->mu_cpl);
      CUDA_TRY(h, cudaFuncSetAttribute(pick_multi(h->d.family, b->mu_cpl), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)b->mu_smem));
    } else {
      b->mu_cpl = 0;
    }
  }
  const size_t P = h->P, ld = b->ld;
  CUDA_TRY(h, cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  for (double** q : {&b->Q, &b->Pm, &b->Gd, &b->IM, &b->theta_c, &b->p_half}) {
    CUDA_TRY(h, cudaMalloc(q, sizeof(double) * P * ld));
    CUDA_TRY(h, cudaMemset(*q, 0, sizeof(double) * P * ld));
  }
  {
    std::vector<double> ones(P * ld, 1.0);
    CUDA_TRY(h, cudaMemcpy(b->IM, ones.data(), sizeof(double) * P * ld, cudaMemcpyHostToDevice));
  }
  CUDA_TRY(h, cudaMalloc(&b->V, sizeof(double) * ld));
  CUDA_TRY(h, cudaMemset(b->V, 0, sizeof(double) * ld));
  const size_t n_part = (size_t)(b->sms + ld / BATCH_CB) * (h->d.K + 2) * BATCH_CB;
  CUDA_TRY(h, cudaMalloc(&b->partials, sizeof(double) * n_part));
  CUDA_TRY(h, cudaMalloc(&b->reduced, sizeof(double) * (size_t)(h->d.K + 2) * ld));
  CUDA_TRY(h, cudaMalloc(&b->result, sizeof(double) * ld * (P + 2)));
  CUDA_TRY(h, cudaMalloc(&b->state_out, sizeof(double) * ld * (3 * P + 1)));
  CUDA_TRY(h, cudaMalloc(&b->theta_in, sizeof(double) * ld * (4 * P + 1)));   // also set_state staging
  CUDA_TRY(h, cudaMalloc(&b->eps_d, sizeof(double) * ld));
  CUDA_TRY(h, cudaMalloc(&b->chains_d, sizeof(int32_t) * ld));
  CUDA_TRY(h, cudaMallocHost(&b->h_pin, sizeof(double) * ld * (4 * P + 2)));
  CUDA_TRY(h, cudaMallocHost(&b->h_pin_i, sizeof(int32_t) * ld));
  h->batch = b;
  return B200GLM_OK;
}

int b200glm_log_prob_grad_batched(b200glm_handle* h, int32_t n, const double* theta, int32_t propto,
                                  int32_t jacobian, double* lp, double* grad, int32_t* status) {
  int rc = batch_check(h, n);
  if (rc) return rc;
  if (!theta || !lp) {
    h->set_error("null pointer argument");
    return B200GLM_INVALID;
  }
  Batch* b = h->batch;
  std::lock_guard<std::mutex> lk(b->mu);
  CUDA_TRY(h, cudaSetDevice(h->d.device));
  const size_t P = h->P;
  std::memcpy(b->h_pin, theta, sizeof(double) * n * P);
  CUDA_TRY(h, cudaMemcpyAsync(b->theta_in, b->h_pin, sizeof(double) * n * P, cudaMemcpyHostToDevice, b->stream));
  rc = enqueue_batched(h, n, MODE_THETA, propto ? 1 : 0, jacobian ? 1 : 0, nullptr, nullptr, 0.0, false);
  if (rc) return rc;
  double* hres = b->h_pin;
  CUDA_TRY(h, cudaMemcpyAsync(hres, b->result, sizeof(double) * n * (P + 2), cudaMemcpyDeviceToHost, b->stream));
  CUDA_TRY(h, cudaStreamSynchronize(b->stream));
  int worst = B200GLM_OK;
  for (int i = 0; i < n; ++i) {
    const double* r = hres + (size_t)i * (P + 2);
    const int st = (h->bad_y || r[P + 1] != 0.0) ? B200GLM_DOMAIN : B200GLM_OK;
    if (status) status[i] = st;
    if (st) worst = st;
    lp[i] = r[0];
    if (grad) std::memcpy(grad + (size_t)i * P, r + 1, sizeof(double) * P);
  }
  if (worst && !status) {
    h->set_error("non-finite log density or gradient in at least one chain");
    return worst;
  }
  return B200GLM_OK;
}

int b200glm_set_state_batched(b200glm_handle* h, int32_t n, const int32_t* chains, const double* q,
                              const double* p, const double* g, const double* V, const double* inv_metric) {
  int rc = batch_check(h, n);
  if (rc) return rc;
  if (!q || !p || !g || !V) {
    h->set_error("null pointer argument");
    return B200GLM_INVALID;
  }
  Batch* b = h->batch;
  std::lock_guard<std::mutex> lk(b->mu);
  CUDA_TRY(h, cudaSetDevice(h->d.device));
  const size_t P = h->P;
  for (int i = 0; i < n; ++i) {
    const int c = chains ? chains[i] : i;
    if (c < 0 || c >= b->max_chains) {
      h->set_error("chain slot out of range");
      return B200GLM_INVALID;
    }
  }
  // one pinned staging block [n][4P + 1] = (q, p, g, inv_metric, V), one H2D copy, one scatter kernel
  const size_t W = 4 * P + 1;
  for (int i = 0; i < n; ++i) {
    double* st = b->h_pin + (size_t)i * W;
    std::memcpy(st, q + (size_t)i * P, sizeof(double) * P);
    std::memcpy(st + P, p + (size_t)i * P, sizeof(double) * P);
    std::memcpy(st + 2 * P, g + (size_t)i * P, sizeof(double) * P);
    if (inv_metric) std::memcpy(st + 3 * P, inv_metric + (size_t)i * P, sizeof(double) * P);
    st[4 * P] = V[i];
    b->h_pin_i[i] = chains ? chains[i] : i;
  }
  CUDA_TRY(h, cudaMemcpyAsync(b->theta_in, b->h_pin, sizeof(double) * n * W, cudaMemcpyHostToDevice, b->stream));
  CUDA_TRY(h, cudaMemcpyAsync(b->chains_d, b->h_pin_i, sizeof(int32_t) * n, cudaMemcpyHostToDevice, b->stream));
  const long long tot = (long long)n * (P + 1);
  batched_scatter_state_kernel<<<(int)std::min<long long>((tot + 255) / 256, 4 * b->sms), 256, 0, b->stream>>>(
      n, (int)P, b->ld, b->chains_d, b->theta_in, inv_metric ? 1 : 0, b->Q, b->Pm, b->Gd, b->IM, b->V);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  CUDA_TRY(h, cudaStreamSynchronize(b->stream));
  return B200GLM_OK;
}

int b200glm_leapfrog_batched(b200glm_handle* h, int32_t n, const int32_t* chains, const double* eps, double* q,
                             double* p, double* g, double* V, int32_t* status) {
  int rc = batch_check(h, n);
  if (rc) return rc;
  if (!eps) {
    h->set_error("null pointer argument");
    return B200GLM_INVALID;
  }
  Batch* b = h->batch;
  std::lock_guard<std::mutex> lk(b->mu);
  CUDA_TRY(h, cudaSetDevice(h->d.device));
  const size_t P = h->P;
  for (int i = 0; i < n; ++i) {
    const int c = chains ? chains[i] : i;
    if (c < 0 || c >= b->max_chains) {
      h->set_error("chain slot out of range");
      return B200GLM_INVALID;
    }
    b->h_pin_i[i] = c;
    b->h_pin[i] = eps[i];
  }
  CUDA_TRY(h, cudaMemcpyAsync(b->chains_d, b->h_pin_i, sizeof(int32_t) * n, cudaMemcpyHostToDevice, b->stream));
  CUDA_TRY(h, cudaMemcpyAsync(b->eps_d, b->h_pin, sizeof(double) * n, cudaMemcpyHostToDevice, b->stream));
  rc = enqueue_batched(h, n, MODE_LEAPFROG, 1, 1, b->chains_d, b->eps_d, 0.0, true);
  if (rc) return rc;
  double* hs = b->h_pin;
  const size_t W = 3 * P + 1;
  CUDA_TRY(h, cudaMemcpyAsync(hs, b->state_out, sizeof(double) * n * W, cudaMemcpyDeviceToHost, b->stream));
  CUDA_TRY(h, cudaStreamSynchronize(b->stream));
  for (int i = 0; i < n; ++i) {
    const double* r = hs + (size_t)i * W;
    if (q) std::memcpy(q + (size_t)i * P, r, sizeof(double) * P);
    if (p) std::memcpy(p + (size_t)i * P, r + P, sizeof(double) * P);
    if (g) std::memcpy(g + (size_t)i * P, r + 2 * P, sizeof(double) * P);
    if (V) V[i] = r[3 * P];
    if (status) status[i] = (h->bad_y || std::isinf(r[3 * P])) ? B200GLM_DOMAIN : B200GLM_OK;
  }
  return B200GLM_OK;
}

int b200glm_leapfrog_batched_async(b200glm_handle* h, int32_t n, double eps) {
  int rc = batch_check(h, n);
  if (rc) return rc;
  CUDA_TRY(h, cudaSetDevice(h->d.device));
  return enqueue_batched(h, n, MODE_LEAPFROG, 1, 1, nullptr, nullptr, eps, false);
}

int b200glm_batch_sync(b200glm_handle* h) {
  int rc = batch_check(h, 1);
  if (rc) return rc;
  CUDA_TRY(h, cudaStreamSynchronize(h->batch->stream));
  return B200GLM_OK;
}

void* b200glm_batch_stream(b200glm_handle* h) { return (h && h->batch) ? (void*)h->batch->stream : nullptr; }

// ---- device-side NUTS: nuts_tree.cuh / nuts_kernels.cuh behind the batched leapfrog ----
static NutsDeviceParams nuts_params(b200glm_handle* h) {
  Batch* b = h->batch;
  Batch::Nuts* u = b->nuts;
  NutsDeviceParams p;
  std::memset(&p, 0, sizeof(p));
  p.cfg = u->cfg;
  p.chains = u->chains;
  p.vec = u->vec;
  p.vstride = u->vstride;
  p.Q = b->Q;
  p.Pm = b->Pm;
  p.Gd = b->Gd;
  p.IM = b->IM;
  p.V = b->V;
  p.ld = (size_t)b->ld;
  p.eps_c = u->eps_c;
  p.normals = u->normals;
  p.unif = u->uniforms;
  p.status = u->status;
  p.draws = u->draws;
  p.metric = u->metric;
  return p;
}

int b200glm_nuts_reserve(b200glm_handle* h, int32_t n_chains, const b200glm_nuts_config* c) {
  if (!h) return B200GLM_INVALID;
  if (!c || n_chains < 1 || c->max_depth < 1 || c->max_depth > NUTS_DEPTH_CAP || c->num_warmup < 0 || c->num_samples < 0) {
    h->set_error("b200glm_nuts_reserve: needs n_chains >= 1, 1 <= max_depth <= 16, num_warmup / num_samples >= 0");
    return B200GLM_INVALID;
  }
  int rc = b200glm_batch_reserve(h, n_chains);
  if (rc) return rc;
  Batch* b = h->batch;
  std::lock_guard<std::mutex> lk(b->mu);
  CUDA_TRY(h, cudaSetDevice(h->d.device));
  CUDA_TRY(h, cudaStreamSynchronize(b->stream));
  free_nuts(b);
  Batch::Nuts* u = new Batch::Nuts();
  b->nuts = u;
  const size_t P = h->P, n = (size_t)n_chains;
  u->n_chains = n_chains;
  u->cfg.P = h->P;
  u->cfg.max_depth = c->max_depth;
  u->cfg.max_deltaH = c->max_deltaH;
  u->cfg.delta = c->delta;
  u->cfg.gamma = c->gamma;
  u->cfg.kappa = c->kappa;
  u->cfg.t0 = c->t0;
  u->cfg.w_num_warmup = c->w_num_warmup;
  u->cfg.w_init_buffer = c->w_init_buffer;
  u->cfg.w_term_buffer = c->w_term_buffer;
  u->cfg.w_base_window = c->w_base_window;
  u->cfg.w_size0 = c->w_size0;
  u->cfg.w_next0 = c->w_next0;
  u->cfg.num_warmup = c->num_warmup;
  u->cfg.num_samples = c->num_samples;
  u->cfg.stepsize_jitter = c->stepsize_jitter;
  u->vstride = nuts_vec_doubles(h->P, c->max_depth);
  u->use_graphs = !std::getenv("B200GLM_NO_GRAPH") && h->d.world <= 1;   // row shards: the round contains a collective
  CUDA_TRY(h, cudaMalloc(&u->chains, sizeof(NutsChain) * n));
  CUDA_TRY(h, cudaMemset(u->chains, 0, sizeof(NutsChain) * n));
  CUDA_TRY(h, cudaMalloc(&u->vec, sizeof(double) * u->vstride * n));
  CUDA_TRY(h, cudaMemset(u->vec, 0, sizeof(double) * u->vstride * n));
  CUDA_TRY(h, cudaMalloc(&u->eps_c, sizeof(double) * b->ld));
  CUDA_TRY(h, cudaMemset(u->eps_c, 0, sizeof(double) * b->ld));
  CUDA_TRY(h, cudaMalloc(&u->lanes_d, sizeof(int32_t) * b->ld));
  CUDA_TRY(h, cudaMallocHost(&u->normals, sizeof(double) * n * P));
  CUDA_TRY(h, cudaMallocHost(&u->uniforms, sizeof(double) * n * NUTS_UNIF_STRIDE));
  CUDA_TRY(h, cudaMallocHost(&u->draws, sizeof(double) * n * nuts_draw_doubles((int)P)));
  CUDA_TRY(h, cudaMallocHost(&u->metric, sizeof(double) * n * P));
  CUDA_TRY(h, cudaMallocHost(&u->status, sizeof(NutsStatus) * n));
  std::memset(u->normals, 0, sizeof(double) * n * P);
  std::memset(u->uniforms, 0, sizeof(double) * n * NUTS_UNIF_STRIDE);
  std::memset(u->draws, 0, sizeof(double) * n * nuts_draw_doubles((int)P));
  std::memset(u->metric, 0, sizeof(double) * n * P);
  std::memset(u->status, 0, sizeof(NutsStatus) * n);
  return B200GLM_OK;
}

static int nuts_check(b200glm_handle* h) {
  if (!h) return B200GLM_INVALID;
  if (!h->batch || !h->batch->nuts) {
    h->set_error("b200glm_nuts_reserve was not called");
    return B200GLM_INVALID;
  }
  return B200GLM_OK;
}

int b200glm_nuts_buffers(b200glm_handle* h, double** normals, double** uniforms, b200glm_nuts_status** status,
                         double** draws, double** metric) {
  int rc = nuts_check(h);
  if (rc) return rc;
  static_assert(sizeof(b200glm_nuts_status) == sizeof(NutsStatus), "b200glm_nuts_status mirrors NutsStatus");
  Batch::Nuts* u = h->batch->nuts;
  if (normals) *normals = u->normals;
  if (uniforms) *uniforms = u->uniforms;
  if (status) *status = reinterpret_cast<b200glm_nuts_status*>(u->status);
  if (draws) *draws = u->draws;
  if (metric) *metric = u->metric;
  return B200GLM_OK;
}

int b200glm_nuts_init_chain(b200glm_handle* h, int32_t chain, const double* q0, const double* inv_metric,
                            double stepsize) {
  int rc = nuts_check(h);
  if (rc) return rc;
  Batch* b = h->batch;
  Batch::Nuts* u = b->nuts;
  if (chain < 0 || chain >= u->n_chains || !q0 || !inv_metric) {
    h->set_error("b200glm_nuts_init_chain: chain out of range or null pointer");
    return B200GLM_INVALID;
  }
  std::lock_guard<std::mutex> lk(b->mu);
  CUDA_TRY(h, cudaSetDevice(h->d.device));
  const size_t P = h->P;
  std::memcpy(u->normals + (size_t)chain * P, q0, sizeof(double) * P);        // staged in the chain's own rows
  std::memcpy(u->metric + (size_t)chain * P, inv_metric, sizeof(double) * P);
  NutsDeviceParams p = nuts_params(h);
  p.init_chain = chain;
  p.stepsize = stepsize;
  nuts_init_kernel<<<1, 32, 0, b->stream>>>(p);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  CUDA_TRY(h, cudaStreamSynchronize(b->stream));
  return B200GLM_OK;
}

int b200glm_nuts_round(b200glm_handle* h, int32_t n, const int32_t* chains) {
  int rc = nuts_check(h);
  if (rc) return rc;
  rc = batch_check(h, n);
  if (rc) return rc;
  Batch* b = h->batch;
  Batch::Nuts* u = b->nuts;
  if (!chains) {
    h->set_error("null pointer argument");
    return B200GLM_INVALID;
  }
  std::lock_guard<std::mutex> lk(b->mu);
  CUDA_TRY(h, cudaSetDevice(h->d.device));
  bool any_begin = false;
  for (int i = 0; i < n; ++i) {
    if (chains[i] < 0 || chains[i] >= u->n_chains) {
      h->set_error("chain out of range");
      return B200GLM_INVALID;
    }
    any_begin = any_begin || u->status[chains[i]].need_normals != 0;   // the status of the previous round (pinned)
  }
  // the lane list only changes when a chain finishes: upload it then (the previous round has completed, so the
  // pinned staging block is free)
  if ((int)u->lanes_h.size() != n || std::memcmp(u->lanes_h.data(), chains, sizeof(int32_t) * n) != 0) {
    u->lanes_h.assign(chains, chains + n);
    std::memcpy(b->h_pin_i, chains, sizeof(int32_t) * n);
    CUDA_TRY(h, cudaMemcpyAsync(u->lanes_d, b->h_pin_i, sizeof(int32_t) * n, cudaMemcpyHostToDevice, b->stream));
  }
  NutsDeviceParams p = nuts_params(h);
  p.n_lanes = n;
  p.lanes = u->lanes_d;
  auto enqueue_round = [&](bool with_begin) -> int {
    if (with_begin) {   // chains that were waiting for normal variates (the caller has supplied them); others pass through
      nuts_begin_kernel<<<(n + 7) / 8, 256, 0, b->stream>>>(p);
      h->launches++;
    }
    int r = enqueue_batched(h, n, MODE_LEAPFROG, 1, 1, u->lanes_d, u->eps_c, 0.0, false, true);
    if (r) return r;
    nuts_step_kernel<<<(n + 7) / 8, 256, 0, b->stream>>>(p);
    h->launches++;
    return B200GLM_OK;
  };
  if (u->use_graphs) {
    auto it = u->graphs.find(n);
    if (it == u->graphs.end()) {
      const long long l0 = h->launches;
      cudaGraph_t graph = nullptr;
      cudaGraphExec_t exec = nullptr;
      bool ok = cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
      if (ok) {
        rc = enqueue_round(true);
        ok = cudaStreamEndCapture(b->stream, &graph) == cudaSuccess && rc == B200GLM_OK && graph != nullptr;
      }
      const int inside = (int)(h->launches - l0);
      h->launches = l0;
      if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
      if (graph) cudaGraphDestroy(graph);
      if (!ok) {            // capture not possible here: fall back to plain launches for the rest of the run
        cudaGetLastError();
        u->use_graphs = false;
      } else {
        it = u->graphs.emplace(n, std::make_pair(exec, inside)).first;
      }
    }
    if (u->use_graphs) {
      CUDA_TRY(h, cudaGraphLaunch(it->second.first, b->stream));
      h->launches += it->second.second;
      CUDA_TRY(h, cudaStreamSynchronize(b->stream));
      return B200GLM_OK;
    }
  }
  rc = enqueue_round(any_begin);
  if (rc) return rc;
  CUDA_TRY(h, cudaGetLastError());
  CUDA_TRY(h, cudaStreamSynchronize(b->stream));
  return B200GLM_OK;
}

int b200glm_append_rows(b200glm_handle* h, int64_t n, const double* X, int64_t ldx, const int32_t* y_int,
                        const double* y_real, const int32_t* trials) {
  if (!h) return B200GLM_INVALID;
  const b200glm_desc& d = h->d;
  if (!h->streamed || h->ready) {
    h->set_error("b200glm_append_rows: not a streamed handle, or already finalized");
    return B200GLM_INVALID;
  }
  if (n <= 0 || h->rows_appended + n > d.N) {
    h->set_error("b200glm_append_rows: chunk is empty or exceeds the rows reserved at b200glm_create");
    return B200GLM_INVALID;
  }
  if (h->rows_appended + n < d.N && n % h->panel_rows != 0) {
    h->set_error("b200glm_append_rows: every chunk but the last must be a multiple of " + std::to_string(h->panel_rows) + " rows");
    return B200GLM_INVALID;
  }
  if ((d.K > 0 && (!X || ldx < n)) || (d.family == B200GLM_NORMAL_ID ? !y_real : !y_int)
      || (d.family == B200GLM_BINOMIAL_LOGIT && !trials)) {
    h->set_error("b200glm_append_rows: null pointer or ldx < n");
    return B200GLM_INVALID;
  }
  CUDA_TRY(h, cudaSetDevice(d.device));
  cudaStream_t st = h->slots[0]->stream;
  const double* dX = X;
  const int32_t* dy = y_int;
  const double* dyr = y_real;
  const int32_t* dt = trials;
  void *tX = nullptr, *ty = nullptr, *tt = nullptr;
  int rc = B200GLM_OK;
  auto fail_cuda = [&](cudaError_t e, const char* what) {
    h->set_error(std::string(what) + ": " + cudaGetErrorString(e));
    rc = B200GLM_CUDA;
  };
  if (!d.data_on_device) {   // host chunk: stage it (the chunk, not the matrix, is what is resident twice)
    cudaError_t e;
    if (d.K > 0) {
      if ((e = cudaMalloc(&tX, sizeof(double) * (size_t)n * d.K)) != cudaSuccess) fail_cuda(e, "cudaMalloc(chunk)");
      else if ((e = cudaMemcpy2D(tX, sizeof(double) * n, X, sizeof(double) * ldx, sizeof(double) * n, d.K,
                                 cudaMemcpyHostToDevice)) != cudaSuccess) fail_cuda(e, "cudaMemcpy2D(chunk)");
      dX = static_cast<const double*>(tX);
      ldx = n;
    }
    const size_t ybytes = (d.family == B200GLM_NORMAL_ID ? sizeof(double) : sizeof(int32_t)) * (size_t)n;
    if (rc == B200GLM_OK && (e = cudaMalloc(&ty, ybytes)) != cudaSuccess) fail_cuda(e, "cudaMalloc(y)");
    if (rc == B200GLM_OK && (e = cudaMemcpy(ty, d.family == B200GLM_NORMAL_ID ? (const void*)y_real : (const void*)y_int,
                                            ybytes, cudaMemcpyHostToDevice)) != cudaSuccess) fail_cuda(e, "cudaMemcpy(y)");
    dy = static_cast<const int32_t*>(ty);
    dyr = static_cast<const double*>(ty);
    if (rc == B200GLM_OK && trials) {
      if ((e = cudaMalloc(&tt, sizeof(int32_t) * (size_t)n)) != cudaSuccess) fail_cuda(e, "cudaMalloc(trials)");
      else if ((e = cudaMemcpy(tt, trials, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice)) != cudaSuccess)
        fail_cuda(e, "cudaMemcpy(trials)");
      dt = static_cast<const int32_t*>(tt);
    }
  }
  if (rc == B200GLM_OK) {
    const long long r0 = h->rows_appended;
    // y / trials pointers are chunk-relative: relayout_kernel indexes them with the GLOBAL row, so shift them back
    const int32_t* y_g = d.family == B200GLM_NORMAL_ID ? nullptr : dy - r0;
    const double* yr_g = d.family == B200GLM_NORMAL_ID ? dyr - r0 : nullptr;
    const int32_t* t_g = dt ? dt - r0 : nullptr;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, d.device);
    relayout_kernel<<<4 * prop.multiProcessorCount, 256, 0, st>>>(dX, ldx, r0, y_g, yr_g, nullptr, t_g, nullptr, r0, n,
                                                                  d.N, d.K, h->C, 0, h->Cpad, h->panels, h->panel_rows,
                                                                  h->Cpad, h->wide ? 0 : 1);
    if (d.family != B200GLM_NORMAL_ID) {   // data checks + the propto=false constant of this chunk
      const int nb = 296;
      double* d_stats = nullptr;
      cudaError_t e = cudaMalloc(&d_stats, sizeof(double) * 2 * nb);
      if (e != cudaSuccess) fail_cuda(e, "cudaMalloc(stats)");
      else {
        y_stats_kernel<<<nb, 256, 0, st>>>(dy, dt, n, d.family, d_stats, 0);
        std::vector<double> hs(2 * nb);
        if ((e = cudaMemcpyAsync(hs.data(), d_stats, sizeof(double) * 2 * nb, cudaMemcpyDeviceToHost, st)) != cudaSuccess
            || (e = cudaStreamSynchronize(st)) != cudaSuccess)
          fail_cuda(e, "y_stats");
        for (int i = 0; i < nb && rc == B200GLM_OK; ++i) {
          h->bad_count += hs[2 * i];
          h->lgamma_sum += hs[2 * i + 1];
        }
        cudaFree(d_stats);
      }
    }
    cudaError_t e = cudaStreamSynchronize(st);   // the caller may reuse the chunk buffer as soon as this returns
    if (rc == B200GLM_OK && e != cudaSuccess) fail_cuda(e, "relayout");
    if (rc == B200GLM_OK && (e = cudaGetLastError()) != cudaSuccess) fail_cuda(e, "relayout");
  }
  cudaFree(tX);
  cudaFree(ty);
  cudaFree(tt);
  if (rc == B200GLM_OK) {
    h->rows_appended += n;
    h->launches++;
  }
  return rc;
}

int b200glm_finalize(b200glm_handle* h) {
  if (!h) return B200GLM_INVALID;
  if (!h->streamed) return B200GLM_OK;
  if (h->rows_appended != h->d.N) {
    h->set_error("b200glm_finalize: " + std::to_string(h->rows_appended) + " of " + std::to_string((long long)h->d.N)
                 + " rows appended");
    return B200GLM_INVALID;
  }
  h->bad_y = h->bad_y_local = h->bad_count > 0;
  h->lgamma_sum_total = h->lgamma_sum;
  h->ready = true;
  return B200GLM_OK;
}

int b200glm_measure_peaks(int32_t device, double* read_gbs, double* dmma_tflops) {
  if (cudaSetDevice(device) != cudaSuccess) return B200GLM_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return B200GLM_CUDA;
  const int sms = prop.multiProcessorCount;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int rc = B200GLM_OK;
  if (read_gbs) {
    const size_t bytes = (size_t)4 << 30;     // 4 GiB >> 126 MB L2
    double2* buf = nullptr;
    double* out = nullptr;
    if (cudaMalloc(&buf, bytes) != cudaSuccess || cudaMalloc(&out, sizeof(double) * sms * 4) != cudaSuccess) {
      cudaFree(buf);
      rc = B200GLM_CUDA;
    } else {
      cudaMemset(buf, 0, bytes);
      float best = 1e30f;
      for (int i = 0; i < 8; ++i) {
        cudaEventRecord(e0);
        read_stream_kernel<<<sms * 4, 512>>>(buf, bytes / 16, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (i >= 2 && ms < best) best = ms;
      }
      *read_gbs = (double)bytes / (best * 1e-3) / 1e9;
      if (cudaGetLastError() != cudaSuccess) rc = B200GLM_CUDA;
    }
    cudaFree(buf);
    cudaFree(out);
  }
  if (dmma_tflops && rc == B200GLM_OK) {
    double* out = nullptr;
    const int grid = sms * 4, iters = 8000;
    if (cudaMalloc(&out, sizeof(double) * grid * 256) != cudaSuccess) {
      rc = B200GLM_CUDA;
    } else {
      float best = 1e30f;
      for (int i = 0; i < 4; ++i) {
        cudaEventRecord(e0);
        dmma_peak_kernel<<<grid, 256>>>(out, iters, 0.999, 0.001);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (i >= 1 && ms < best) best = ms;
      }
      // one m8n8k4 DMMA = 8*8*4*2 = 512 flops per warp; 8 warps per CTA, 16 accumulator tiles per iteration
      *dmma_tflops = 512.0 * 16 * iters * 8.0 * grid / (best * 1e-3) / 1e12;
      if (cudaGetLastError() != cudaSuccess) rc = B200GLM_CUDA;
      cudaFree(out);
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}

int b200glm_timeline_enable(b200glm_handle* h, int32_t slot, int32_t on) {
  int rc = validate_slot(h, slot);
  if (rc) return rc;
  Slot* s = h->slots[slot];
  std::lock_guard<std::recursive_mutex> lk(s->mu);
  CUDA_TRY(h, cudaSetDevice(h->d.device));
  CUDA_TRY(h, cudaStreamSynchronize(s->stream));
  if (on && !s->tl) {
    CUDA_TRY(h, cudaMalloc(&s->tl, sizeof(unsigned long long) * 16 * (h->grid + 2)));
    CUDA_TRY(h, cudaMemset(s->tl, 0, sizeof(unsigned long long) * 16 * (h->grid + 2)));
  } else if (!on && s->tl) {
    cudaFree(s->tl);
    s->tl = nullptr;
  }
  return B200GLM_OK;
}

int b200glm_timeline_read(b200glm_handle* h, int32_t slot, uint64_t* out, int32_t* rows) {
  int rc = validate_slot(h, slot);
  if (rc) return rc;
  Slot* s = h->slots[slot];
  if (rows) *rows = h->grid + 2;
  if (!out) return B200GLM_OK;
  if (!s->tl) {
    h->set_error("b200glm_timeline_enable was not called for this slot");
    return B200GLM_INVALID;
  }
  CUDA_TRY(h, cudaSetDevice(h->d.device));
  CUDA_TRY(h, cudaStreamSynchronize(s->stream));
  CUDA_TRY(h, cudaMemcpy(out, s->tl, sizeof(unsigned long long) * 16 * (h->grid + 2), cudaMemcpyDeviceToHost));
  return B200GLM_OK;
}

}  // extern "C"

extern "C" {

int b200glm_peer_export(b200glm_handle* h, void* ipc_handle_64) {
  if (!h || !ipc_handle_64) return B200GLM_INVALID;
  if (h->d.world < 2 || h->d.world > MAX_PEERS) {
    h->set_error("peer mailboxes need 2 <= world <= 8");
    return B200GLM_INVALID;
  }
  CUDA_TRY(h, cudaSetDevice(h->d.device));
  if (!h->mbox) {
    h->peer_stride = (h->P + 2 + 15) & ~15;  // payload, rounded up to whole 128-byte lines
    const size_t n = (size_t)h->slots.size() * PEER_BUFS * h->d.world * h->peer_stride;
    CUDA_TRY(h, cudaMalloc(&h->mbox, sizeof(double) * n));
    CUDA_TRY(h, cudaMemset(h->mbox, 0xFF, sizeof(double) * n));   // every word = PEER_EMPTY
    CUDA_TRY(h, cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t ih;
  CUDA_TRY(h, cudaIpcGetMemHandle(&ih, h->mbox));
  static_assert(sizeof(ih) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::memcpy(ipc_handle_64, &ih, sizeof(ih));
  return B200GLM_OK;
}

int b200glm_peer_connect(b200glm_handle* h, const void* all_handles, int32_t world) {
  if (!h || !all_handles) return B200GLM_INVALID;
  if (world != h->d.world || !h->mbox) {
    h->set_error("b200glm_peer_connect: world differs from b200glm_create, or b200glm_peer_export was not called");
    return B200GLM_INVALID;
  }
  CUDA_TRY(h, cudaSetDevice(h->d.device));
  for (int r = 0; r < world; ++r) {
    if (r == h->d.rank) {
      h->peer_mbox[r] = h->mbox;
      continue;
    }
    cudaIpcMemHandle_t ih;
    std::memcpy(&ih, static_cast<const char*>(all_handles) + (size_t)r * sizeof(ih), sizeof(ih));
    void* ptr = nullptr;
    CUDA_TRY(h, cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess));
    h->peer_mbox[r] = static_cast<double*>(ptr);
  }
  // the propto=false poisson constant is a sum over all shards: exchanged through the same mailboxes
  // by the first evaluation is not possible (it is a create-time constant), so the caller passes it:
  h->peer_on = true;
  return B200GLM_OK;
}

int b200glm_set_lgamma_sum_total(b200glm_handle* h, double total) {
  if (!h) return B200GLM_INVALID;
  h->lgamma_sum_total = total;
  return B200GLM_OK;
}

double b200glm_lgamma_sum_local(const b200glm_handle* h) { return h ? h->lgamma_sum : 0.0; }

int b200glm_shard_constants_local(const b200glm_handle* h, double out[2]) {
  if (!h || !out) return B200GLM_INVALID;
  out[0] = h->lgamma_sum;
  out[1] = h->bad_y_local ? 1.0 : 0.0;
  return B200GLM_OK;
}

int b200glm_set_shard_constants_total(b200glm_handle* h, const double total[2]) {
  if (!h || !total) return B200GLM_INVALID;
  h->lgamma_sum_total = total[0];
  h->bad_y = total[1] > 0.0;
  return B200GLM_OK;
}

int b200glm_comm_unique_id(void* unique_id_128) {
  if (!unique_id_128 || !nccl().ok) return B200GLM_CUDA;
  ncclUniqueId id;
  if (nccl().GetUniqueId(&id) != 0) return B200GLM_CUDA;
  std::memcpy(unique_id_128, &id, sizeof(id));
  return B200GLM_OK;
}

int b200glm_comm_init(b200glm_handle* h, const void* unique_id_128, int32_t rank, int32_t world) {
  if (!h || !unique_id_128) return B200GLM_INVALID;
  if (!nccl().ok) {
    h->set_error("libnccl.so.2 could not be loaded");
    return B200GLM_CUDA;
  }
  if (rank != h->d.rank || world != h->d.world) {
    h->set_error("rank/world differ from the ones given to b200glm_create");
    return B200GLM_INVALID;
  }
  CUDA_TRY(h, cudaSetDevice(h->d.device));
  ncclUniqueId id;
  std::memcpy(&id, unique_id_128, sizeof(id));
  ncclResult_t r = nccl().CommInitRank(&h->comm, world, id, rank);
  if (r != 0) {
    h->set_error(std::string("ncclCommInitRank: ") + (nccl().GetErrorString ? nccl().GetErrorString(r) : "?"));
    return B200GLM_CUDA;
  }
  // per-shard create-time constants are sums over all shards: the propto=false constant of poisson_log /
  // binomial_logit / neg_binomial_2_log (0 for the other families) and the count of out-of-range y, so that
  // every rank reports the same data-check status
  {
    double hv[2] = {h->lgamma_sum, h->bad_y_local ? 1.0 : 0.0};
    double* dv;
    CUDA_TRY(h, cudaMalloc(&dv, sizeof(hv)));
    CUDA_TRY(h, cudaMemcpy(dv, hv, sizeof(hv), cudaMemcpyHostToDevice));
    cudaStream_t st = h->slots[0]->stream;
    if (nccl().AllReduce(dv, dv, 2, ncclFloat64, ncclSum, h->comm, st) != 0) {
      cudaFree(dv);
      h->set_error("ncclAllReduce(shard constants) failed");
      return B200GLM_CUDA;
    }
    CUDA_TRY(h, cudaMemcpyAsync(hv, dv, sizeof(hv), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    cudaFree(dv);
    h->lgamma_sum_total = hv[0];
    h->bad_y = hv[1] > 0.0;
  }
  return B200GLM_OK;
}

}  // extern "C"
