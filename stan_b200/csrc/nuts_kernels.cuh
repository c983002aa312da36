// Kernels around the per-chain NUTS state machine of nuts_tree.cuh: one WARP per chain, features strided
// over the lanes, dot products by shuffles; 8 chains per CTA.  A round of the device-side sampler is
//   nuts_begin_kernel -> batched leapfrog (batched_begin / glm_batched / batched_reduce / batched_finish) -> nuts_step_kernel
// on the batch stream; the chain slots (q, p, g, V, inverse metric; feature-major) are the batched kernels' own, the tree
// state is chain-major.  Host-visible outputs (status, draws, adapted metric) and inputs (normal / uniform variates) are
// pinned host memory the kernels access directly: a few hundred bytes per chain and round, P doubles per momentum refresh.
#pragma once
#include "nuts_tree.cuh"

namespace b200glm {

struct NutsWarp {
  static __device__ __forceinline__ int lane() { return threadIdx.x & 31; }
  static __device__ __forceinline__ int lanes() { return 32; }
  static __device__ __forceinline__ double sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
};

struct NutsDeviceParams {
  NutsConfig cfg;
  int n_lanes;
  const int32_t* lanes;       // [n_lanes] chain ids (device)
  NutsChain* chains;          // [C]
  double* vec;                // [C][vstride]
  size_t vstride;
  double *Q, *Pm, *Gd, *IM, *V;   // chain slots, [P][ld] / [ld]
  size_t ld;
  double* eps_c;              // [ld] step of every chain's next leapfrog lane (read by the batched kernels)
  const double* normals;      // pinned [C][P]
  const double* unif;         // pinned [C][NUTS_UNIF_STRIDE]
  NutsStatus* status;         // pinned [C]
  double* draws;              // pinned [C][nuts_draw_doubles(P)]
  double* metric;             // pinned [C][P]
  double stepsize;            // nuts_init_kernel only
  int init_chain;             // nuts_init_kernel only
};

__device__ __forceinline__ NutsSlot nuts_slot(const NutsDeviceParams& p, int c) {
  return NutsSlot{p.Q, p.Pm, p.Gd, p.IM, p.V, p.ld, c};
}

// one chain: q0 (staged in its row of `normals`) and the inverse metric (its row of `metric`) into the slot, state machine reset
__global__ void __launch_bounds__(32) nuts_init_kernel(const NutsDeviceParams p) {
  const int c = p.init_chain, P = p.cfg.P, lane = threadIdx.x;
  const NutsSlot s = nuts_slot(p, c);
  for (int k = lane; k < P; k += 32) {
    s.q(k) = p.normals[(size_t)c * P + k];
    s.p(k) = 0.0;
    s.g(k) = 0.0;
    s.im(k) = p.metric[(size_t)c * P + k];
  }
  if (lane == 0) s.v() = 0.0;
  NutsChain ch;
  nuts_chain_init<NutsWarp>(p.cfg, ch, p.vec + (size_t)c * p.vstride, s, p.stepsize);
  __syncwarp();
  if (lane == 0) {
    p.chains[c] = ch;
    p.eps_c[c] = ch.lane_eps;
    nuts_publish(ch, p.status[c]);
  }
}

// chains whose normal variates have arrived: start the transition / the init_stepsize iteration
__global__ void __launch_bounds__(256) nuts_begin_kernel(const NutsDeviceParams p) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= p.n_lanes) return;
  const int c = p.lanes[i], P = p.cfg.P;
  NutsChain ch = p.chains[c];
  if (!ch.need_normals || !(ch.phase == NPH_SS_FIRST || ch.phase == NPH_SS_LOOP || ch.phase == NPH_TREE)) return;
  nuts_begin<NutsWarp>(p.cfg, ch, p.vec + (size_t)c * p.vstride, nuts_slot(p, c), p.normals + (size_t)c * P,
                       p.unif + (size_t)c * NUTS_UNIF_STRIDE);
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    p.chains[c] = ch;
    p.eps_c[c] = ch.lane_eps;
  }
}

// everything the reference does between two leapfrog steps of a chain, for every lane of the round
__global__ void __launch_bounds__(256) nuts_step_kernel(const NutsDeviceParams p) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= p.n_lanes) return;
  const int c = p.lanes[i], P = p.cfg.P;
  NutsChain ch = p.chains[c];
  nuts_after_leapfrog<NutsWarp>(p.cfg, ch, p.vec + (size_t)c * p.vstride, nuts_slot(p, c),
                                p.unif + (size_t)c * NUTS_UNIF_STRIDE, p.draws + (size_t)c * nuts_draw_doubles(P),
                                p.metric + (size_t)c * P);
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    p.chains[c] = ch;
    p.eps_c[c] = ch.lane_eps;
    nuts_publish(ch, p.status[c]);
  }
}

}  // namespace b200glm
