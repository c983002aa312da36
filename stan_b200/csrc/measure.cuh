// measure.cuh -- roofline denominators measured on the device a bench run uses, in the same process and under
// the same clocks as the kernels they bound (bench.py prints them next to MEASURED_PEAKS.json's copy figure):
//   read_stream_kernel   HBM bandwidth of a READ-ONLY stream: the GLM gradient reads X once and writes nothing,
//                        while the driver's hbm_gbs is a read+write copy -- a read-only stream can beat it, which
//                        is why roofline.frac against the copy figure may exceed 1
//   dmma_peak_kernel     fp64 tensor-pipe peak: a register-resident mma.sync.m8n8k4.f64 loop (SASS DMMA.8x8x4), the
//                        denominator of the batched-chains kernel (tcgen05 has no fp64 kind on sm_100a)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b200glm {

__global__ void __launch_bounds__(512) read_stream_kernel(const double2* __restrict__ src, size_t n16, double* out) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {
    const double2 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride),
                  d = __ldcs(src + i + 3 * stride);
    s0 += a.x + a.y;
    s1 += b.x + b.y;
    s2 += c.x + c.y;
    s3 += d.x + d.y;
  }
  for (; i < n16; i += stride) {
    const double2 a = __ldcs(src + i);
    s0 += a.x + a.y;
  }
  const double s = (s0 + s1) + (s2 + s3);
  if (s == 1.2345e-300) out[blockIdx.x] = s;   // keeps the loads alive, (practically) never taken
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a, double b) {
  double c0[16], c1[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    c0[i] = threadIdx.x;
    c1[i] = i;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace b200glm
