// fp64_peak.cu -- measures the fp64 roofline denominators on this B200: DFMA (FP64 pipe) and
// DMMA (mma.sync.m8n8k4.f64, the only fp64 tensor path on sm_100a) register-resident loops.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters, double a, double b) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c0[i] = threadIdx.x; c1[i] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  const int iters = 20000;
  for (int cps = 1; cps <= 8; cps *= 2) {
    const int grid = sms * cps;
    float ms = time_ms([&] { dfma_loop<<<grid, 256>>>(out, iters, 0.999, 0.001); });
    double fl = 2.0 * 16 * iters * 256.0 * grid;
    printf("DFMA  ctas/sm=%d  %.3f ms  %.2f TFLOP/s\n", cps, ms, fl / ms / 1e9);
    ms = time_ms([&] { dmma_loop<8><<<grid, 256>>>(out, iters, 0.999, 0.001); });
    fl = 512.0 * 8 * iters * 8.0 * grid;
    printf("DMMA8 ctas/sm=%d  %.3f ms  %.2f TFLOP/s\n", cps, ms, fl / ms / 1e9);
    ms = time_ms([&] { dmma_loop<16><<<grid, 256>>>(out, iters, 0.999, 0.001); });
    fl = 512.0 * 16 * iters * 8.0 * grid;
    printf("DMMA16 ctas/sm=%d  %.3f ms  %.2f TFLOP/s\n", cps, ms, fl / ms / 1e9);
  }
  return 0;
}
