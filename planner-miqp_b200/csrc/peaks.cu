// peaks.cu -- FP64 FMA throughput micro-benchmark.
//
// MEASURED_PEAKS.json (driver-written) holds the HBM and bf16 tensor peaks of this pool's
// B200s but no FP64 figure; the node kernel is double-precision FMA work, so the roofline
// denominator for it is measured here, live, on the same device and in the same process as
// the benchmark: every thread runs 8 independent dependent-chains of DFMA, 4 CTAs of 256
// threads per SM, long enough to reach steady clocks.
#include <cuda_runtime.h>

namespace miqp {

__global__ void __launch_bounds__(256) fp64_fma_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0;
  double x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
#pragma unroll 1
  for (int k = 0; k < iters; ++k) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// returns TFLOP/s (2 flops per FMA), best of `reps`
double measure_fp64_tflops(int num_sms, cudaStream_t st, int reps) {
  const int ctas = num_sms * 4, threads = 256, iters = 4096;
  double *d = nullptr;
  if (cudaMalloc(&d, sizeof(double) * ctas * threads) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int r = 0; r < reps + 1; ++r) {
    cudaEventRecord(e0, st);
    fp64_fma_kernel<<<ctas, threads, 0, st>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8.0 * 16.0 * iters * (double)ctas * threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (r > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  return best;
}

}  // namespace miqp
