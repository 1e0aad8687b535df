// Dependent-issue latencies on sm_100a that shape the node kernel: DFMA chain, shared-memory load->use,
// MUFU.RCP64H, __syncwarp, warp shuffle of a double.  One warp per SM sub-partition, clock64 deltas.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, int n) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  double a = out[threadIdx.x], b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) a = fma(a, b, c);
  long long t1 = clock64();
  double s = a;
  int idx = threadIdx.x;
#pragma unroll 16
  for (int i = 0; i < n; ++i) { double v = sm[idx]; idx = (int)(v) + ((idx + 33) & 1023) - 1; s += v; }   // load -> address dependence
  long long t2 = clock64();
  double r = a;
#pragma unroll 16
  for (int i = 0; i < n; ++i) { double q; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(r)); r = q + 1.5; }
  long long t3 = clock64();
  double h = r;
#pragma unroll 16
  for (int i = 0; i < n; ++i) h = __shfl_xor_sync(0xffffffffu, h, 1) + 1.0;
  long long t4 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) { __syncwarp(); sm[threadIdx.x] = h; __syncwarp(); h = sm[(threadIdx.x + 1) & 31] + 1.0; }
  long long t5 = clock64();
  double d = h;
#pragma unroll 16
  for (int i = 0; i < n; ++i) d = d / (1.0 + 1e-9 * d);
  long long t6 = clock64();
  out[threadIdx.x] = a + s + r + h + d;
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; }
}
int main() {
  double *out; long long *cyc, h[6];
  cudaMalloc(&out, 1024 * 8); cudaMemset(out, 0, 1024 * 8); cudaMalloc(&cyc, 64);
  const int n = 4096;
  for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(out, cyc, n);
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  const char *names[6] = {"DFMA dependent", "LDS -> address -> LDS (+cvt, add)", "MUFU.RCP64H + DADD", "SHFL(double) + DADD", "syncwarp+STS+syncwarp+LDS+DADD", "IEEE DDIV (+DFMA)"};
  for (int i = 0; i < 6; ++i) printf("%-40s %.1f cycles\n", names[i], (double)h[i] / n);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
