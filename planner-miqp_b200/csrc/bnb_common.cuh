// bnb_common.cuh -- small device helpers shared by the node kernels (bnb.cu, bnb_multi.cu).
#pragma once
#include "dev_problem.cuh"

namespace miqp {

__device__ __forceinline__ void atomic_min_double(double *addr, double v) {
  unsigned long long *a = reinterpret_cast<unsigned long long *>(addr);
  unsigned long long old = *a;
  while (__longlong_as_double((long long)old) > v) {
    unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}

// Takes the spin lock `lk` for the calling warp: lane 0 tries, the outcome is broadcast, and all 32 lanes loop together until
// it succeeds.  (A single lane spinning on atomicCAS while the other 31 lanes of its warp wait at a warp-collective or a CTA
// barrier produced intermittent "illegal instruction" faults and hangs on sm_100a under contention -- not reproducible under
// compute-sanitizer; profiles/r2a_scheduler_experiments.md.  Keeping the warp converged while it waits removed them.)
__device__ __forceinline__ void warp_lock(int *lk, int lane) {
  unsigned ns = 20;
  for (;;) {
    int got = 0;
    if (lane == 0) got = (atomicCAS(lk, 0, 1) == 0);
    got = __shfl_sync(0xffffffffu, got, 0);
    if (got) break;
    __nanosleep(ns);
    if (ns < 320) ns *= 2;
  }
  __threadfence();
}
__device__ __forceinline__ void warp_unlock(int *lk, int lane) {   // every lane's writes are fenced before lane 0 releases
  __threadfence();
  __syncwarp();
  if (lane == 0) atomicExch(lk, 0);
  __syncwarp();
}

__device__ __forceinline__ unsigned long long ordered_bits(double v) {
  long long b = __double_as_longlong(v);
  unsigned long long u = (unsigned long long)b;
  return (b < 0) ? ~u : (u | 0x8000000000000000ULL);
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

}  // namespace miqp
