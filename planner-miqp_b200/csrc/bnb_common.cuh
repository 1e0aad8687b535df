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
