// Single-CTA vector helpers.  Every vector on the hot path has at most m + n + p <= ~60k entries
// (256-480 KB), so one 1024-thread CTA streams it in a few microseconds; doing so keeps every
// reduction a fixed tree (deterministic, no atomics) and lets one kernel fuse a whole
// "evaluate + reduce + decide" stage of the Newton step.
#pragma once
#include "common.cuh"

namespace cvxb {

constexpr int VT = 1024;   // threads of a vector kernel

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_or(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// All threads of the CTA must call; returns the total to every thread.  `buf` >= 33 doubles.
__device__ __forceinline__ double block_sum(double v, double* buf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) buf[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = lane < nw ? buf[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) buf[32] = t;
  }
  __syncthreads();
  return buf[32];
}
__device__ __forceinline__ double block_min(double v, double* buf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_min(v);
  __syncthreads();
  if (lane == 0) buf[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = lane < nw ? buf[lane] : 1e308;
    t = warp_min(t);
    if (lane == 0) buf[32] = t;
  }
  __syncthreads();
  return buf[32];
}
__device__ __forceinline__ int block_or(int v, int* ibuf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_or(v);
  __syncthreads();
  if (lane == 0) ibuf[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int t = lane < nw ? ibuf[lane] : 0;
    t = warp_or(t);
    if (lane == 0) ibuf[32] = t;
  }
  __syncthreads();
  return ibuf[32];
}

// MatrixUtils.relativeSize (MatrixUtils.scala:437-442) from the two norms
__device__ __forceinline__ double relative_size(double norm_a, double norm_b, double tol) {
  double f = (norm_b < tol) ? tol : tol + norm_b;
  return norm_a / f;
}

}  // namespace cvxb
