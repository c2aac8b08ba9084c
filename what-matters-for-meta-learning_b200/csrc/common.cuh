// Shared helpers for libb200np (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "b200np.h"

#define B200NP_VERSION 100

namespace b200np {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs (grid sizing default)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Kernel-launch counter (b200np_launch_count): every API notes how many kernels it enqueued.
extern std::atomic<long long> g_launches;
inline int launch_status(int kernels = 1) {
  g_launches.fetch_add(kernels, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? B200NP_OK : B200NP_E_LAUNCH;
}

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline long long ceil_div(long long a, long long b) { return (a + b - 1) / b; }

// grid size for a grid-stride elementwise kernel: a few waves over 148 SMs
inline int ew_grid(long long work_items, int block) {
  long long g = ceil_div(work_items, block);
  long long cap = (long long)kNumSMs * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; every thread gets the result.  `red` = shared float[32].
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

// out[map(j)] = sum_{p < nparts} part[p * n + j], j < n.  Deterministic (fixed order), coalesced, 8 partial
// lanes per output.  map: 0 = identity; 1 = weight-gradient permutation (t, co, ci) -> (co, ci, t) with
// a = ntaps, b = Cout, c = Cin (an extra tap index a, if present, goes to out2[co][ci]);  2 = small-conv partials [co][KGp] -> dw[co][KK] / db[co] (a = KK, b = KGp).
struct ReduceMap { int kind, a, b, c; float* out2; };
int launch_reduce_partials(const float* part, float* out, int nparts, int n, ReduceMap map, cudaStream_t st);

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace b200np
