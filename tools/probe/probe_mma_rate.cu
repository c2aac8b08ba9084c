// Hardware probe: cycles per tcgen05.mma.kind::tf32 (M=128 or 64, K=8) as a function of N and of the A operand's
// row-group stride (SBO), both operands in shared memory.  One CTA, one issuing thread, back-to-back MMAs.
// (M = 64 added at the end of round 1, not yet run: is a single-tap MMA cheaper than the unused half of a tap pair in
// csrc/tapwgrad_halo.cu role 1?)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128) rate_kernel(long long* out, int N, int sbo_rows, int iters, int M) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + 120 * 1024);
  uint32_t* slot = (uint32_t*)(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 120 * 1024 / 4; i += 128) ((float*)smem)[i] = 1.0f;
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(slot))); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tmem = *slot;
  if (tid == 0) {
    uint32_t a16 = (smem_u32(smem) & 0x3FFFF) >> 4, b16 = (smem_u32(smem + 64 * 1024) & 0x3FFFF) >> 4;
    uint32_t a_hi = ((uint32_t)(sbo_rows * 128) >> 4) | (1u << 14) | (2u << 29), b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t al = (a16 + 2 * k) | (1u << 16), bl = (b16 + 2 * k) | (1u << 16), acc = (it | k) != 0;
        asm volatile("{.reg .pred p; .reg .b64 da, db; mov.b64 da, {%1, %2}; mov.b64 db, {%3, %4}; setp.ne.b32 p, %6, 0; "
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;}" ::"r"(tmem), "r"(al), "r"(a_hi), "r"(bl), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
      }
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
    long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 2 * sizeof(long long));
  size_t smem = 122 * 1024 + 1024;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int iters = 2000;
  for (int M : {128, 64}) for (int ncta : {1, 148}) for (int N : {64, 128, 256}) for (int sbo : {8, 10}) {
    if (M == 64 && sbo != 8) continue;
    rate_kernel<<<ncta, 128, smem>>>(d, N, sbo, iters, M);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("M=%3d ctas=%3d N=%3d sbo_rows=%2d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (%s)\n", M, ncta, N, sbo, (double)h[0] / (4.0 * iters), (double)h[1] / (4.0 * iters), cudaGetErrorString(e));
  }
  return 0;
}
