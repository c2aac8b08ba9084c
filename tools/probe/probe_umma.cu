// Hardware probe (diagnostic, not product): does a K-major SWIZZLE_128B UMMA operand work when
//  (a) its start address is offset by `row_off` 128-byte rows inside a 1024-byte swizzle atom, and
//  (b) the stride between 8-row groups (SBO) is `sbo_rows`*128 B, not a multiple of 1024 B,
// if data is stored with the absolute-address swizzle chunk ^ ((addr >> 7) & 7)?
// mode 0: base_offset field = 0; mode 1: base_offset = (start >> 7) & 7.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ buf, int buf_rows, const float* __restrict__ B,
                                                    float* __restrict__ D, int row_off, int sbo_rows, int mode) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_s = smem;                    // buf_rows x 128 B
  uint8_t* b_s = smem + 48 * 1024;        // 64 x 128 B
  uint64_t* bar = (uint64_t*)(smem + 60 * 1024);
  uint32_t* slot = (uint32_t*)(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < buf_rows * 8; i += 128) {
    int r = i >> 3, c = i & 7;
    uint32_t addr = r * 128;
    uint32_t off = addr + ((c ^ ((addr >> 7) & 7)) << 4);
    *(float4*)(a_s + off) = *(const float4*)(buf + r * 32 + c * 4);
  }
  for (int i = tid; i < 64 * 8; i += 128) {
    int r = i >> 3, c = i & 7;
    uint32_t off = (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4);
    *(float4*)(b_s + off) = *(const float4*)(B + r * 32 + c * 4);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tmem = *slot;
  if (tid == 0) {
    uint32_t sa = smem_u32(a_s) + row_off * 128, sb = smem_u32(b_s);
    auto desc = [&](uint32_t addr, uint32_t sbo_bytes, uint32_t bo) {
      uint64_t d = 0;
      d |= (uint64_t)((addr & 0x3FFFF) >> 4);
      d |= (uint64_t)1 << 16;
      d |= (uint64_t)(sbo_bytes >> 4) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)(bo & 7) << 49;
      d |= (uint64_t)2 << 61;
      return d;
    };
    uint32_t bo = mode ? ((sa >> 7) & 7) : 0;
    uint64_t ad = desc(sa, sbo_rows * 128, bo), bd = desc(sb, 1024, 0);
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64 >> 3) << 17) | ((128 >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      uint32_t acc = k != 0;
      asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(tmem),
                   "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(acc)
                   : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
    if (++spins > (1u << 22)) __trap();
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t r[32];
  for (int half = 0; half < 2; ++half) {
    uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + half * 32;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
        "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(ta));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[tid * 64 + half * 32 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

extern "C" int probe_umma(const float* buf, int buf_rows, const float* B, float* D, int row_off, int sbo_rows, int mode) {
  size_t smem = 62 * 1024 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem>>>(buf, buf_rows, B, D, row_off, sbo_rows, mode);
  return (int)cudaDeviceSynchronize();
}
