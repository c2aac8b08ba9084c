// Hardware probe (diagnostic, not product): MN-major SWIZZLE_128B_BASE32B tf32 operands (the weight gradient's layout:
// atom = 4 pixel rows x 128 B, 32-byte units permuted by unit ^ (row & 3), SBO = next group of 4 pixels, LBO = next
// 32-channel block).  Question for the halo formulation of the weight gradient: if the data is
// stored with the swizzle taken from ABSOLUTE shared-memory address bits (unit ^ ((addr >> 7) & 3)), may a descriptor
//  (a) start at any 128-byte row (a tap = the same staged pixels shifted by whole pixels), and
//  (b) use an LBO that is not a multiple of 4096 B / 512 B (channel blocks of one halo row `lbo_rows` pixels apart)?
// A[m = b*32 + c][k] = buf[row_off + b*lbo_rows + k][c], 4 K-steps of 8 pixels, B = standard MN-major 64 x 32 tile.
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
    uint32_t addr = r * 128;   // a_s is 1024-aligned: (absolute address >> 7) & 3 == r & 3
    uint32_t off = addr + ((((c >> 1) ^ ((addr >> 7) & 3)) << 5) | ((c & 1) << 4));
    *(float4*)(a_s + off) = *(const float4*)(buf + r * 32 + c * 4);
  }
  for (int i = tid; i < 64 * 8; i += 128) {
    int r = i >> 3, c = i & 7;
    uint32_t off = r * 128 + ((((c >> 1) ^ (r & 3)) << 5) | ((c & 1) << 4));   // r = block*32 + pixel
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
    auto desc = [&](uint32_t addr, uint32_t lbo_bytes, uint32_t bo) {
      uint64_t d = 0;
      d |= (uint64_t)((addr & 0x3FFFF) >> 4);
      d |= (uint64_t)(lbo_bytes >> 4) << 16;   // next 32-channel block
      d |= (uint64_t)(512 >> 4) << 32;         // SBO: next group of 4 pixels
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)(bo & 7) << 49;
      d |= (uint64_t)1 << 61;                  // SWIZZLE_128B_BASE32B
      return d;
    };
    uint32_t bo = mode ? ((sa >> 7) & 3) : 0;
    const int lbo_rows = sbo_rows;             // (argument re-used: rows between the channel blocks of A)
    uint64_t ad = desc(sa, lbo_rows * 128, bo), bd = desc(sb, 4096, 0);
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((64 >> 3) << 17) | ((128 >> 4) << 24);
    for (int k = 0; k < 4; ++k) {              // next 8 pixels: +1024 B = +64 in the address field
      uint32_t acc = k != 0;
      asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(tmem),
                   "l"(ad + 64 * k), "l"(bd + 64 * k), "r"(idesc), "r"(acc)
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

extern "C" int probe_umma_mn(const float* buf, int buf_rows, const float* B, float* D, int row_off, int sbo_rows, int mode) {
  size_t smem = 62 * 1024 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem>>>(buf, buf_rows, B, D, row_off, sbo_rows, mode);
  return (int)cudaDeviceSynchronize();
}
