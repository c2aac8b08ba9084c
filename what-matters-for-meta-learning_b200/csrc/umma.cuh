// tcgen05 / TMEM / mbarrier building blocks shared by the tensor-core kernels (sm_100a inline PTX).
#pragma once
#include "common.cuh"

namespace b200np {
namespace umma {

constexpr int kBM = 128, kBN = 64, kBK = 32;          // K-block = 32 channels = one 128 B swizzle row
constexpr int kStages = 2;
constexpr uint32_t kABytes = kBM * kBK * 4;           // 16 KB
constexpr uint32_t kBBytes = kBN * kBK * 4;           //  8 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (sticky error the host sees) instead of hanging the GPU.  The bound is WALL TIME
// (20 s on %globaltimer, sampled every 64 K failed probes), not a spin count: under compute-sanitizer / ncu replay
// or when a co-resident stream starves the CTA, a healthy wait can take millions of probes.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xffffu) == 0) {
      uint64_t t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > 20000000000ull) __trap();
    }
  }
}
// One lane of a converged warp (the lowest active one).  The MMA-issuing code runs warp-uniformly and only
// the tcgen05.mma / tcgen05.commit instructions sit under this predicate, so that descriptors stay in
// uniform registers (see tapconv_halo.cu).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// `issue`: instruction-level guard.  tcgen05.mma / tcgen05.commit read their operands from UNIFORM registers,
// and the compiler keeps a value there only if it is computed in warp-uniform control flow.  The issuing warp
// therefore runs its loop with all 32 lanes and passes issue = (this lane is the elected one); wrapping the
// whole loop in `if (tid == 0)` instead costs an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall per MMA.
__device__ __forceinline__ void umma_commit(uint64_t* bar, uint32_t issue = 1u) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar)), "r"(issue)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 inputs, fp32 accumulate, issued by one thread.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate, uint32_t issue = 1u) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start address >> 4 in [0,14), LBO >> 4 in [16,30) (unused for swizzled K-major; 1 by convention),
// SBO >> 4 in [32,46) = 1024 B between 8-row groups, version 1 in [46,48), layout type 2 in [61,64).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) at [4,6), a/b format
// TF32 (2) at [7,10)/[10,13), both K-major, N>>3 at [17,23), M>>4 at [24,29).
constexpr uint32_t kIdescTf32_128x64 = (1u << 4) | (2u << 7) | (2u << 10) | ((kBN >> 3) << 17) | ((kBM >> 4) << 24);

__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) {
  return static_cast<uint32_t>((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}
// fp32 -> tf32 with round-to-nearest, ties away from zero (the tensor core itself truncates: measured 2.3x the
// error of cuDNN's RN path).  The result has its low 13 mantissa bits clear, so hardware truncation is a no-op.
// Done on the bit pattern: adding half a tf32 ulp to the magnitude and clearing the low bits IS `cvt.rna.tf32.f32`
// (Inf stays Inf, NaN stays NaN), in 2 integer instructions -- ptxas expands the cvt into ~5 (FADD/FSETP/SEL/LOP3),
// which made the operand conversion the largest share of every staging loop (SASS: 530 instructions per thread and
// K-block in the weight-gradient kernel, half of them this).
__device__ __forceinline__ float to_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
// X3: hi = rn_tf32(x), lo = x - hi (exact in fp32; the tensor core truncates it to tf32); single pass: rn_tf32(x).
// Rounding lo as well (B200NP_LO_RN=1) halves the residual of the split -- 2^-23 |x| instead of 2^-22 |x| on average,
// random in sign either way because sign(lo) is -- at 2 more integer instructions per element in loops that are
// instruction-issue bound; both are below the fp32 rounding of the products' sum.
#ifndef B200NP_LO_RN
#define B200NP_LO_RN 0
#endif
__device__ __forceinline__ float lo_part(float x, float hi) { return B200NP_LO_RN ? to_tf32(x - hi) : x - hi; }
// The stores are explicit st.shared with 32-bit addresses: through the 1024-byte alignment arithmetic on the
// dynamic shared-memory base the compiler loses the address space and emits generic 64-bit ST.E (and splits some
// of them into scalar stores).
__device__ __forceinline__ void sts128(uint32_t saddr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// Split for operands that a TMA copy has already placed in shared memory as RAW fp32: the tensor core truncates its
// fp32 operands to tf32 (finding 1), so the raw plane IS a valid `hi` operand (hi = trunc_tf32(x)); only the
// remainder lo = x - trunc_tf32(x) has to be written (exact in fp32: it is the low 13 mantissa bits; the tensor core's
// truncation of it leaves a residual <= 3 * 2^-23 |x|, against 2^-23 |x| for the round-to-nearest split).
__device__ __forceinline__ float4 lo_of_truncated(float4 v) {
  float4 l;
  l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
  l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
  l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
  l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
  return l;
}
__device__ __forceinline__ float4 to_tf32_4(float4 v) {
  return make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
}
__device__ __forceinline__ void split_store(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, float4 v, bool x3) {
  float4 h;
  h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
  sts128(smem_u32(hi_base) + off, h);
  if (x3) {
    float4 l;
    l.x = lo_part(v.x, h.x); l.y = lo_part(v.y, h.y); l.z = lo_part(v.z, h.z); l.w = lo_part(v.w, h.w);
    sts128(smem_u32(lo_base) + off, l);
  }
}

// 32 lanes x 32 columns of fp32 from TMEM into registers (warp-collective)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Accumulator blocks in TMEM.  The tensor core adds into its fp32 accumulator with truncation, which
// biases a long K loop (measured 3.7e-6 relative at K = 640 -- 10x the CUDA-core fp32 result).  The
// fp32-grade mode therefore spreads the hi*hi products round-robin over kAccX3 independent 64-column
// accumulators (each sees 1/kAccX3 of the K-blocks), keeps the 2^-11-smaller cross terms in a block
// of their own, and sums the blocks with round-to-nearest adds in the epilogue.
constexpr int kAccX3 = 3;
template <bool X3> struct AccCfg {
  static constexpr int kHi = X3 ? kAccX3 : 1;
  static constexpr int kBlocks = X3 ? kAccX3 + 1 : 1;
  static constexpr uint32_t kCols = X3 ? 256 : 64;  // power of two >= 64 * kBlocks
};
// sum the accumulator blocks of one 32-column half into acc[0..32)
template <bool X3>
__device__ __forceinline__ void gather_acc(uint32_t taddr, int col0, int hi_used, float (&acc)[32]) {
  uint32_t r[32];
  if (X3) {
    tmem_ld32(taddr + AccCfg<X3>::kHi * 64 + col0, r);  // cross terms first (smallest magnitude)
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  }
#pragma unroll
  for (int b = 0; b < AccCfg<X3>::kHi; ++b) {
    if (b < hi_used) {
      tmem_ld32(taddr + b * 64 + col0, r);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]);
    }
  }
}
}  // namespace umma
}  // namespace b200np
