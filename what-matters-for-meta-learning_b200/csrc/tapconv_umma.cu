// tcgen05 / TMEM implementation of the tap convolution (tapconv.cuh) for the 64-channel layers.
//
// GEMM view per CTA: D[128 pixels x 64 couts] (fp32 accumulator in TMEM, 64 columns)
//                    += sum over K-blocks (tap, 32-channel half) of A[128 x 32] * B[64 x 32]^T
// Operands are K-major (channels contiguous) in the canonical SWIZZLE_128B shared-memory layout:
// a row is 128 B (32 tf32), 8 rows form a 1024 B swizzle atom, 16-byte chunk c of row r sits at
// chunk position c ^ (r & 7).  tcgen05.mma.kind::tf32 consumes K = 8 per instruction, so a K-block
// is 4 instructions per pass.
//
// Precision modes
//   TF32    one pass on operands rounded to nearest tf32 while they are staged.
//   TF32X3  fp32-grade: every operand is split into hi = rn_tf32(x) and lo = rn_tf32(x - hi) while it
//           is staged, and D += hi*hi + lo*hi + hi*lo (the dropped lo*lo term is < 2^-22 relative).
//
// The im2col gather (zero padding, stride, second source for the fused skip projection, parity
// classes of the stride-2 data gradient) is done by the 128 threads with 16-byte loads straight
// into the swizzled layout -- the operand split needs the values in registers anyway.  A 2-stage
// shared-memory ring decouples the gather from the asynchronous MMAs via mbarriers
// (tcgen05.commit); two CTAs per SM overlap one CTA's gather latency with the other's MMAs.
#include "tapconv.cuh"
#include "umma.cuh"

namespace b200np {

namespace {

using namespace umma;


constexpr int kConvThreads = 256;
constexpr uint32_t kIdescTf32_128x128 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// One 128-pixel tile per CTA, K-blocks = (tap, channel half).  256 threads: eight consecutive lanes hold the
// eight 16-byte chunks of one 128-byte row (pixel x 32 channels, or weight row), so every warp load covers four
// whole lines; the epilogue goes through a shared-memory transpose for the same reason (see gemm_umma.cu).
// 3xTF32 uses the stacked [W_hi ; W_lo] operand: an N = 128 and an N = 64 MMA per k-step, two rotating
// 128-column accumulator blocks.
template <bool X3>
__global__ void __launch_bounds__(kConvThreads, 2) tapconv_umma_kernel(const TapConvArgs a) {
  constexpr uint32_t kStageBytes = (X3 ? 2u : 1u) * (kABytes + kBBytes);
  constexpr uint32_t kCols = X3 ? 256u : 64u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);  // [kStages] stage free, [1] accum done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kStages + 1);
  long long* off_s = reinterpret_cast<long long*>(bars + kStages + 2);         // [128] output offset per tile pixel, -1: none

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  const uint32_t leader = elect_one_sync();
  const long long M = (long long)a.N * a.OH * a.OW;
  const long long m0 = (long long)blockIdx.x * kBM;

  if (tid == 0) {
    for (int s = 0; s <= kStages; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < kBM) {  // output offsets of the tile's pixels, for the transposed epilogue
    const long long p = m0 + tid;
    long long off = -1;
    if (p < M) {
      const int ox = (int)(p % a.OW);
      const long long q = p / a.OW;
      const int oy = (int)(q % a.OH);
      const long long n = q / a.OH;
      off = ((n * a.dstH + (long long)oy * a.dst_s + a.dst_oy) * a.dstW + (long long)ox * a.dst_s + a.dst_ox) * 64;
    }
    off_s[tid] = off;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  // gather roles: chunk g_c of rows g_r + 32 i (A: i = 0..3 pixels of the tile, B: i = 0..1 weight rows)
  const int g_c = tid & 7, g_r = tid >> 3;
  int pn[4], poy[4], pox[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long p = m0 + g_r + 32 * i;
    pn[i] = -1; poy[i] = 0; pox[i] = 0;
    if (p < M) {
      pox[i] = (int)(p % a.OW);
      const long long q = p / a.OW;
      poy[i] = (int)(q % a.OH);
      pn[i] = (int)(q / a.OH);
    }
  }
  uint32_t a_off[4], b_off[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) a_off[i] = sw128_offset(g_r + 32 * i, g_c);
#pragma unroll
  for (int j = 0; j < 2; ++j) b_off[j] = sw128_offset(g_r + 32 * j, g_c);

  const int KB = a.ntaps * 2;
  // two register buffers: the loads of K-block kb+2 are in flight while kb+1 is staged and kb runs
  float4 av0[4], bv0[2], av1[4], bv1[2];
  auto fetch = [&](int kb, float4 (&av)[4], float4 (&bv)[2]) {
    const int t = kb >> 1, c0 = (kb & 1) * kBK + g_c * 4;
    const Tap tp = a.taps[t];
    const int s = tp.src;
    const float* src = a.src[s];
    const int sH = a.srcH[s], sW = a.srcW[s], ss = a.in_s[s];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int iy = poy[i] * ss + tp.dy, ix = pox[i] * ss + tp.dx;
      const bool ok = pn[i] >= 0 && iy >= 0 && iy < sH && ix >= 0 && ix < sW;
      av[i] = ok ? ldg4(src + (((long long)pn[i] * sH + iy) * sW + ix) * 64 + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float* bp = a.w[s] + ((long long)tp.slab * 64 + g_r) * 64 + c0;
#pragma unroll
    for (int j = 0; j < 2; ++j) bv[j] = ldg4(bp + 32 * 64 * j);
  };

  if (KB > 0) fetch(0, av0, bv0);
  if (KB > 1) fetch(1, av1, bv1);
  auto step = [&](int kb, float4 (&av)[4], float4 (&bv)[2]) {
    const int s = kb % kStages;
    const int use = kb / kStages;
    if (use >= 1) mbar_wait(bars + s, (use - 1) & 1);  // MMAs that read this stage have retired
    uint8_t* st = smem + s * kStageBytes;
    uint8_t* a_hi = st;
    uint8_t* a_lo = st + kABytes;                       // only in X3
    uint8_t* b_hi = st + (X3 ? 2 : 1) * kABytes;
    uint8_t* b_lo = b_hi + kBBytes;                     // only in X3; [W_hi ; W_lo] = one 128-row K-major tile
#pragma unroll
    for (int i = 0; i < 4; ++i) split_store(a_hi, a_lo, a_off[i], av[i], X3);
#pragma unroll
    for (int j = 0; j < 2; ++j) split_store(b_hi, b_lo, b_off[j], bv[j], X3);
    fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (warp_u == 0) {  // all 32 lanes: descriptors stay in uniform registers, the elected lane issues
      tc_fence_after();
      const uint64_t ah = make_kmajor_sw128_desc(smem_u32(a_hi)), bh = make_kmajor_sw128_desc(smem_u32(b_hi));
      if (X3) {
        const uint64_t al = make_kmajor_sw128_desc(smem_u32(a_lo));
        const uint32_t d_blk = tmem_d + (uint32_t)(kb & 1) * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k)  // +32 B (8 tf32) along K inside the swizzle atom = +2 in the address field
          umma_tf32(d_blk, ah + 2 * k, bh + 2 * k, kIdescTf32_128x128, (kb >= 2) | (k != 0), leader);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_tf32(d_blk + 64, al + 2 * k, bh + 2 * k, kIdescTf32_128x64, 1u, leader);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_tf32(tmem_d, ah + 2 * k, bh + 2 * k, kIdescTf32_128x64, (kb | k) != 0, leader);
      }
      umma_commit(bars + s, leader);                       // stage reusable once these MMAs retire
      if (kb == KB - 1) umma_commit(bars + kStages, leader);  // accumulator complete
    }
    if (kb + 2 < KB) fetch(kb + 2, av, bv);
  };
  for (int kb = 0; kb < KB; kb += 2) {
    step(kb, av0, bv0);
    if (kb + 1 < KB) step(kb + 1, av1, bv1);
  }

  // ---- epilogue: TMEM -> registers -> shared-memory transpose -> bias / ReLU-mask / activation -> NHWC global.
  // warps w and w + 4 share TMEM lane quarter w % 4 and take 32 of the 64 output channels each ----
  if (KB > 0) {
    mbar_wait(bars + kStages, 0);   // all MMAs retired: the stages are free and serve as staging
    tc_fence_after();
  }
  const int q4 = warp & 3, half = warp >> 2;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  if (KB > 0) {
    const uint32_t taddr = tmem_d + (static_cast<uint32_t>(q4 * 32) << 16);
    uint32_t r[32];
    if (X3) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {  // cross-term columns first (smallest magnitude), then the main products
        const int blk = q & 1, cross = q < 2;
        if (blk < KB) {
          tmem_ld32(taddr + blk * 128 + cross * 64 + half * 32, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]);
        }
      }
    } else {
      tmem_ld32(taddr + half * 32, r);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]);
    }
  }
  uint8_t* stg = smem + warp * 4096;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
        make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
  __syncwarp();
  {
    const int cj = lane & 7, c = half * 32 + cj * 4;   // this lane's four output channels
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.bias) bsum = ldg4(a.bias + c);
    if (a.bias2) {
      const float4 b = ldg4(a.bias2 + c);
      bsum.x += b.x; bsum.y += b.y; bsum.z += b.z; bsum.w += b.w;
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rr = it * 4 + (lane >> 3);
      const long long off = off_s[q4 * 32 + rr];
      float4 o = *reinterpret_cast<const float4*>(stg + rr * 128 + ((cj ^ (rr & 7)) << 4));
      if (off < 0) continue;
      o.x += bsum.x; o.y += bsum.y; o.z += bsum.z; o.w += bsum.w;
      if (a.mask_bits) {
        const uint32_t nb4 = __ldg(a.mask_bits + (off >> 6) * 2 + half) >> (4 * cj);
        o.x = (nb4 & 1u) ? o.x : 0.f; o.y = (nb4 & 2u) ? o.y : 0.f;
        o.z = (nb4 & 4u) ? o.z : 0.f; o.w = (nb4 & 8u) ? o.w : 0.f;
      } else if (a.mask) {
        const float4 mk = ldg4(a.mask + off + c);
        o.x = mk.x > 0.f ? o.x : 0.f; o.y = mk.y > 0.f ? o.y : 0.f;
        o.z = mk.z > 0.f ? o.z : 0.f; o.w = mk.w > 0.f ? o.w : 0.f;
      }
      if (a.act == B200NP_ACT_RELU) {
        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
      }
      *reinterpret_cast<float4*>(a.dst + off + c) = o;
    }
    if (a.relu_bits) {  // gates of the outputs just written: the 8 lanes of a pixel's 32-channel half OR their nibbles
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rr = it * 4 + (lane >> 3);
        const long long off = off_s[q4 * 32 + rr];
        const float4 o = off >= 0 ? *reinterpret_cast<const float4*>(a.dst + off + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t wb = (uint32_t)((o.x > 0.f) | ((o.y > 0.f) << 1) | ((o.z > 0.f) << 2) | ((o.w > 0.f) << 3)) << (4 * cj);
        wb |= __shfl_xor_sync(0xffffffffu, wb, 1);
        wb |= __shfl_xor_sync(0xffffffffu, wb, 2);
        wb |= __shfl_xor_sync(0xffffffffu, wb, 4);
        if (cj == 0 && off >= 0) a.relu_bits[(off >> 6) * 2 + half] = wb;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kCols) : "memory");
  }
}

template <bool X3>
int launch(const TapConvArgs& a, cudaStream_t st) {
  constexpr uint32_t kStageBytes = (X3 ? 2u : 1u) * (kABytes + kBBytes);
  const size_t smem = kStages * kStageBytes + 1024 + 64 + 1024;  // + per-pixel output offsets
  static bool configured = false;  // idempotent attribute; benign if two threads race
  if (!configured) {
    if (cudaFuncSetAttribute(tapconv_umma_kernel<X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return B200NP_E_LAUNCH;
    configured = true;
  }
  const long long M = (long long)a.N * a.OH * a.OW;
  if (M <= 0) return B200NP_OK;
  tapconv_umma_kernel<X3><<<(unsigned)ceil_div(M, kBM), kConvThreads, smem, st>>>(a);
  return launch_status();
}

// ------------------------------------------------------------------------------------------------
// Weight gradient on tcgen05.  Per CTA: one pair of taps (t0, t0+1) over one chunk of pixels.
//   D[128 x 64] : row m = tap_local*64 + ci, column = co
//   D += sum over pixels p of  X_tap[gather(p)][ci] * dY[p][co]
// i.e. A = [X_t0^T ; X_t1^T] (M = 128) and B = dY^T (N = 64), contraction over PIXELS.  Both operands
// are contiguous along their M/N index (channels) and strided along K (pixels): "MN-major" in UMMA
// terms.  For 32-bit operands the only MN-major shared-memory layout the tensor core accepts is
// SWIZZLE_128B_BASE32B (cute: Layout_MN_SW128_32B_Atom): an atom is 4 pixels x 32 channels (4 rows of
// 128 B) and the swizzle permutes 32-BYTE units -- unit j of pixel-row r sits at unit position
// j ^ (r & 3).  Atoms of successive 4-pixel groups are SBO = 512 B apart, the 32-channel blocks
// LBO = 4096 B apart (one 32-pixel K-block).  One tcgen05.mma (K = 8) consumes two pixel groups; its
// descriptor start advances by 1024 B per k-step.
// ------------------------------------------------------------------------------------------------
constexpr int kWgK = 32;                             // pixels per K-block
constexpr int kWgThreads = 256;
constexpr uint32_t kWgABytes = 128 * kWgK * 4;        // 16 KB : 4 channel blocks x 4 pixel groups x 1 KB
constexpr uint32_t kWgBBytes = 64 * kWgK * 4;         //  8 KB : 2 channel blocks x 4 pixel groups x 1 KB
constexpr uint32_t kIdescTf32_128x64_MN = kIdescTf32_128x64 | (1u << 15) | (1u << 16);  // a_major = b_major = MN
constexpr uint32_t kIdescTf32_128x128_MN =
    (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(4096 >> 4) << 16;  // LBO: next 32-channel block
  d |= static_cast<uint64_t>(512 >> 4) << 32;   // SBO: next group of 4 pixels
  d |= static_cast<uint64_t>(1) << 46;          // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(1) << 61;          // layout type 1 = SWIZZLE_128B_BASE32B
  return d;
}
// byte offset of 16-byte chunk `chunk` (0..7) of pixel-row k (0..31) in channel block `cblock`
__device__ __forceinline__ uint32_t mn_offset(int cblock, int k, int chunk) {
  return static_cast<uint32_t>(cblock * 4096 + (k >> 2) * 512 + (k & 3) * 128 + (((chunk >> 1) ^ (k & 3)) << 5) +
                               ((chunk & 1) << 4));
}

template <bool X3>
__global__ void __launch_bounds__(kWgThreads, 2) tapwgrad_umma_kernel(const TapWgradArgs a) {
  constexpr uint32_t kStageBytes = (X3 ? 2u : 1u) * (kWgABytes + kWgBBytes);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kStages + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  const uint32_t leader = elect_one_sync();
  const int chunk = blockIdx.y, pair = blockIdx.x;
  const long long M = (long long)a.N * a.OH * a.OW;
  const long long p_begin = (long long)chunk * a.pix_per_chunk;
  long long p_end = p_begin + a.pix_per_chunk;
  if (p_end > M) p_end = M;

  if (tid == 0) {
    for (int s = 0; s <= kStages; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(AccCfg<X3>::kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  // Gather roles.  The 16 lanes of a half-warp walk the 16 consecutive 16-byte chunks of ONE pixel (256 B),
  // so a warp-wide load touches 4 full 128-byte lines.  (With a thread per 128-byte half-pixel every load
  // touched 32 different lines = 32 wavefronts of the L1 data pipe, and that pipe -- shared with the
  // shared-memory stores and the tensor core's operand reads -- was 92% busy: the kernel's bound.)
  // thread = (chunk g_c of pixels g_k + 16 i, i = 0..1); A: both taps of the pair, B: dY.  256 threads: the
  // gather is a chain of dependent address arithmetic, so it is the number of warps per scheduler (4 with
  // two CTAs per SM) that hides it.
  const int g_c = tid & 15, g_k = tid >> 4;
  // shared-memory offsets of this thread's chunks do not depend on the K-block
  uint32_t g_off[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) g_off[i] = mn_offset(g_c >> 3, g_k + 16 * i, g_c & 7);
  const int t0 = pair * 2, t1 = pair * 2 + 1;
  const bool t1_ok = t1 < a.ntaps;
  const Tap tpa = a.taps[t0], tpb = a.taps[t1_ok ? t1 : t0];
  const long long KB = p_begin < p_end ? ((p_end - p_begin + kWgK - 1) / kWgK) : 0;
  // fetches run over consecutive K-blocks: walk (n, oy, ox) of pixel g_k incrementally (no divisions).
  // Everything is 32-bit (the launcher rejects more than 2^31 pixels) and the per-tap constants are hoisted:
  // the gather is the instruction-issue bound of this kernel.
  int f_p = (int)p_begin + g_k, f_n, f_ox, f_oy;
  {
    f_ox = f_p % a.OW;
    const int q = f_p / a.OW;
    f_oy = q % a.OH;
    f_n = q / a.OH;
  }
  const int pe = (int)p_end;
  struct TapC { const float* src; int sH, sW, ss, dy, dx; };
  auto tapc = [&](const Tap& tp) {
    TapC c;
    c.src = (tp.src ? a.src2 : a.src) + g_c * 4;
    c.sH = tp.src ? a.src2H : a.srcH; c.sW = tp.src ? a.src2W : a.srcW; c.ss = tp.src ? a.in_s2 : a.in_s;
    c.dy = tp.dy; c.dx = tp.dx;
    return c;
  };
  const TapC ta = tapc(tpa), tb = tapc(tpb);
  const float* const dyp = a.dy + g_c * 4;
  float4 av0[4], bv0[2], av1[4], bv1[2];
  const bool want_db = a.part_db != nullptr && pair == 0;   // one CTA per chunk sums dy for the bias gradient
  float4 dbs = make_float4(0.f, 0.f, 0.f, 0.f);
  auto gather = [&](const TapC& t, bool ok, int oy, int ox, int n) -> float4 {
    const int iy = oy * t.ss + t.dy, ix = ox * t.ss + t.dx;
    if (ok && (unsigned)iy < (unsigned)t.sH && (unsigned)ix < (unsigned)t.sW)
      return ldg4(t.src + (long long)((n * t.sH + iy) * t.sW + ix) * 64);
    return make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto fetch = [&](long long, float4 (&av)[4], float4 (&bv)[2]) {
    int ox = f_ox, oy = f_oy, n = f_n;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int p = f_p + 16 * i;
      const bool pix_ok = p < pe;
      av[i] = gather(ta, pix_ok, oy, ox, n);
      av[2 + i] = gather(tb, pix_ok && t1_ok, oy, ox, n);
      bv[i] = pix_ok ? ldg4(dyp + (long long)p * 64) : make_float4(0.f, 0.f, 0.f, 0.f);
      ox += 16;
      while (ox >= a.OW) {
        ox -= a.OW;
        if (++oy == a.OH) { oy = 0; ++n; }
      }
    }
    f_p += kWgK;  // after two steps of 16 the walk stands at pixel g_k of the next K-block
    f_ox = ox; f_oy = oy; f_n = n;
  };

  if (KB > 0) fetch(0, av0, bv0);
  if (KB > 1) fetch(1, av1, bv1);
  auto step = [&](long long kb, float4 (&av)[4], float4 (&bv)[2]) {
    const int s = (int)(kb % kStages);
    const long long use = kb / kStages;
    if (use >= 1) mbar_wait(bars + s, (uint32_t)((use - 1) & 1));
    uint8_t* st = smem + s * kStageBytes;
    uint8_t* a_hi = st;
    uint8_t* a_lo = st + kWgABytes;
    uint8_t* b_hi = st + (X3 ? 2 : 1) * kWgABytes;
    uint8_t* b_lo = b_hi + kWgBBytes;
    // a quarter-warp holds the eight chunks of one 32-channel block of one pixel: all 32 banks, no conflicts
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      split_store(a_hi, a_lo, g_off[i], av[i], X3);
      split_store(a_hi, a_lo, g_off[i] + 2 * 4096, av[2 + i], X3);  // second tap: channel blocks 2, 3
      split_store(b_hi, b_lo, g_off[i], bv[i], X3);
      if (want_db) { dbs.x += bv[i].x; dbs.y += bv[i].y; dbs.z += bv[i].z; dbs.w += bv[i].w; }
    }
    fence_proxy_async();
    __syncthreads();
    if (warp_u == 0) {  // all 32 lanes: descriptors stay in uniform registers, the elected lane issues
      tc_fence_after();
      const uint64_t ah = make_mnmajor_sw128_desc(smem_u32(a_hi)), bh = make_mnmajor_sw128_desc(smem_u32(b_hi));
      if (X3) {
        // dY_hi and dY_lo are adjacent 32-channel blocks of one MN-major tile, so a single N = 128 MMA forms
        // X_hi * [dY_hi | dY_lo] (64 vs 2 x 54 cycles, and X_hi is read from shared memory once); X_lo * dY_hi
        // goes on top of the cross-term columns.  Two rotating 128-column blocks (see AccCfg) per CTA.
        const uint64_t al = make_mnmajor_sw128_desc(smem_u32(a_lo));
        const uint32_t d_blk = tmem_d + (uint32_t)(kb & 1) * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k)  // next group of 8 pixels: +1024 B = +64 in the address field
          umma_tf32(d_blk, ah + 64 * k, bh + 64 * k, kIdescTf32_128x128_MN, (kb >= 2) | (k != 0), leader);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_tf32(d_blk + 64, al + 64 * k, bh + 64 * k, kIdescTf32_128x64_MN, 1u, leader);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_tf32(tmem_d, ah + 64 * k, bh + 64 * k, kIdescTf32_128x64_MN, (kb | k) != 0, leader);
      }
      umma_commit(bars + s, leader);
      if (kb == KB - 1) umma_commit(bars + kStages, leader);
    }
    if (kb + 2 < KB) fetch(kb + 2, av, bv);
  };
  for (long long kb = 0; kb < KB; kb += 2) {
    step(kb, av0, bv0);
    if (kb + 1 < KB) step(kb + 1, av1, bv1);
  }

  // epilogue: warps w and w + 4 share TMEM lane quarter w % 4 (rows m = 32 (w % 4) + l = tap_local*64 + ci)
  // and take 32 of the 64 cout columns each
  const int q4 = warp & 3, half = warp >> 2;
  const int tl = q4 >> 1, ci = (q4 & 1) * 32 + lane;
  const int tap = pair * 2 + tl;
  if (KB > 0) {
    mbar_wait(bars + kStages, 0);
    tc_fence_after();
  }
  if (want_db) {  // CTA-uniform.  All MMAs have retired: the stages are free for the 16-row reduction
    float* red = reinterpret_cast<float*>(smem);
    *reinterpret_cast<float4*>(red + g_k * 64 + g_c * 4) = dbs;
    __syncthreads();
    if (tid < 64) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) t += red[r * 64 + tid];
      a.part_db[(long long)chunk * 64 + tid] = t;
    }
  }
  {
    const uint32_t taddr = tmem_d + (static_cast<uint32_t>(q4 * 32) << 16);
    float* po = a.part + ((long long)chunk * a.ntaps + (tap < a.ntaps ? tap : 0)) * 64 * 64 + ci;
    {
      float acc[32];
      if (KB > 0) {
        if (X3) {  // cross-term columns first (smallest magnitude), then the main products, both blocks
          uint32_t r[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int blk = q & 1, cross = q < 2;
            if (blk < KB) {
              tmem_ld32(taddr + blk * 128 + cross * 64 + half * 32, r);
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]);
            }
          }
        } else {
          gather_acc<false>(taddr, half * 32, 1, acc);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.f;
      }
      if (tap < a.ntaps) {
#pragma unroll
        for (int j = 0; j < 32; ++j) po[(half * 32 + j) * 64] = acc[j];  // lanes = consecutive ci: coalesced
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(AccCfg<X3>::kCols) : "memory");
  }
}

template <bool X3>
int launch_wgrad(const TapWgradArgs& a, cudaStream_t st) {
  constexpr uint32_t kStageBytes = (X3 ? 2u : 1u) * (kWgABytes + kWgBBytes);
  const size_t smem = kStages * kStageBytes + 1024 + 64;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(tapwgrad_umma_kernel<X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return B200NP_E_LAUNCH;
    configured = true;
  }
  dim3 grid((a.ntaps + 1) / 2, a.chunks);  // pair fastest: the CTAs that share a pixel chunk run together and share it in L2
  tapwgrad_umma_kernel<X3><<<grid, kWgThreads, smem, st>>>(a);
  return launch_status();
}

}  // namespace

int launch_tapconv_umma(const TapConvArgs& a, int precision, cudaStream_t st) {
  if (a.Cin != 64 || a.Cout != 64 || a.ntaps > kMaxTaps) return B200NP_E_UNSUPPORTED;
  if (a.act != B200NP_ACT_NONE && a.act != B200NP_ACT_RELU) return B200NP_E_UNSUPPORTED;
  return precision == B200NP_PREC_TF32 ? launch<false>(a, st) : launch<true>(a, st);
}

int launch_tapwgrad_umma(const TapWgradArgs& a, int precision, cudaStream_t st) {
  if (a.Cin != 64 || a.Cout != 64 || a.ntaps > kMaxTaps || a.ntaps < 1 || a.pix_per_chunk % kWgK != 0)
    return B200NP_E_UNSUPPORTED;
  if ((long long)a.N * a.OH * a.OW >= (1LL << 31) - 64 || (long long)a.N * a.srcH * a.srcW >= (1LL << 31) ||
      (long long)a.N * a.src2H * a.src2W >= (1LL << 31))
    return B200NP_E_UNSUPPORTED;  // the kernel indexes pixels with 32-bit integers
  return precision == B200NP_PREC_TF32 ? launch_wgrad<false>(a, st) : launch_wgrad<true>(a, st);
}

}  // namespace b200np
