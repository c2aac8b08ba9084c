// PROTOTYPE 2 (not product; written at the end of round 1, NOT YET RUN ON A GPU): wgrad_halo_proto.cu with the MMA issue
// moved to a warp of its own.  In prototype 1 warp 0 issues the 48 MMAs of a tile AND takes part in staging the next one,
// so a tile costs T_issue + T_store (5.5 K cycles measured against 2.8 K cycles of tensor work); here eight producer
// warps and one issuing warp hand stages over through full / empty mbarriers, and a tile should cost max(T_issue, T_store).
// Original header:
// PROTOTYPE (not product, not linked into libb200np.so): weight gradient of a 3x3 stride-1 64->64 convolution as a HALO
// kernel -- the round-2 plan of DESIGN.md section 9.  The product kernel (csrc/tapconv_umma.cu: tapwgrad_umma_kernel)
// gathers x once per TAP and dY once per tap PAIR; here a CTA stages the x halo and the dY tile of a 4 x 16 pixel
// tile ONCE and forms six taps from it:
//   * planes [halo row][channel block][pixel][32 ch] (128 B per pixel and block, hi and lo copies), stored with the
//     32-byte-unit swizzle of SWIZZLE_128B_BASE32B taken from absolute shared-memory address bits;
//   * a tap is a descriptor whose start is shifted by whole pixels (profiles/umma_mn_descriptor_probe_r1.txt);
//   * the vertical neighbours (dy, dx), (dy + 1, dx) are the two halves of M = 128: channel-block stride (LBO) = one
//     channel block of one halo row, and the next halo row follows at 2 * LBO;
//   * dY is [row][hi | lo][channel block][pixel]: one N = 128 MMA gives x_hi*dY_hi and x_hi*dY_lo, an N = 64 MMA adds
//     x_lo*dY_hi (the product kernel's 3xTF32 scheme).
// Role 0 computes the taps dy in {-1, 0}, role 1 the taps dy in {0, +1} (dy = 0 twice: the prototype keeps one code path);
// three pairs (dx = -1, 0, +1) x 128 TMEM columns per CTA, accumulated over all tiles of the CTA's chunk.
// Output: part[role][chunk][pair][m = 128][co = 64], m = (dy - dy0) * 64 + ci.
#include "umma.cuh"

using namespace b200np;
using namespace b200np::umma;

namespace {

constexpr int RT = 4, TW = 16, HW = TW + 2;
constexpr uint32_t S_A = HW * 128;                  // one channel block of one halo row (2304 B)
constexpr uint32_t A_PLANE = (RT + 1) * 2 * S_A;    // hi (or lo) copy of the x halo (23040 B = 45 * 512)
constexpr uint32_t S_B = TW * 128;                  // one channel block of one dY row (2048 B)
constexpr uint32_t B_TILE = RT * 4 * S_B;           // [row][hi | lo][block][pixel] (32768 B)
constexpr uint32_t STAGE = 2 * A_PLANE + B_TILE;    // 78848 B = 77 * 1024
constexpr int XT = ((RT + 1) * HW * 16 + 255) / 256;   // 16-byte tasks per thread: x halo (6)
constexpr int YT = RT * TW * 16 / 256;                 // dY tile (4)
constexpr uint32_t kIdesc128x64 = kIdescTf32_128x64 | (1u << 15) | (1u << 16);
constexpr uint32_t kIdesc128x128 =
    (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;   // next 32-channel block
  d |= static_cast<uint64_t>(512 >> 4) << 32;   // SBO: next group of 4 pixels
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(1) << 61;          // SWIZZLE_128B_BASE32B, base_offset 0
  return d;
}
// byte offset of 16-byte chunk `chunk` (0..7) inside the 128-byte row at byte offset `row` of a 1024-aligned region
__device__ __forceinline__ uint32_t swz(uint32_t row, int chunk) {
  return row + ((((chunk >> 1) ^ ((row >> 7) & 3)) << 5) | ((chunk & 1) << 4));
}

__global__ void __launch_bounds__(288, 1) wgrad_halo_proto2_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                  float* __restrict__ part, int N, int H, int W, int X3) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * STAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t leader = elect_one_sync();
  const int role = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  const int dy0 = role - 1;
  const int tiles_x = W / TW, tiles_y = H / RT, tiles_img = tiles_x * tiles_y;
  const long long tiles = (long long)N * tiles_img;
  const long long per = (tiles + chunks - 1) / chunks;
  const long long t_begin = chunk * per, t_end = t_begin + per < tiles ? t_begin + per : tiles;

  if (tid == 0) {
    mbar_init(bars + 0, 1);     // empty[0]: the tensor core has read stage 0 (commit)
    mbar_init(bars + 1, 1);     // empty[1]
    mbar_init(bars + 2, 1);     // all MMAs of the chunk retired
    mbar_init(bars + 3, 256);   // full[0]: the 256 producer threads have stored stage 0
    mbar_init(bars + 4, 256);   // full[1]
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  float4 xv[XT], yv[YT];
  auto fetch = [&](long long tile) {
    const int n = (int)(tile / tiles_img), rem = (int)(tile - (long long)n * tiles_img);
    const int oy0 = (rem / tiles_x) * RT, ox0 = (rem % tiles_x) * TW;
#pragma unroll
    for (int i = 0; i < XT; ++i) {
      const int id = tid + i * 256, px = id >> 4, c16 = id & 15;
      const int hr = px / HW, pc = px - hr * HW;
      const int iy = oy0 + hr + dy0, ix = ox0 + pc - 1;
      xv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (px < (RT + 1) * HW && (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W)
        xv[i] = ldg4(x + (((long long)n * H + iy) * W + ix) * 64 + c16 * 4);
    }
#pragma unroll
    for (int i = 0; i < YT; ++i) {
      const int id = tid + i * 256, px = id >> 4, c16 = id & 15;
      const int r = px / TW, pc = px - r * TW;
      yv[i] = ldg4(dy + (((long long)n * H + oy0 + r) * W + ox0 + pc) * 64 + c16 * 4);
    }
  };
  auto store = [&](uint8_t* st) {
    uint8_t* a_hi = st;
    uint8_t* a_lo = st + A_PLANE;
    uint8_t* b = st + 2 * A_PLANE;
#pragma unroll
    for (int i = 0; i < XT; ++i) {
      const int id = tid + i * 256, px = id >> 4, c16 = id & 15;
      const int hr = px / HW, pc = px - hr * HW;
      if (px < (RT + 1) * HW)
        split_store(a_hi, a_lo, swz((uint32_t)(hr * 2 + (c16 >> 3)) * S_A + pc * 128, c16 & 7), xv[i], X3 != 0);
    }
#pragma unroll
    for (int i = 0; i < YT; ++i) {
      const int id = tid + i * 256, px = id >> 4, c16 = id & 15;
      const int r = px / TW, pc = px - r * TW;
      // hi copy in blocks 0/1 of the row, lo copy in blocks 2/3
      split_store(b, b + 2 * S_B, swz((uint32_t)(r * 4 + (c16 >> 3)) * S_B + pc * 128, c16 & 7), yv[i], X3 != 0);
    }
  };

  if (warp_u < 8) {
    // ---- producers: tid 0..255 ----
    if (t_begin < t_end) fetch(t_begin);
    long long it = 0;
    for (long long tile = t_begin; tile < t_end; ++tile, ++it) {
      const int s = (int)(it & 1);
      if (it >= 2) mbar_wait(bars + s, (uint32_t)(((it >> 1) - 1) & 1));   // the MMAs of tile it-2 have read this stage
      store(smem + s * STAGE);
      fence_proxy_async();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bars + 3 + s)) : "memory");
      if (tile + 1 < t_end) fetch(tile + 1);
    }
  } else {
    // ---- MMA issuer: all 32 lanes run the loop (descriptors stay in uniform registers), the elected lane issues ----
    long long it = 0;
    for (long long tile = t_begin; tile < t_end; ++tile, ++it) {
      const int s = (int)(it & 1);
      mbar_wait(bars + 3 + s, (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const uint32_t a_hi = smem_u32(smem + s * STAGE), a_lo = a_hi + A_PLANE, b0 = a_hi + 2 * A_PLANE;
#pragma unroll 1
      for (int r = 0; r < RT; ++r) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const uint64_t bd = mn_desc(b0 + r * 4 * S_B + hf * 8 * 128, S_B);
#pragma unroll
          for (int j = 0; j < 3; ++j) {   // dx = j - 1: halo column of output pixel k is hf*8 + k + j
            const uint32_t aoff = r * 2 * S_A + (hf * 8 + j) * 128;
            const uint32_t d = tmem + j * 128;
            const uint32_t first = (it == 0 && r == 0 && hf == 0) ? 0u : 1u;
            if (X3) {
              umma_tf32(d, mn_desc(a_hi + aoff, S_A), bd, kIdesc128x128, first, leader);
              umma_tf32(d + 64, mn_desc(a_lo + aoff, S_A), bd, kIdesc128x64, 1u, leader);
            } else {
              umma_tf32(d, mn_desc(a_hi + aoff, S_A), bd, kIdesc128x64, first, leader);
            }
          }
        }
      }
      umma_commit(bars + s, leader);
      if (tile + 1 == t_end) umma_commit(bars + 2, leader);
    }
  }
  if (t_begin < t_end && warp_u < 8) {
    mbar_wait(bars + 2, 0);
    tc_fence_after();
    const int q = warp & 3, half = warp >> 2;
    float* out = part + ((long long)(role * chunks + chunk) * 3) * 128 * 64;
    for (int j = 0; j < 3; ++j) {
      uint32_t r0[32], r1[32];
      const uint32_t ta = tmem + (static_cast<uint32_t>(q * 32) << 16) + j * 128 + half * 32;
      tmem_ld32(ta, r0);
      if (X3) tmem_ld32(ta + 64, r1);
      float* o = out + ((long long)j * 128 + q * 32 + lane) * 64 + half * 32;
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(r0[i]) + (X3 ? __uint_as_float(r1[i]) : 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace

namespace b200np { std::atomic<long long> g_launches{0}; }

// part: [2][chunks][3][128][64] floats (zero it first: CTAs without tiles write nothing).  H % 4 == 0, W % 16 == 0.
extern "C" int wgrad_halo_proto2(const float* x, const float* dy, float* part, int N, int H, int W, int chunks, int x3,
                                void* stream) {
  if (H % RT || W % TW || chunks < 1) return -1;
  const size_t smem = 2 * STAGE + 1024 + 1024;
  if (cudaFuncSetAttribute(wgrad_halo_proto2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return -2;
  wgrad_halo_proto2_kernel<<<dim3(chunks, 2), 288, smem, reinterpret_cast<cudaStream_t>(stream)>>>(x, dy, part, N, H, W, x3);
  return (int)cudaGetLastError();
}
