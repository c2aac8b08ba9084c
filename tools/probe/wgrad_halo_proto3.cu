// PROTOTYPE 3 (not product; written at the end of round 1, NOT YET RUN ON A GPU): csrc/tapwgrad_halo.cu -- both CTA roles,
// fused skip tap, fused bias gradient, the product's partial-tile layout -- with the MMA issue on a ninth warp and
// full / empty mbarriers instead of the block-wide barrier per tile (see wgrad_halo_proto2.cu for the reasoning).
// Entry: wgrad_halo_proto3(x, dy, xs|NULL, part [chunks][ntaps][64][64], part_db|NULL [chunks][64], N, H, W, x3, stream)
// returns the number of chunks (> 0) or a negative error.  Compare with b200np_conv_wgrad (run_wgrad_halo_proto3.py).
#include "tapconv.cuh"
#include "umma.cuh"

namespace b200np {

using namespace umma;

namespace {

constexpr int RT = 4, TW = 16, HW = TW + 2;
constexpr uint32_t S_A = HW * 128;                       // one channel block of one (halo) row: 2304 B
constexpr uint32_t A0_PLANE = (RT + 1) * 2 * S_A;        // role 0: hi (or lo) copy of the x halo, rows hr = 0..RT
constexpr uint32_t A1_PLANE = RT * 4 * S_A;              // role 1: [row][x cb0 | x cb1 | skip cb0 | skip cb1]
constexpr uint32_t S_B = TW * 128;                       // one channel block of one dY row: 2048 B
constexpr uint32_t B_TILE = RT * 4 * S_B;                // [row][hi | lo][block][pixel]
constexpr uint32_t STAGE = 2 * A1_PLANE + B_TILE;        // sized for role 1 (the larger): 106496 B = 104 * 1024
constexpr int kThreads = 256;       // producer threads; the CTA has one more warp, the MMA issuer
constexpr int XT0 = ((RT + 1) * HW * 16 + kThreads - 1) / kThreads;   // 16-byte tasks per thread: role 0 x halo (6)
constexpr int XT1 = (RT * (HW + TW) * 16 + kThreads - 1) / kThreads;  // role 1: x rows + skip rows (9)
constexpr int YT = RT * TW * 16 / kThreads;                           // dY tile (4)
constexpr uint32_t kIdesc128x64 = kIdescTf32_128x64 | (1u << 15) | (1u << 16);
constexpr uint32_t kIdesc128x128 =
    (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
static_assert(A0_PLANE % 512 == 0 && A1_PLANE % 512 == 0 && STAGE % 1024 == 0, "swizzle phase of the regions");

struct HaloWgradArgs {
  const float* x;      // [N, H, W, 64]
  const float* dy;     // [N, H, W, 64]
  const float* xs;     // nullable: skip source [N, 2H, 2W, 64], read at (2 oy, 2 ox)
  float* part;         // [chunks][ntaps][co][ci]
  float* part_db;      // nullable: [chunks][64]
  int N, H, W, ntaps;  // ntaps = 9 or 10 (with skip)
  long long tiles, per;
};

__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;   // next 32-channel block
  d |= static_cast<uint64_t>(512 >> 4) << 32;   // SBO: next group of 4 pixels
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(1) << 61;          // SWIZZLE_128B_BASE32B, base_offset 0
  return d;
}
// byte offset of 16-byte chunk `chunk` (0..7) inside the 128-byte row at byte offset `row` of a region whose base is a
// multiple of 512 B from a 1024-aligned stage
__device__ __forceinline__ uint32_t swz(uint32_t row, int chunk) {
  return row + ((((chunk >> 1) ^ ((row >> 7) & 3)) << 5) | ((chunk & 1) << 4));
}

template <bool X3, int ROLE>
__device__ __forceinline__ void run_role(const HaloWgradArgs& a, uint8_t* smem, uint64_t* bars, uint32_t tmem) {
  constexpr int XT = ROLE == 0 ? XT0 : XT1;
  constexpr uint32_t A_PLANE = ROLE == 0 ? A0_PLANE : A1_PLANE;
  constexpr uint32_t ROW_A = ROLE == 0 ? 2 * S_A : 4 * S_A;   // distance between the first halves of consecutive rows
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t leader = elect_one_sync();
  const int chunk = blockIdx.x;
  const int H = a.H, W = a.W;
  const int tiles_x = W / TW, tiles_img = tiles_x * (H / RT);
  const long long t_begin = chunk * a.per, t_end = t_begin + a.per < a.tiles ? t_begin + a.per : a.tiles;
  const int c16 = tid & 15;                      // this thread's 16-byte chunk of a pixel (the same for all its tasks)

  float4 xv[XT], yv[YT];
  float4 dbs = make_float4(0.f, 0.f, 0.f, 0.f);
  auto fetch = [&](long long tile) {
    const int n = (int)(tile / tiles_img), rem = (int)(tile - (long long)n * tiles_img);
    const int oy0 = (rem / tiles_x) * RT, ox0 = (rem % tiles_x) * TW;
#pragma unroll
    for (int i = 0; i < XT; ++i) {
      const int px = (tid + i * kThreads) >> 4;
      xv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ROLE == 0) {          // x halo: row hr <-> image row oy0 + hr - 1, column pc <-> ox0 + pc - 1
        const int hr = px / HW, pc = px - hr * HW;
        const int iy = oy0 + hr - 1, ix = ox0 + pc - 1;
        if (px < (RT + 1) * HW && (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W)
          xv[i] = ldg4(a.x + (((long long)n * H + iy) * W + ix) * 64 + c16 * 4);
      } else if (px < RT * HW) {   // x rows for dy = +1: row r <-> image row oy0 + r + 1
        const int r = px / HW, pc = px - r * HW;
        const int iy = oy0 + r + 1, ix = ox0 + pc - 1;
        if ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W)
          xv[i] = ldg4(a.x + (((long long)n * H + iy) * W + ix) * 64 + c16 * 4);
      } else if (px < RT * (HW + TW)) {   // skip rows: pixel k of row r <-> xs[2 (oy0 + r)][2 (ox0 + k)]
        const int q = px - RT * HW, r = q / TW, k = q - r * TW;
        if (a.xs) xv[i] = ldg4(a.xs + (((long long)n * 2 * H + 2 * (oy0 + r)) * 2 * W + 2 * (ox0 + k)) * 64 + c16 * 4);
      }
    }
#pragma unroll
    for (int i = 0; i < YT; ++i) {
      const int px = (tid + i * kThreads) >> 4;
      const int r = px / TW, pc = px - r * TW;
      yv[i] = ldg4(a.dy + (((long long)n * H + oy0 + r) * W + ox0 + pc) * 64 + c16 * 4);
    }
  };
  auto store = [&](uint8_t* st) {
    uint8_t* a_hi = st;
    uint8_t* a_lo = st + A_PLANE;
    uint8_t* b = st + 2 * A_PLANE;
    const int cb = c16 >> 3, ch = c16 & 7;
#pragma unroll
    for (int i = 0; i < XT; ++i) {
      const int px = (tid + i * kThreads) >> 4;
      if (ROLE == 0) {
        const int hr = px / HW, pc = px - hr * HW;
        if (px < (RT + 1) * HW) split_store(a_hi, a_lo, swz((uint32_t)(hr * 2 + cb) * S_A + pc * 128, ch), xv[i], X3);
      } else if (px < RT * HW) {
        const int r = px / HW, pc = px - r * HW;
        split_store(a_hi, a_lo, swz((uint32_t)(r * 4 + cb) * S_A + pc * 128, ch), xv[i], X3);
      } else if (px < RT * (HW + TW)) {   // skip pixel k sits at column k + 1: the dx = 0 pair (start column 1) reads it
        const int q = px - RT * HW, r = q / TW, k = q - r * TW;
        split_store(a_hi, a_lo, swz((uint32_t)(r * 4 + 2 + cb) * S_A + (k + 1) * 128, ch), xv[i], X3);
      }
    }
#pragma unroll
    for (int i = 0; i < YT; ++i) {
      const int px = (tid + i * kThreads) >> 4;
      const int r = px / TW, pc = px - r * TW;
      split_store(b, b + 2 * S_B, swz((uint32_t)(r * 4 + cb) * S_B + pc * 128, ch), yv[i], X3);   // lo copy: blocks 2, 3
      if (ROLE == 0) { dbs.x += yv[i].x; dbs.y += yv[i].y; dbs.z += yv[i].z; dbs.w += yv[i].w; }
    }
  };

  if (ROLE == 1) {   // (all 288 threads)
    // columns 0 and 17 of the skip blocks are never written by `store`; the unused dx = -1 / +1 second halves read
    // them, and although those accumulators are discarded they must not hold NaN/Inf garbage that an exception-free
    // tensor core would still multiply: zero both stages' operand regions once
    for (uint32_t o = tid * 16; o < 2 * STAGE; o += blockDim.x * 16) *reinterpret_cast<float4*>(smem + o) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
  }
  if (warp_u < 8) {
    // ---- producers: tid 0..255 ----
    fetch(t_begin);
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
    // ---- MMA issuer (warp 8): all 32 lanes run the loop, the elected lane issues ----
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
          for (int j = 0; j < 3; ++j) {   // dx = j - 1: the (halo) column of output pixel k is hf*8 + k + j
            const uint32_t aoff = r * ROW_A + (hf * 8 + j) * 128;
            const uint32_t d = tmem + j * 128;
            const uint32_t acc = (it == 0 && r == 0 && hf == 0) ? 0u : 1u;
            if (X3) {
              umma_tf32(d, mn_desc(a_hi + aoff, S_A), bd, kIdesc128x128, acc, leader);
              umma_tf32(d + 64, mn_desc(a_lo + aoff, S_A), bd, kIdesc128x64, 1u, leader);
            } else {
              umma_tf32(d, mn_desc(a_hi + aoff, S_A), bd, kIdesc128x64, acc, leader);
            }
          }
        }
      }
      umma_commit(bars + s, leader);
      if (tile + 1 == t_end) umma_commit(bars + 2, leader);
    }
    return;   // the issuer takes no part in the epilogue
  }
  // every chunk has at least one tile (the host sizes `per` that way)
  mbar_wait(bars + 2, 0);
  tc_fence_after();
  if (ROLE == 0 && a.part_db) {   // all MMAs have retired: the stages are free for the 16-row reduction of the dY sums
    float* red = reinterpret_cast<float*>(smem);
    *reinterpret_cast<float4*>(red + (tid >> 4) * 64 + c16 * 4) = dbs;
    asm volatile("bar.sync 1, 256;" ::: "memory");   // the 256 producer threads
    if (tid < 64) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) t += red[r * 64 + tid];
      a.part_db[(long long)chunk * 64 + tid] = t;
    }
  }
  // rows m = 32 q + lane = half_tap * 64 + ci; warps w and w + 4 share lane quarter q = w % 4 and take 32 co each
  const int q = warp & 3, half = warp >> 2, second = q >> 1, ci = (q & 1) * 32 + lane;
#pragma unroll 1
  for (int j = 0; j < 3; ++j) {
    int tap;
    if (ROLE == 0) tap = second * 3 + j;                                // (dy = -1 | 0, dx = j - 1)
    else tap = second ? ((j == 1 && a.ntaps > 9) ? 9 : -1) : 6 + j;     // (dy = +1, dx = j - 1) | the skip tap
    uint32_t r0[32], r1[32];
    const uint32_t ta = tmem + (static_cast<uint32_t>(q * 32) << 16) + j * 128 + half * 32;
    tmem_ld32(ta, r0);                 // warp-collective: every warp reads, only valid taps are written
    if (X3) tmem_ld32(ta + 64, r1);
    if (tap >= 0) {
      float* po = a.part + ((long long)chunk * a.ntaps + tap) * 64 * 64 + ci;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        po[(half * 32 + i) * 64] = (X3 ? __uint_as_float(r1[i]) : 0.f) + __uint_as_float(r0[i]);   // lanes = consecutive ci
    }
  }
}

template <bool X3>
__global__ void __launch_bounds__(kThreads + 32, 1) wgrad_halo_proto3_kernel(const HaloWgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * STAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bars + 0, 1);   // stage 0 drained by the tensor core
    mbar_init(bars + 1, 1);   // stage 1
    mbar_init(bars + 2, 1);   // all MMAs of the chunk retired
    mbar_init(bars + 3, kThreads);   // full[0]: every producer thread has stored stage 0
    mbar_init(bars + 4, kThreads);   // full[1]
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
  if (blockIdx.y == 0) run_role<X3, 0>(a, smem, bars, tmem);   // (the issuing warp returns from run_role early)
  else run_role<X3, 1>(a, smem, bars, tmem);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace
}  // namespace b200np

namespace b200np { std::atomic<long long> g_launches{0}; }

extern "C" int wgrad_halo_proto3(const float* x, const float* dy, const float* xs, float* part, float* part_db, int N, int H,
                                 int W, int x3, void* stream) {
  using namespace b200np;
  if (N <= 0 || H % RT || W % TW) return -1;
  HaloWgradArgs h{};
  h.x = x; h.dy = dy; h.xs = xs; h.part = part; h.part_db = part_db;
  h.N = N; h.H = H; h.W = W; h.ntaps = xs ? 10 : 9;
  h.tiles = (long long)N * (H / RT) * (W / TW);
  h.per = (h.tiles + kNumSMs - 1) / kNumSMs;
  const int chunks = (int)((h.tiles + h.per - 1) / h.per);
  const size_t smem = 2 * STAGE + 1024 + 1024;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (x3) {
    e = cudaFuncSetAttribute(wgrad_halo_proto3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -2;
    wgrad_halo_proto3_kernel<true><<<dim3(chunks, 2), kThreads + 32, smem, st>>>(h);
  } else {
    e = cudaFuncSetAttribute(wgrad_halo_proto3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -2;
    wgrad_halo_proto3_kernel<false><<<dim3(chunks, 2), kThreads + 32, smem, st>>>(h);
  }
  return cudaGetLastError() == cudaSuccess ? chunks : -3;
}
