// Halo-tile tcgen05 implementation of the tap convolution for unit-input-stride stencils
// (forward 3x3 s1 [+ fused 1x1 s2 skip projection], data gradient of a 3x3 s1 conv, and the four
// parity classes of a stride-2 data gradient [+ skip gradient]).
//
// Why: the gather kernel (tapconv_umma.cu) re-reads and re-splits every input pixel once per tap
// (9x) and is bound by the latency of those loads (ncu: long_scoreboard, tensor pipe 14 %).  Here an
// output tile is 16 rows x 8 pixels of one image; its input HALO ((16+dy_span) x (8+dx_span) pixels) is
// loaded, split into tf32 hi/lo and stored in shared memory ONCE per 32-channel half, in the K-major
// SWIZZLE_128B layout with one 128-byte "slot" per halo pixel.  A tap is then nothing but a
// descriptor: start = slot(dy,dx), 8 consecutive slots = 8 consecutive output pixels, stride between
// the 16 row groups (SBO) = halo pitch * 128 B.  (The tensor core applies the 128B swizzle to absolute
// shared-memory address bits, so start addresses at any 128-byte slot and any SBO are legal as long as
// the data is stored with the same absolute-address swizzle -- verified on B200 by tools/probe.)
//
// Weights arrive pre-split (hi/lo) and pre-swizzled from b200np_pack_conv_weight, one 16 KB K-block
// (tap, channel half) per cp.async.bulk into a 3-deep ring.
//
// Persistent, warp-specialised CTA (one per SM), tiles round-robin:
//   warps 0-7   halo producers (gather + split + store), 2 stages of one channel half each; eight warps
//               because the producers are instruction-issue bound (ncu: 1 warp per scheduler stalled on
//               fixed-latency dependencies), not memory bound
//   warp  8     weight producer (one lane issues bulk copies); also owns the TMEM allocation
//   warp  9     MMA issuer (one lane)
//   warps 10-13 epilogue (TMEM -> registers -> bias / ReLU mask / activation -> NHWC global)
// Two accumulator sets in TMEM (2 x 256 columns in the fp32-grade mode) let the epilogue of tile i
// overlap the MMAs of tile i+1.
#include "tapconv.cuh"
#include "umma.cuh"

namespace b200np {

using namespace umma;

// Diagnostic hook (tools/halo_stalls.py): when set, every CTA writes the cycles its MMA lane, producers and
// epilogue spent blocked on each barrier.  nullptr in normal operation.
static long long* g_halo_dbg = nullptr;
static int g_halo_min_taps = 1;
extern "C" void b200np_debug_set_halo_min_taps(int n) { g_halo_min_taps = n; }
extern "C" void b200np_debug_set_halo_timing(long long* buf) { g_halo_dbg = buf; }

namespace {

constexpr int kTileRows = 16, kTileCols = 8;
constexpr int kMaxHaloSlots = 184;                       // 18 x 10 = 180, rounded up to a multiple of 8
constexpr uint32_t kHaloBytes = kMaxHaloSlots * 128;     // 23,552 B (multiple of 1024)
constexpr uint32_t kSkipBytes = 128 * 128;               // 16 KB
constexpr int kAStages = 2, kMaxBStages = 10;
constexpr int kRot = 2;   // rotating accumulator blocks per accumulator set
constexpr uint32_t kBSlotBytes = 2 * kBBytes;            // hi + lo = 16 KB
constexpr int kProducerWarps = 8, kProducerThreads = 32 * kProducerWarps;
constexpr int kWeightWarp = kProducerWarps, kMmaWarp = kProducerWarps + 1, kEpiWarp0 = kProducerWarps + 2;
constexpr int kThreads = 32 * (kProducerWarps + 2 + 4);
constexpr int kSlotsPerPass = kProducerThreads / 8;                       // slots covered by one pass of the producers
constexpr int kMaxTasks = (kMaxHaloSlots + 128 + kSlotsPerPass - 1) / kSlotsPerPass;  // chunk tasks per producer thread

struct HaloArgs {
  TapConvArgs t;
  const float* bp[2];   // pre-split, pre-swizzled weights per source: [half][slab][hi 8 KB | lo 8 KB]
  int nslabs[2];
  int dy_min, dx_min, HR, HC;   // halo geometry (src 0)
  int has_skip;                 // one extra tap on src 1 at offset (0,0)
  int tiles_x, tiles_total;
  // shared-memory plan (bytes from the 1024-aligned base), sized per launch: the weight ring takes
  // whatever the two halo stages leave -- its depth is what hides the L2 latency of the bulk copies
  uint32_t halo_bytes, stage_bytes, b_off, bar_off;
  int nb;                       // weight ring depth
  long long* dbg;               // optional [gridDim.x][8] stall-cycle counters
};

constexpr uint32_t kSmemBudget = 227 * 1024;
template <bool X3>
constexpr uint32_t b_slot_bytes() { return X3 ? kBSlotBytes : kBBytes; }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void mbar_wait_timed(uint64_t* bar, uint32_t parity, long long& acc, bool timed) {
  if (!timed) { mbar_wait(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}

struct Ring {
  int idx = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int n) {
    if (++idx == n) { idx = 0; phase ^= 1; }
  }
};

template <bool X3>
__global__ void __launch_bounds__(kThreads, 1) tapconv_halo_kernel(const HaloArgs h) {
  constexpr uint32_t kBSlot = b_slot_bytes<X3>();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + h.bar_off);
  uint64_t* a_full = bars;                  // [kAStages] producers -> MMA            (count 128)
  uint64_t* a_empty = bars + 2;             // [kAStages] MMA commit -> producers     (count 1)
  uint64_t* acc_full = bars + 4;            // [2] MMA commit -> epilogue             (count 1)
  uint64_t* acc_empty = bars + 6;           // [2] epilogue -> MMA                    (count 128)
  uint64_t* b_full = bars + 8;              // [nb] bulk copy tx -> MMA               (count 1 + tx)
  uint64_t* b_empty = b_full + kMaxBStages; // [nb] MMA commit -> weight warp         (count 1)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_empty + kMaxBStages);
  const int nb = h.nb;

  const TapConvArgs& a = h.t;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t kAccCols = X3 ? 256u : 128u;       // columns per accumulator set: kRot blocks of 128 | 64
  constexpr uint32_t kTmemCols = 2 * kAccCols;

  if (tid == 0) {
    for (int s = 0; s < kAStages; ++s) { mbar_init(a_full + s, kProducerThreads); mbar_init(a_empty + s, 1); }
    for (int s = 0; s < nb; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWeightWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int hslots = h.HR * h.HC;
  const int ntaps = a.ntaps;

  if (warp < kProducerWarps) {
    // ===================== halo producers =====================
    // Thread = fixed 16-byte chunk c of slots s0, s0+16, s0+32, ...  The (row, column) of those slots
    // advances by a constant (16 / HC, 16 % HC) step, so no division appears in the per-tile loop (with
    // one producer warp per scheduler, address arithmetic was the producers' critical path).
    // Slot space of a stage plane: [0, hslots) halo pixels, [hpad, hpad + 128) the skip tile (hpad = hslots
    // rounded up to 8, so the skip tile starts 1024-aligned and the plain K-major layout "row m at m*128,
    // chunk c at c ^ (m & 7)" coincides with the halo's absolute-address swizzle: ONE store formula).
    Ring st;
    long long dbg_a = 0;
    const long long dbg_t0 = clock64();
    const int c = tid & 7, s0 = tid >> 3;
    const int step_y = kSlotsPerPass / h.HC, step_x = kSlotsPerPass - step_y * h.HC;
    const int hy0 = s0 / h.HC, hx0 = s0 - hy0 * h.HC;
    const int hpad = (hslots + 7) & ~7;
    const int nslots = h.has_skip ? hpad + 128 : hslots;
    const long long row_pitch = (long long)a.srcW[0] * 64;
    const uint32_t plane = h.halo_bytes + (h.has_skip ? kSkipBytes : 0u);
    for (int tile = blockIdx.x; tile < h.tiles_total; tile += gridDim.x) {
      const int xt = tile % h.tiles_x, rb = tile / h.tiles_x;
      const int r0 = rb * kTileRows;
      const int n = r0 / a.OH, oy0 = r0 - n * a.OH, ox0 = xt * kTileCols;
      const int iy_base = oy0 + h.dy_min, ix_base = ox0 + h.dx_min;
      for (int half = 0; half < 2; ++half) {
        const float* base0 = a.src[0] + ((long long)n * a.srcH[0] * a.srcW[0]) * 64 + half * 32 + c * 4;
        const float* base1 = h.has_skip ? a.src[1] + ((long long)n * a.srcH[1] * a.srcW[1]) * 64 + half * 32 + c * 4
                                        : nullptr;
        float4 v[kMaxTasks];
        int slot = s0, hy = hy0, hx = hx0;
#pragma unroll
        for (int i = 0; i < kMaxTasks; ++i) {
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (slot < hslots) {
            const int iy = iy_base + hy, ix = ix_base + hx;
            if (iy >= 0 && iy < a.srcH[0] && ix >= 0 && ix < a.srcW[0]) v[i] = ldg4(base0 + iy * row_pitch + ix * 64);
          } else if (slot >= hpad && slot < nslots) {
            const int m = slot - hpad;                   // output pixel (row group m>>3, column m&7)
            const int iy = (oy0 + (m >> 3)) * a.in_s[1], ix = (ox0 + (m & 7)) * a.in_s[1];
            v[i] = ldg4(base1 + ((long long)iy * a.srcW[1] + ix) * 64);
          }
          slot += kSlotsPerPass;
          hy += step_y;
          hx += step_x;
          if (hx >= h.HC) { hx -= h.HC; ++hy; }
        }
        mbar_wait_timed(a_empty + st.idx, st.phase ^ 1, dbg_a, h.dbg != nullptr);
        uint8_t* plane_hi = smem + st.idx * h.stage_bytes;
        uint8_t* plane_lo = plane_hi + plane;
        slot = s0;
#pragma unroll
        for (int i = 0; i < kMaxTasks; ++i) {
          if (slot < nslots) split_store(plane_hi, plane_lo, (uint32_t)(slot * 128 + ((c ^ (slot & 7)) << 4)), v[i], X3);
          slot += kSlotsPerPass;
        }
        fence_proxy_async();
        mbar_arrive(a_full + st.idx);
        st.advance(kAStages);
      }
    }
    if (h.dbg && tid == 0) { h.dbg[blockIdx.x * 8 + 0] = dbg_a; h.dbg[blockIdx.x * 8 + 1] = clock64() - dbg_t0; }
  } else if (warp == kWeightWarp) {
    // ===================== weight producer =====================
    if (lane == 0) {
      Ring bs;
      long long dbg_b = 0;
      for (int tile = blockIdx.x; tile < h.tiles_total; tile += gridDim.x) {
        for (int half = 0; half < 2; ++half) {
          for (int t = 0; t < ntaps; ++t) {
            const Tap tp = a.taps[t];
            const float* src = h.bp[tp.src] + ((long long)half * h.nslabs[tp.src] + tp.slab) * (kBSlotBytes / 4);
            mbar_wait_timed(b_empty + bs.idx, bs.phase ^ 1, dbg_b, h.dbg != nullptr);
            mbar_expect_tx(b_full + bs.idx, kBSlot);
            bulk_g2s(smem + h.b_off + bs.idx * kBSlot, src, kBSlot, b_full + bs.idx);
            bs.advance(nb);
          }
        }
      }
      if (h.dbg) h.dbg[blockIdx.x * 8 + 2] = dbg_b;
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    // One lane issues every tcgen05.mma of the CTA, so this loop's instruction count IS the kernel's critical
    // path (measured: 21 k of 25 k cycles per tile were spent here, not waiting).  Everything that does not
    // change per tile is precomputed: per-tap descriptor words live in a small shared table, descriptors are
    // assembled from 32-bit halves, the k-step is an immediate add on the low word.
    uint32_t* tap_tab = tmem_slot + 4;                    // [ntaps][2]: {A offset >> 4 (bit 31: skip tile), desc high word}
    if (lane < ntaps) {
      const Tap tp = a.taps[lane];
      const uint32_t sbo = tp.src == 0 ? (uint32_t)h.HC * 128u : 1024u;
      const uint32_t off = tp.src == 0 ? (uint32_t)((tp.dy - h.dy_min) * h.HC + (tp.dx - h.dx_min)) * 8u : 0x80000000u;
      tap_tab[lane * 2 + 0] = off;
      tap_tab[lane * 2 + 1] = (sbo >> 4) | (1u << 14) | (2u << 29);   // SBO, descriptor version 1, SWIZZLE_128B
    }
    __syncwarp();
    if (lane == 0) {
      Ring st, bs;
      int acc_set = 0;
      uint32_t acc_phase = 0;
      long long w_acc = 0, w_a = 0, w_b = 0;
      const bool timed = h.dbg != nullptr;
      const long long t_start = clock64();
      const uint32_t plane16 = (h.halo_bytes + (h.has_skip ? kSkipBytes : 0u)) >> 4;
      const uint32_t skip16 = h.halo_bytes >> 4;
      const uint32_t lbo_bits = 1u << 16;                 // LBO field (ignored for swizzled K-major; 1 by convention)
      const uint32_t b_hi_word = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t b_base16 = (smem_u32(smem + h.b_off) & 0x3FFFFu) >> 4;
      // fp32-grade mode: the weight slot holds [B_hi (64 rows) ; B_lo (64 rows)] contiguously, i.e. one 128-row
      // K-major operand.  A_hi x [B_hi;B_lo] as ONE N=128 MMA yields hi*hi (columns 0-63) and hi*lo (64-127)
      // for 64 cycles instead of 2 x 54.5 (N=64 MMAs are shared-memory-bandwidth bound: tools/probe), then
      // A_lo x B_hi (N=64) adds the other cross term into columns 64-127.  Blocks of 128 columns rotate
      // (kRot = 2) to bound the truncating accumulator's bias; the epilogue sums all column groups.
      constexpr uint32_t kIdescN128 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      auto mma = [&](uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t acc, uint32_t idesc) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %2};\n\t"
            "mov.b64 db, {%3, %4};\n\t"
            "setp.ne.b32 p, %6, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
            ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi_word), "r"(idesc), "r"(acc)
            : "memory");
      };
      for (int tile = blockIdx.x; tile < h.tiles_total; tile += gridDim.x) {
        mbar_wait_timed(acc_empty + acc_set, acc_phase ^ 1, w_acc, timed);   // epilogue has drained this set
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc_set * kAccCols;
        uint32_t rot = 0;                                 // accumulator block rotation (kb % kRot without a division)
        int kb = 0;
        for (int half = 0; half < 2; ++half) {
          mbar_wait_timed(a_full + st.idx, st.phase, w_a, timed);
          tc_fence_after();
          const uint32_t stage16 = (smem_u32(smem + st.idx * h.stage_bytes) & 0x3FFFFu) >> 4;
          for (int t = 0; t < ntaps; ++t, ++kb) {
            const uint32_t off = tap_tab[t * 2], a_hi_word = tap_tab[t * 2 + 1];
            const uint32_t a16 = stage16 + ((off & 0x80000000u) ? skip16 : off);
            const uint32_t ah = a16 | lbo_bits, al = (a16 + plane16) | lbo_bits;
            const uint32_t b16 = b_base16 + bs.idx * (kBSlot >> 4);
            const uint32_t bh = b16 | lbo_bits;
            mbar_wait_timed(b_full + bs.idx, bs.phase, w_b, timed);
            tc_fence_after();
            const uint32_t d_blk = d0 + rot * (X3 ? 128u : 64u);
            const uint32_t acc_first = kb >= kRot;
            if (X3) {
#pragma unroll
              for (int k = 0; k < 4; ++k) mma(d_blk, ah + 2 * k, a_hi_word, bh + 2 * k, k == 0 ? acc_first : 1u, kIdescN128);
#pragma unroll
              for (int k = 0; k < 4; ++k) mma(d_blk + 64, al + 2 * k, a_hi_word, bh + 2 * k, 1u, kIdescTf32_128x64);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) mma(d_blk, ah + 2 * k, a_hi_word, bh + 2 * k, k == 0 ? acc_first : 1u, kIdescTf32_128x64);
            }
            umma_commit(b_empty + bs.idx);
            bs.advance(nb);
            if (++rot == kRot) rot = 0;
          }
          umma_commit(a_empty + st.idx);
          st.advance(kAStages);
        }
        umma_commit(acc_full + acc_set);
        if (++acc_set == 2) { acc_set = 0; acc_phase ^= 1; }
      }
      if (timed) {
        long long* o = h.dbg + blockIdx.x * 8;
        o[3] = w_acc; o[4] = w_a; o[5] = w_b; o[6] = clock64() - t_start;
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                              // TMEM lane quarter this warp may access (warp id % 4)
    const int m = q * 32 + lane;                         // accumulator row = output pixel of the tile
    int acc_set = 0;
    uint32_t acc_phase = 0;
    long long w_e = 0;
    const int kb_total = 2 * ntaps;
    const int blocks_used = kb_total < kRot ? kb_total : kRot;
    for (int tile = blockIdx.x; tile < h.tiles_total; tile += gridDim.x) {
      const int xt = tile % h.tiles_x, rb = tile / h.tiles_x;
      const int r0 = rb * kTileRows;
      const int n = r0 / a.OH, oy = r0 - n * a.OH + (m >> 3), ox = xt * kTileCols + (m & 7);
      const long long off =
          (((long long)n * a.dstH + (long long)oy * a.dst_s + a.dst_oy) * a.dstW + (long long)ox * a.dst_s + a.dst_ox) * 64;
      mbar_wait_timed(acc_full + acc_set, acc_phase, w_e, h.dbg != nullptr);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc_set * kAccCols + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float acc[32];
        {   // sum the rotating blocks (and, fp32-grade, their cross-term column group), smallest terms first
          uint32_t r[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = 0.f;
          if (X3) {
#pragma unroll
            for (int b = 0; b < kRot; ++b)
              if (b < blocks_used) {
                tmem_ld32(taddr + b * 128 + 64 + hf * 32, r);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]);
              }
          }
#pragma unroll
          for (int b = 0; b < kRot; ++b)
            if (b < blocks_used) {
              tmem_ld32(taddr + b * (X3 ? 128 : 64) + hf * 32, r);
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]);
            }
        }
        if (hf == 1) {                                   // all TMEM reads of this set are done
          tc_fence_before();
          mbar_arrive(acc_empty + acc_set);
        }
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) {
          const int c = hf * 32 + 4 * qq;
          float4 o = make_float4(acc[4 * qq], acc[4 * qq + 1], acc[4 * qq + 2], acc[4 * qq + 3]);
          if (a.bias) {
            const float4 b = ldg4(a.bias + c);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          }
          if (a.bias2) {
            const float4 b = ldg4(a.bias2 + c);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          }
          if (a.mask) {
            const float4 mk = ldg4(a.mask + off + c);
            o.x = mk.x > 0.f ? o.x : 0.f; o.y = mk.y > 0.f ? o.y : 0.f;
            o.z = mk.z > 0.f ? o.z : 0.f; o.w = mk.w > 0.f ? o.w : 0.f;
          }
          if (a.act == B200NP_ACT_RELU) {
            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
          }
          *reinterpret_cast<float4*>(a.dst + off + c) = o;
        }
      }
      if (++acc_set == 2) { acc_set = 0; acc_phase ^= 1; }
    }
    if (h.dbg && warp == kEpiWarp0 && lane == 0) h.dbg[blockIdx.x * 8 + 7] = w_e;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWeightWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

template <bool X3>
int launch_halo(HaloArgs& h, cudaStream_t st) {
  // plan shared memory: two halo stages (+ skip tiles when fused), then as deep a weight ring as fits
  const uint32_t planes = X3 ? 2u : 1u;
  h.halo_bytes = ((uint32_t)(h.HR * h.HC) * 128u + 1023u) & ~1023u;
  h.stage_bytes = planes * (h.halo_bytes + (h.has_skip ? kSkipBytes : 0u));
  h.b_off = kAStages * h.stage_bytes;
  const uint32_t tail = 1024 /*alignment slack*/ + 512 /*barriers + tap table*/;
  int nb = (int)((kSmemBudget - tail - h.b_off) / b_slot_bytes<X3>());
  if (nb > kMaxBStages) nb = kMaxBStages;
  if (nb < 2) return B200NP_E_UNSUPPORTED;
  h.nb = nb;
  h.bar_off = h.b_off + nb * b_slot_bytes<X3>();
  const size_t smem = h.bar_off + tail;
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(tapconv_halo_kernel<X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return B200NP_E_LAUNCH;
    configured = smem;
  }
  int grid = h.tiles_total < kNumSMs ? h.tiles_total : kNumSMs;
  tapconv_halo_kernel<X3><<<grid, kThreads, smem, st>>>(h);
  return launch_status();
}

}  // namespace

// Eligibility + geometry.  `bp0` / `bp1`: pre-split weights of source 0 / 1 (see b200np_pack_conv_weight).
int launch_tapconv_halo(const TapConvArgs& a, const float* bp0, int nslabs0, const float* bp1, int nslabs1,
                        int precision, cudaStream_t st) {
  if (a.Cin != 64 || a.Cout != 64 || a.ntaps < 1 || a.ntaps > kMaxTaps || !bp0) return B200NP_E_UNSUPPORTED;
  if (a.act != B200NP_ACT_NONE && a.act != B200NP_ACT_RELU) return B200NP_E_UNSUPPORTED;
  if (a.OH % kTileRows != 0 || a.OW % kTileCols != 0 || a.in_s[0] != 1) return B200NP_E_UNSUPPORTED;
  // Even the sparse parity classes of a stride-2 data gradient (1-2 taps) are faster here than in the gather
  // kernel since the producers went to 8 warps (measured 1.39 ms vs 1.96 ms for the 4 classes of layer1).
  if (a.ntaps < g_halo_min_taps) return B200NP_E_UNSUPPORTED;
  HaloArgs h{};
  h.t = a;
  h.dbg = g_halo_dbg;
  h.bp[0] = bp0; h.bp[1] = bp1; h.nslabs[0] = nslabs0; h.nslabs[1] = nslabs1;
  int dy_min = 127, dy_max = -127, dx_min = 127, dx_max = -127, n0 = 0, n1 = 0;
  for (int t = 0; t < a.ntaps; ++t) {
    const Tap& tp = a.taps[t];
    if (tp.src == 0) {
      ++n0;
      dy_min = tp.dy < dy_min ? tp.dy : dy_min; dy_max = tp.dy > dy_max ? tp.dy : dy_max;
      dx_min = tp.dx < dx_min ? tp.dx : dx_min; dx_max = tp.dx > dx_max ? tp.dx : dx_max;
    } else {
      ++n1;
      if (tp.dy != 0 || tp.dx != 0 || !bp1) return B200NP_E_UNSUPPORTED;
    }
  }
  if (n0 < 1 || n1 > 1) return B200NP_E_UNSUPPORTED;
  h.dy_min = dy_min; h.dx_min = dx_min;
  h.HR = kTileRows + dy_max - dy_min; h.HC = kTileCols + dx_max - dx_min;
  if (h.HR * h.HC > kMaxHaloSlots) return B200NP_E_UNSUPPORTED;
  h.has_skip = n1;
  h.tiles_x = a.OW / kTileCols;
  const long long tiles = (long long)a.N * a.OH / kTileRows * h.tiles_x;
  if (tiles <= 0) return B200NP_OK;
  if (tiles > 0x7fffffff) return B200NP_E_UNSUPPORTED;
  h.tiles_total = (int)tiles;
  return precision == B200NP_PREC_TF32 ? launch_halo<false>(h, st) : launch_halo<true>(h, st);
}

}  // namespace b200np
