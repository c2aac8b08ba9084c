// tcgen05 / TMEM implementation of the strided, grouped GEMM of gemm.cu (same descriptor, same epilogue)
// for the dense layers and the FAVOR+ feature projections.
//
//   C[128 x 64 tile] (+)= A[128 x 32] * B[32 x 64] per K-block, fp32 accumulators in TMEM.
//
// The three autograd roles of a dense layer differ only in which index of an operand is contiguous in
// memory, which maps 1:1 onto the tensor core's operand "major" modes:
//   forward   A k-contiguous (K-major)   B = W[n][k]  k-contiguous (K-major)
//   dgrad     A k-contiguous (K-major)   B = W[k][n]  n-contiguous (MN-major)
//   wgrad     A = dZ[k][m] m-contiguous (MN-major)   B = X[k][n] n-contiguous (MN-major)
// K-major tiles use SWIZZLE_128B (row = 32 k), MN-major tiles SWIZZLE_128B_BASE32B (row = one k, 32 m|n),
// see tapconv_umma.cu.  Operands are staged through registers (rn_tf32 rounding / hi-lo split as in the
// convolutions), two smem stages, loads of K-block kb+2 in flight while kb+1 is staged.
#include "common.cuh"
#include "umma.cuh"

namespace b200np {

using namespace umma;

namespace {

constexpr uint32_t kGA = 128 * 32 * 4, kGB = 64 * 32 * 4;  // 16 KB, 8 KB per plane

struct GemmUArgs {
  b200np_gemm_desc d;
  int a_vec, b_vec;  // 16-byte loads legal for A / B (pointer and leading-dimension alignment)
  int splits;        // split-K: blockIdx.z = group * splits + split; partial tiles are atomically added to C
};

// MN-major descriptor / offsets (same as the weight-gradient kernel): 32-wide blocks LBO = 4096 B apart,
// 4-row groups SBO = 512 B apart, 32-byte swizzle units
__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(4096 >> 4) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(1) << 61;
  return d;
}
__device__ __forceinline__ uint32_t mn_off(int block, int k, int chunk) {
  return static_cast<uint32_t>(block * 4096 + (k >> 2) * 512 + (k & 3) * 128 + (((chunk >> 1) ^ (k & 3)) << 5) +
                               ((chunk & 1) << 4));
}
// 4 consecutive floats from p[0..3] where only the first `valid` exist (valid <= 0: all zero)
__device__ __forceinline__ float4 load4_guarded(const float* p, int valid, bool vec) {
  if (valid >= 4 && vec) return ldg4(p);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid > 0) v.x = __ldg(p);
  if (valid > 1) v.y = __ldg(p + 1);
  if (valid > 2) v.z = __ldg(p + 2);
  if (valid > 3) v.w = __ldg(p + 3);
  return v;
}

constexpr int kGemmThreads = 256;
constexpr uint32_t kIdescN128 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// AK / BK: operand is k-contiguous (K-major) or m|n-contiguous (MN-major).
// 256 threads.  Every global access of a warp covers whole 128-byte lines: operand loads put consecutive lanes
// on consecutive 16-byte chunks of a row (a "thread = row" mapping makes each warp load touch 32 lines, i.e. 32
// wavefronts of the L1 data pipe), and the epilogue transposes the accumulator tile through shared memory so
// that bias / beta*C / row_scale*addend / activation and the store all run on coalesced float4 rows.
// 3xTF32: the B tile's hi and lo halves are adjacent, so one N = 128 MMA forms A_hi*[B_hi|B_lo] and an N = 64
// MMA adds A_lo*B_hi to the cross-term columns; two rotating 128-column accumulator blocks (see umma.cuh).
// CONV: 0 = plain operands; 1 / 2 = operand A / B is the virtual im2col matrix of an implicit convolution (own
// instantiations, so that the plain GEMMs carry none of its registers or branches)
template <bool X3, bool AK, bool BK, int CONV = 0>
__global__ void __launch_bounds__(kGemmThreads, 2) gemm_umma_kernel(const GemmUArgs g) {
  constexpr uint32_t kStageBytes = (X3 ? 2u : 1u) * (kGA + kGB);
  constexpr uint32_t kCols = X3 ? 256u : 64u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kStages + 1);
  const b200np_gemm_desc& d = g.d;
  // sum_groups: the groups are K-slices of one product (group pointers switch inside the K loop)
  const int grp = d.sum_groups ? 0 : blockIdx.z / g.splits;
  const int sp = d.sum_groups ? blockIdx.z : blockIdx.z - grp * g.splits;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  const uint32_t leader = elect_one_sync();
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * 64;
  const int M = d.M, N = d.N, K = d.K;

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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  // this CTA's K-block range (split-K: the small dense layers have 16 output tiles for 148 SMs)
  const int KB_grp = (K + 31) / 32;                       // K-blocks per group
  const int KB_all = d.sum_groups ? KB_grp * d.groups : KB_grp;
  const int kb_per = (KB_all + g.splits - 1) / g.splits;
  const int kb_lo = sp * kb_per;
  const int KB = KB_all - kb_lo < kb_per ? (KB_all - kb_lo > 0 ? KB_all - kb_lo : 0) : kb_per;

  // shared-memory offsets of this thread's chunks (K-block invariant)
  //   A K-major : rows (tid >> 3) + 32 i, chunk tid & 7        A MN-major: k rows (tid >> 5) + 8 i, m chunk tid & 31
  //   B K-major : rows (tid >> 3) + 32 j, chunk tid & 7        B MN-major: k rows (tid >> 4) + 16 j, n chunk tid & 15
  uint32_t a_off[4], b_off[2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    a_off[i] = AK ? sw128_offset((tid >> 3) + 32 * i, tid & 7) : mn_off((tid & 31) >> 3, (tid >> 5) + 8 * i, tid & 7);
#pragma unroll
  for (int j = 0; j < 2; ++j)
    b_off[j] = BK ? sw128_offset((tid >> 3) + 32 * j, tid & 7) : mn_off((tid & 15) >> 3, (tid >> 4) + 16 * j, tid & 7);

  // implicit convolution (d.conv_operand): per-thread geometry of the virtual im2col operand.  cv_y / cv_x: input
  // coordinates of tap (0, 0) of the thread's pixels (2*oy - 1, 2*ox - 1), cv_img: element offset of their image;
  // cv_r / cv_s / cv_ci: the tap and first channel of the thread's four consecutive k.
  int cv_y[4] = {0, 0, 0, 0}, cv_x[4] = {0, 0, 0, 0}, cv_r = 0, cv_s = 0, cv_ci = 0;
  long long cv_img[4] = {0, 0, 0, 0};
  const int cv_OW = d.conv_W >> 1, cv_OH = d.conv_H >> 1;
  if (CONV != 0) {
    const int rows = CONV == 1 ? 4 : 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i >= rows) break;
      // A: output pixel = GEMM row m (K-block invariant); B: output pixel = contraction index of the first K-block
      const long long pix = CONV == 1 ? (long long)m0 + (tid >> 3) + 32 * i
                                                : (long long)kb_lo * 32 + (tid >> 4) + 16 * i;
      const int ox = (int)(pix % cv_OW);
      const long long q = pix / cv_OW;
      cv_x[i] = 2 * ox - 1;
      cv_y[i] = 2 * (int)(q % cv_OH) - 1;
      cv_img[i] = (q / cv_OH) * d.conv_H * d.conv_W * d.conv_C;
      if (CONV == 1 && pix >= M) {   // a row of the last tile beyond the matrix: never in range (B rows are guarded by k < K)
        cv_y[i] = -(1 << 28);
        cv_img[i] = 0;
      }
    }
    if (CONV == 2) {        // B: this thread's four consecutive columns n = one tap, four channels
      const int nn = n0 + (tid & 15) * 4, tap = nn / d.conv_C;
      cv_ci = nn - tap * d.conv_C;
      cv_r = tap / 3;
      cv_s = tap - 3 * cv_r;
    }
  }
  float4 av0[4], bv0[2], av1[4], bv1[2];
  auto fetch = [&](int kb, float4 (&av)[4], float4 (&bv)[2]) {
    const int kbt = kb_lo + kb;
    const int gsel = d.sum_groups ? kbt / KB_grp : grp;
    const int k0 = (kbt - (d.sum_groups ? gsel * KB_grp : 0)) * 32;
    const float* __restrict__ A = d.A[gsel];
    const float* __restrict__ B = d.B[gsel];
    if (CONV == 1) {        // this thread's 4 consecutive k of the K-block: one tap, four channels
      const int kk = k0 + (tid & 7) * 4, tap = kk / d.conv_C;
      cv_ci = kk - tap * d.conv_C;
      cv_r = tap / 3;
      cv_s = tap - 3 * cv_r;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (AK) {
        const int m = m0 + (tid >> 3) + 32 * i, k = k0 + (tid & 7) * 4;
        if (CONV == 1) {   // virtual im2col row: the pixel geometry is K-block invariant, (tap, ci) row invariant
          const int iy = cv_y[i] + cv_r, ix = cv_x[i] + cv_s;
          av[i] = (k < K && iy >= 0 && iy < d.conv_H && ix >= 0 && ix < d.conv_W)
                      ? ldg4(A + cv_img[i] + ((long long)iy * d.conv_W + ix) * d.conv_C + cv_ci)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
        } else av[i] = load4_guarded(A + (long long)m * d.a_rs + k, m < M ? K - k : 0, g.a_vec);
      } else {
        const int k = k0 + (tid >> 5) + 8 * i, m = m0 + (tid & 31) * 4;
        av[i] = load4_guarded(A + (long long)k * d.a_cs + m, k < K ? M - m : 0, g.a_vec);
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (BK) {
        const int n = n0 + (tid >> 3) + 32 * j, k = k0 + (tid & 7) * 4;
        bv[j] = load4_guarded(B + (long long)n * d.b_cs + k, n < N ? K - k : 0, g.b_vec);
      } else {
        const int k = k0 + (tid >> 4) + 16 * j, n = n0 + (tid & 15) * 4;
        if (CONV == 2) {   // virtual im2col column block: (tap, ci) is thread invariant, the pixel walks with k
          const int iy = cv_y[j] + cv_r, ix = cv_x[j] + cv_s;
          bv[j] = (k < K && n < N && iy >= 0 && iy < d.conv_H && ix >= 0 && ix < d.conv_W)
                      ? ldg4(B + cv_img[j] + ((long long)iy * d.conv_W + ix) * d.conv_C + cv_ci)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
        } else bv[j] = load4_guarded(B + (long long)k * d.b_rs + n, k < K ? N - n : 0, g.b_vec);
      }
    }
    if (CONV == 2) {        // the next call fetches the next K-block: both pixels move on by 32
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        int ox = ((cv_x[j] + 1) >> 1) + 32, oy = (cv_y[j] + 1) >> 1;
        while (ox >= cv_OW) {
          ox -= cv_OW;
          if (++oy == cv_OH) { oy = 0; cv_img[j] += (long long)d.conv_H * d.conv_W * d.conv_C; }
        }
        cv_x[j] = 2 * ox - 1;
        cv_y[j] = 2 * oy - 1;
      }
    }
  };
  constexpr uint32_t kMajorBits = (AK ? 0u : (1u << 15)) | (BK ? 0u : (1u << 16));

  if (KB > 0) fetch(0, av0, bv0);
  if (KB > 1) fetch(1, av1, bv1);
  auto step = [&](int kb, float4 (&av)[4], float4 (&bv)[2]) {
    const int s = kb % kStages;
    const int use = kb / kStages;
    if (use >= 1) mbar_wait(bars + s, (use - 1) & 1);
    uint8_t* st = smem + s * kStageBytes;
    uint8_t* a_hi = st;
    uint8_t* a_lo = st + kGA;
    uint8_t* b_hi = st + (X3 ? 2 : 1) * kGA;
    uint8_t* b_lo = b_hi + kGB;   // adjacent to b_hi: [B_hi ; B_lo] is one 128-row (K-major) / 4-block (MN-major) tile
#pragma unroll
    for (int i = 0; i < 4; ++i) split_store(a_hi, a_lo, a_off[i], av[i], X3);
#pragma unroll
    for (int j = 0; j < 2; ++j) split_store(b_hi, b_lo, b_off[j], bv[j], X3);
    fence_proxy_async();
    __syncthreads();
    if (warp_u == 0) {  // all 32 lanes: descriptors stay in uniform registers, the elected lane issues
      tc_fence_after();
      const uint64_t ah = AK ? make_kmajor_sw128_desc(smem_u32(a_hi)) : mn_desc(smem_u32(a_hi));
      const uint64_t bh = BK ? make_kmajor_sw128_desc(smem_u32(b_hi)) : mn_desc(smem_u32(b_hi));
      constexpr uint64_t ka = AK ? 2 : 64, kbs = BK ? 2 : 64;  // descriptor advance per k-step of 8
      if (X3) {
        const uint64_t al = AK ? make_kmajor_sw128_desc(smem_u32(a_lo)) : mn_desc(smem_u32(a_lo));
        const uint32_t d_blk = tmem_d + (uint32_t)(kb & 1) * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_tf32(d_blk, ah + ka * k, bh + kbs * k, kIdescN128 | kMajorBits, (kb >= 2) | (k != 0), leader);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_tf32(d_blk + 64, al + ka * k, bh + kbs * k, kIdescTf32_128x64 | kMajorBits, 1u, leader);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_tf32(tmem_d, ah + ka * k, bh + kbs * k, kIdescTf32_128x64 | kMajorBits, (kb | k) != 0, leader);
      }
      umma_commit(bars + s, leader);
      if (kb == KB - 1) umma_commit(bars + kStages, leader);
    }
    if (kb + 2 < KB) fetch(kb + 2, av, bv);
  };
  for (int kb = 0; kb < KB; kb += 2) {
    step(kb, av0, bv0);
    if (kb + 1 < KB) step(kb + 1, av1, bv1);
  }

  // ---- epilogue.  warps w and w + 4 share TMEM lane quarter w % 4 and take 32 of the 64 columns each ----
  if (KB > 0) {
    mbar_wait(bars + kStages, 0);   // all MMAs retired: the stages are free and serve as transpose staging
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
  // stage this warp's 32 rows x 32 columns (row pitch 128 B, chunk index XOR row: conflict-free both ways)
  uint8_t* stg = smem + warp * 4096;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
        make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
  __syncwarp();
  {
    const int cj = lane & 7;                         // 16-byte column chunk of this lane
    const int n = n0 + half * 32 + cj * 4;
    const bool split = g.splits > 1;
    float* __restrict__ C = split ? static_cast<float*>(d.workspace) + (long long)sp * M * N : d.C[grp];
    const long long ldc = split ? N : d.ldc;
    const bool vec = n + 3 < N && (ldc & 3) == 0 && aligned16(C) &&
                     (split || ((!d.row_scale || ((d.ld_add & 3) == 0 && aligned16(d.addend)))));
    const float* __restrict__ bias = d.bias[grp];
    const bool use_rs = d.row_scale != nullptr && grp == 0 && !split;
    float bv4[4] = {0.f, 0.f, 0.f, 0.f};
    if (bias && !split) {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (n + e < N) bv4[e] = __ldg(bias + n + e);
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rr = it * 4 + (lane >> 3);
      const int m = m0 + q4 * 32 + rr;
      const float4 t = *reinterpret_cast<const float4*>(stg + rr * 128 + ((cj ^ (rr & 7)) << 4));
      if (m >= M || n >= N) continue;
      float v[4] = {t.x, t.y, t.z, t.w};
      float* cp = C + (long long)m * ldc + n;
      if (split) {  // raw partial sums; splitk_epilogue_kernel reduces the splits in a fixed order
        if (vec) *reinterpret_cast<float4*>(cp) = t;
        else
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (n + e < N) cp[e] = v[e];
        continue;
      }
      float old[4] = {0.f, 0.f, 0.f, 0.f}, add[4] = {0.f, 0.f, 0.f, 0.f};
      const float rs = use_rs ? __ldg(d.row_scale + m) : 0.f;
      if (vec) {
        if (d.beta != 0.f) { const float4 o = *reinterpret_cast<const float4*>(cp); old[0] = o.x; old[1] = o.y; old[2] = o.z; old[3] = o.w; }
        if (use_rs) { const float4 o = ldg4(d.addend + (long long)m * d.ld_add + n); add[0] = o.x; add[1] = o.y; add[2] = o.z; add[3] = o.w; }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (n + e < N) {
            if (d.beta != 0.f) old[e] = cp[e];
            if (use_rs) add[e] = __ldg(d.addend + (long long)m * d.ld_add + n + e);
          }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float x = d.alpha * v[e] + bv4[e];
        if (d.beta != 0.f) x = fmaf(d.beta, old[e], x);
        if (use_rs) x = fmaf(rs, add[e], x);
        if (d.act == B200NP_ACT_RELU) x = fmaxf(x, 0.f);
        else if (d.act == B200NP_ACT_TANH) x = tanhf(x);
        v[e] = x;
      }
      if (vec) *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
      else
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (n + e < N) cp[e] = v[e];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kCols) : "memory");
  }
}

template <bool X3, bool AK, bool BK, int CONV = 0>
int launch(const GemmUArgs& g, cudaStream_t st) {
  constexpr uint32_t kStageBytes = (X3 ? 2u : 1u) * (kGA + kGB);
  const size_t smem = kStages * kStageBytes + 1024 + 64;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_umma_kernel<X3, AK, BK, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return B200NP_E_LAUNCH;
    configured = true;
  }
  dim3 grid((g.d.M + 127) / 128, (g.d.N + 63) / 64, (g.d.sum_groups ? 1 : g.d.groups) * g.splits);
  gemm_umma_kernel<X3, AK, BK, CONV><<<grid, kGemmThreads, smem, st>>>(g);
  return launch_status();
}

// C = act(alpha * sum_s part[s] + bias + beta * C + row_scale * addend): the split-K partial tiles summed in
// split order (deterministic), then the same epilogue as the single-pass kernel
__global__ void splitk_epilogue_kernel(const b200np_gemm_desc d, int splits) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long MN = (long long)d.M * d.N;
  if (i >= MN) return;
  const int m = (int)(i / d.N), n = (int)(i - (long long)m * d.N);
  const float* part = static_cast<const float*>(d.workspace) + i;
  float acc = 0.f;
  for (int s = 0; s < splits; ++s) acc += __ldg(part + s * MN);
  float* cp = d.C[0] + (long long)m * d.ldc + n;
  float v = d.alpha * acc;
  if (d.bias[0]) v += __ldg(d.bias[0] + n);
  if (d.beta != 0.f) v = fmaf(d.beta, *cp, v);
  if (d.row_scale) v = fmaf(__ldg(d.row_scale + m), __ldg(d.addend + (long long)m * d.ld_add + n), v);
  if (d.act == B200NP_ACT_RELU) v = fmaxf(v, 0.f);
  else if (d.act == B200NP_ACT_TANH) v = tanhf(v);
  *cp = v;
}

}  // namespace

// Split-K factor: the dense layers of the path have a few hundred rows, i.e. ~16 output tiles for 148 SMs;
// their K range is spread over the idle SMs and the partial tiles are reduced from the caller's workspace.
int gemm_umma_splits(const b200np_gemm_desc& d) {
  if (d.precision == B200NP_PREC_FP32_SIMT || (d.groups != 1 && !d.sum_groups)) return 1;
  const int tiles = ((d.M + 127) / 128) * ((d.N + 63) / 64);
  const int KB = ((d.K + 31) / 32) * (d.sum_groups ? d.groups : 1);
  if (tiles * 2 > kNumSMs || KB < 8) return 1;
  int s = 2 * kNumSMs / tiles;   // two CTAs are resident per SM, and a CTA's K-blocks run one after the other
  if (s > KB / 4) s = KB / 4;
  return s > 1 ? s : 1;
}

// Returns B200NP_E_UNSUPPORTED when the shape is better served by the CUDA-core kernel.
int launch_gemm_umma(const b200np_gemm_desc& d, int a_vec, int b_vec, cudaStream_t st) {
  if (d.precision == B200NP_PREC_FP32_SIMT) return B200NP_E_UNSUPPORTED;
  if (d.K < 32 || d.N < 16 || (long long)d.M * d.N < 64 * 64) return B200NP_E_UNSUPPORTED;  // tiny: not worth a tile
  const bool AK = d.a_cs == 1, BK = d.b_rs == 1;
  if (d.conv_operand) {   // implicit convolution: only here (no CUDA-core form), so every refusal is an error
    if (d.groups != 1 || d.sum_groups || (d.conv_C & 3) || d.conv_C <= 0 || (d.conv_H & 1) || (d.conv_W & 1) ||
        d.conv_H < 2 || d.conv_W < 2)
      return B200NP_E_BADARG;
    if (d.conv_operand == 1 ? !AK : (d.conv_operand != 2 || BK)) return B200NP_E_BADARG;
    if (!aligned16(d.conv_operand == 1 ? d.A[0] : d.B[0])) return B200NP_E_BADARG;
  } else if (!AK && BK) return B200NP_E_UNSUPPORTED;
  if (!AK && BK) return B200NP_E_BADARG;
  if (d.sum_groups && d.K % 32 != 0) return B200NP_E_UNSUPPORTED;  // (A m-contiguous, B k-contiguous) never occurs on the path
  GemmUArgs g;
  g.d = d;
  g.a_vec = a_vec;
  g.b_vec = b_vec;
  g.splits = gemm_umma_splits(d);
  if ((size_t)g.splits * d.M * d.N * sizeof(float) > d.workspace_bytes || !d.workspace) g.splits = 1;
  const bool x3 = d.precision != B200NP_PREC_TF32;
  int rc;
  if (d.conv_operand == 1) {
    if (!BK) return B200NP_E_BADARG;   // forward form: weights k-contiguous
    rc = x3 ? launch<true, true, true, 1>(g, st) : launch<false, true, true, 1>(g, st);
  } else if (d.conv_operand == 2) {
    if (AK) return B200NP_E_BADARG;    // weight-gradient form: A = dY^T
    rc = x3 ? launch<true, false, false, 2>(g, st) : launch<false, false, false, 2>(g, st);
  } else if (AK && BK) rc = x3 ? launch<true, true, true>(g, st) : launch<false, true, true>(g, st);
  else if (AK && !BK) rc = x3 ? launch<true, true, false>(g, st) : launch<false, true, false>(g, st);
  else rc = x3 ? launch<true, false, false>(g, st) : launch<false, false, false>(g, st);
  if (rc != B200NP_OK || g.splits == 1) return rc;
  const long long n = (long long)d.M * d.N;
  splitk_epilogue_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d, g.splits);
  return launch_status();
}

}  // namespace b200np
