// CUDA-core fp32 implementation of the tap convolution (see tapconv.cuh) and its weight gradient.
// Exact-fp32 arithmetic: it is the validation twin of the tcgen05 kernels and serves the layers
// whose channel counts do not fill a tensor-core tile (encoder_w0: 32/48 channels).
#include "tapconv.cuh"

namespace b200np {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, APAD = 4;

// 128 threads; thread (tm = tid/8, tn = tid%8) owns an 8x8 register tile: pixels tm*8.., couts tn*8..
__global__ void __launch_bounds__(128) tapconv_simt_kernel(const TapConvArgs a) {
  __shared__ __align__(16) float As[BK][BM + APAD];  // [k][pixel]
  __shared__ __align__(16) float Bs[BK][BN];         // [k][cout]
  const int tid = threadIdx.x;
  const long long M = (long long)a.N * a.OH * a.OW;
  const long long m0 = (long long)blockIdx.x * BM;

  // the pixel whose A row this thread gathers
  const long long p = m0 + tid;
  const bool pvalid = p < M;
  int ox = 0, oy = 0, n = 0;
  if (pvalid) {
    ox = (int)(p % a.OW);
    long long q = p / a.OW;
    oy = (int)(q % a.OH);
    n = (int)(q / a.OH);
  }
  // B rows: cout = tid/2, k-half = tid%2 (8 floats)
  const int bco = tid >> 1, bhalf = tid & 1;
  const bool bvalid = bco < a.Cout;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int chunks = a.Cin / BK;
  const int iters = a.ntaps * chunks;
  float4 av[4], bv[2];

  auto fetch = [&](int it) {
    const int t = it / chunks, c0 = (it - t * chunks) * BK;
    const Tap tp = a.taps[t];
    const int s = tp.src;
    const int iy = oy * a.in_s[s] + tp.dy, ix = ox * a.in_s[s] + tp.dx;
    const bool ok = pvalid && iy >= 0 && iy < a.srcH[s] && ix >= 0 && ix < a.srcW[s];
    if (ok) {
      const float* ap = a.src[s] + (((long long)n * a.srcH[s] + iy) * a.srcW[s] + ix) * a.Cin + c0;
#pragma unroll
      for (int q = 0; q < 4; ++q) av[q] = ldg4(ap + 4 * q);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) av[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (bvalid) {
      const float* bp = a.w[s] + ((long long)tp.slab * a.Cout + bco) * a.Cin + c0 + bhalf * 8;
      bv[0] = ldg4(bp);
      bv[1] = ldg4(bp + 4);
    } else {
      bv[0] = bv[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };

  const int tm = tid >> 3, tn = tid & 7;
  if (iters > 0) fetch(0);
  for (int it = 0; it < iters; ++it) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      As[q * 4 + 0][tid] = av[q].x;
      As[q * 4 + 1][tid] = av[q].y;
      As[q * 4 + 2][tid] = av[q].z;
      As[q * 4 + 3][tid] = av[q].w;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      Bs[bhalf * 8 + q * 4 + 0][bco] = bv[q].x;
      Bs[bhalf * 8 + q * 4 + 1][bco] = bv[q].y;
      Bs[bhalf * 8 + q * 4 + 2][bco] = bv[q].z;
      Bs[bhalf * 8 + q * 4 + 3][bco] = bv[q].w;
    }
    __syncthreads();
    if (it + 1 < iters) fetch(it + 1);  // global loads in flight while the FMAs run
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][tm * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][tm * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tn * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][tn * 8 + 4]);
      const float ar[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float br[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }

  // epilogue
  const int co0 = tn * 8;
  if (co0 >= a.Cout) return;
  float bsum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float b = a.bias ? __ldg(a.bias + co0 + j) : 0.f;
    if (a.bias2) b += __ldg(a.bias2 + co0 + j);
    bsum[j] = b;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long pp = m0 + tm * 8 + i;
    if (pp >= M) break;
    const int x_ = (int)(pp % a.OW);
    const long long q = pp / a.OW;
    const int y_ = (int)(q % a.OH);
    const long long n_ = q / a.OH;
    const long long off =
        ((n_ * a.dstH + (long long)y_ * a.dst_s + a.dst_oy) * a.dstW + (long long)x_ * a.dst_s + a.dst_ox) * a.Cout + co0;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = acc[i][j] + bsum[j];
    if (a.mask_bits) {  // Cout == 64 (checked by the caller): two 32-bit words per pixel
      const uint32_t byte = __ldg(a.mask_bits + (off >> 6) * 2 + (co0 >> 5)) >> (co0 & 31);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = ((byte >> j) & 1u) ? o[j] : 0.f;
    } else if (a.mask) {
      const float4 m0v = ldg4(a.mask + off), m1v = ldg4(a.mask + off + 4);
      const float mk[8] = {m0v.x, m0v.y, m0v.z, m0v.w, m1v.x, m1v.y, m1v.z, m1v.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = mk[j] > 0.f ? o[j] : 0.f;
    }
    if (a.act == B200NP_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
    }
    *reinterpret_cast<float4*>(a.dst + off) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(a.dst + off + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
}

// Weight gradient: grid (chunk, tap).  part[chunk][tap][co][ci] = sum over the chunk's pixels of
// dy[pix,co] * src[gather(pix,tap),ci].  128 threads, thread (tco = tid/16, tci = tid%16) owns
// 8 couts x 4 cins.
__global__ void __launch_bounds__(128) tapwgrad_simt_kernel(const TapWgradArgs a) {
  __shared__ __align__(16) float As[BK][64];  // [pixel][cout]
  __shared__ __align__(16) float Bs[BK][64];  // [pixel][cin]
  const int tid = threadIdx.x;
  const int chunk = blockIdx.x, t = blockIdx.y;
  const Tap tp = a.taps[t];
  const long long M = (long long)a.N * a.OH * a.OW;
  const long long p_begin = (long long)chunk * a.pix_per_chunk;
  long long p_end = p_begin + a.pix_per_chunk;
  if (p_end > M) p_end = M;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int tco = tid >> 4, tci = tid & 15;

  // loader mapping: two (pixel, 4-float chunk) slots per thread per operand
  float4 av[2], bv[2];
  auto fetch = [&](long long p0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int idx = tid + h * 128;
      const int pl = idx >> 4, c4 = (idx & 15) * 4;
      const long long p = p0 + pl;
      av[h] = bv[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < p_end) {
        if (c4 < a.Cout) av[h] = ldg4(a.dy + p * a.Cout + c4);
        if (c4 < a.Cin) {
          const int ox = (int)(p % a.OW);
          const long long q = p / a.OW;
          const int oy = (int)(q % a.OH);
          const long long n = q / a.OH;
          const float* sp = tp.src ? a.src2 : a.src;
          const int sH = tp.src ? a.src2H : a.srcH, sW = tp.src ? a.src2W : a.srcW, ss = tp.src ? a.in_s2 : a.in_s;
          const int iy = oy * ss + tp.dy, ix = ox * ss + tp.dx;
          if (iy >= 0 && iy < sH && ix >= 0 && ix < sW) bv[h] = ldg4(sp + ((n * sH + iy) * sW + ix) * a.Cin + c4);
        }
      }
    }
  };

  if (p_begin < p_end) fetch(p_begin);
  for (long long p0 = p_begin; p0 < p_end; p0 += BK) {
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int idx = tid + h * 128;
      const int pl = idx >> 4, c4 = (idx & 15) * 4;
      *reinterpret_cast<float4*>(&As[pl][c4]) = av[h];
      *reinterpret_cast<float4*>(&Bs[pl][c4]) = bv[h];
    }
    __syncthreads();
    if (p0 + BK < p_end) fetch(p0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][tco * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][tco * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tci * 4]);
      const float ar[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float br[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
  if (tci * 4 >= a.Cin) return;
  float* po = a.part + ((long long)chunk * a.ntaps + t) * a.Cout * a.Cin;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int co = tco * 8 + i;
    if (co < a.Cout)
      *reinterpret_cast<float4*>(po + (long long)co * a.Cin + tci * 4) =
          make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

}  // namespace

int launch_tapconv_simt(const TapConvArgs& a, cudaStream_t st) {
  if (a.Cin % BK != 0 || a.Cout % 8 != 0 || a.Cout > BN || a.ntaps > kMaxTaps) return B200NP_E_UNSUPPORTED;
  const long long M = (long long)a.N * a.OH * a.OW;
  if (M <= 0) return B200NP_OK;
  tapconv_simt_kernel<<<(unsigned)ceil_div(M, BM), 128, 0, st>>>(a);
  return launch_status();
}

int launch_tapwgrad_simt(const TapWgradArgs& a, cudaStream_t st) {
  if (a.Cin % 4 != 0 || a.Cout % 4 != 0 || a.Cin > 64 || a.Cout > 64 || a.ntaps > kMaxTaps || a.ntaps < 1)
    return B200NP_E_UNSUPPORTED;
  dim3 grid(a.chunks, a.ntaps);
  tapwgrad_simt_kernel<<<grid, 128, 0, st>>>(a);
  return launch_status();
}

}  // namespace b200np
