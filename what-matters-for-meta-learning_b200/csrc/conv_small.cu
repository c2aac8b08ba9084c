// Small-Cin direct convolution (the 5x5 stride-2 stem with 1 or 3 input channels, and the
// 3x3 stride-2 first conv of encoder_w0).  K = Cin*R*R is 9..75: far too thin for a tensor-core
// tile and 12-32 FLOP/B, so these run on CUDA cores and are bound by the NHWC output stream
// (forward) / the dY read (wgrad).  Input is NCHW exactly as the datasets deliver it.
#include "common.cuh"

using namespace b200np;

namespace {

constexpr int kTile = 16;  // 16x16 output pixels per CTA

// ---------------------------------------------------------------------------------------------
// forward: y[n,oy,ox,:] = act(b + sum_{ci,r,s} w[:,ci,r,s] * x[n,ci,oy*2+r-pad,ox*2+s-pad])
// ---------------------------------------------------------------------------------------------
template <int CIN, int R>
__global__ void __launch_bounds__(256) conv_small_fwd_kernel(const float* __restrict__ x,
                                                            const float* __restrict__ w,
                                                            const float* __restrict__ bias,
                                                            float* __restrict__ y, int H, int W, int Cout,
                                                            int stride, int pad, int relu) {
  constexpr int KK = CIN * R * R;
  constexpr int PH = (kTile - 1) * 2 + R;  // patch extent for stride 2 (stride 1 needs less)
  constexpr int PW = PH + 1;               // +1: odd row pitch
  extern __shared__ float smem[];
  float* patch = smem;                     // [CIN][PH][PW]
  float* ws = smem + ((CIN * PH * PW + 3) & ~3);  // [KK][Cout], 16 B aligned
  const int OH = H / stride, OW = W / stride;
  const int n = blockIdx.z;
  const int oy0 = blockIdx.y * kTile, ox0 = blockIdx.x * kTile;
  const int tid = threadIdx.x;

  for (int i = tid; i < KK * Cout; i += 256) {  // transpose weights to [k][co]
    int co = i / KK, k = i - co * KK;
    ws[k * Cout + co] = __ldg(w + i);
  }
  const int iy0 = oy0 * stride - pad, ix0 = ox0 * stride - pad;
  const int ph = (kTile - 1) * stride + R;
  for (int i = tid; i < CIN * ph * ph; i += 256) {
    int ci = i / (ph * ph), rem = i - ci * ph * ph;
    int py = rem / ph, px = rem - py * ph;
    int iy = iy0 + py, ix = ix0 + px;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + ((long long)(n * CIN + ci) * H + iy) * W + ix);
    patch[(ci * PH + py) * PW + px] = v;
  }
  __syncthreads();

  const int ty = tid / kTile, tx = tid % kTile;
  const int oy = oy0 + ty, ox = ox0 + tx;
  if (oy >= OH || ox >= OW) return;
  float* yo = y + ((long long)(n * OH + oy) * OW + ox) * Cout;
  for (int c0 = 0; c0 < Cout; c0 += 16) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = __ldg(bias + c0 + j);
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int s = 0; s < R; ++s) {
          float v = patch[(ci * PH + ty * stride + r) * PW + tx * stride + s];
          const float4* wp = reinterpret_cast<const float4*>(ws + ((ci * R + r) * R + s) * Cout + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 wv = wp[q];
            acc[q * 4 + 0] = fmaf(v, wv.x, acc[q * 4 + 0]);
            acc[q * 4 + 1] = fmaf(v, wv.y, acc[q * 4 + 1]);
            acc[q * 4 + 2] = fmaf(v, wv.z, acc[q * 4 + 2]);
            acc[q * 4 + 3] = fmaf(v, wv.w, acc[q * 4 + 3]);
          }
        }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 o = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
      if (relu) {
        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
      }
      reinterpret_cast<float4*>(yo + c0)[q] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad: dw[co,k] = sum_pixels dy[pix,co] * patch[pix,k];  db[co] = sum dy[pix,co]
// Stage 1: each CTA walks tiles of 64 output pixels, keeps a [Cout][KKp] partial in registers
// (thread = one co x a strided set of 4-wide k groups), writes it to ws.  Stage 2 folds CTAs.
// ---------------------------------------------------------------------------------------------
constexpr int kWgPix = 64;
template <int CIN, int R>
__global__ void __launch_bounds__(256) conv_small_wgrad_stage1(const float* __restrict__ x,
                                                              const float* __restrict__ dy,
                                                              float* __restrict__ part, int N, int H, int W,
                                                              int Cout, int stride, int pad, long long tiles) {
  constexpr int KK = CIN * R * R;
  constexpr int KKp = (KK + 15) / 16 * 16;
  constexpr int NJ = KKp / 16;  // float4 groups per thread
  extern __shared__ float smem[];
  float* dys = smem;                       // [64][Cout]
  float* ps = smem + kWgPix * Cout;        // [64][KKp]
  const int OH = H / stride, OW = W / stride;
  const long long npix = (long long)N * OH * OW;
  const int tid = threadIdx.x;
  const int kg_count = 256 / Cout;         // k-groups per pixel row handled in parallel (4 or 8)
  const int co = tid % Cout, kg = tid / Cout;
  float acc[NJ][4];
  float accb = 0.f;
#pragma unroll
  for (int j = 0; j < NJ; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long p0 = tile * kWgPix;
    __syncthreads();
    for (int i = tid; i < kWgPix * Cout; i += 256) {
      long long p = p0 + i / Cout;
      dys[i] = p < npix ? __ldg(dy + p0 * Cout + i) : 0.f;
    }
    for (int i = tid; i < kWgPix * KKp; i += 256) {
      int pl = i / KKp, k = i - pl * KKp;
      long long p = p0 + pl;
      float v = 0.f;
      if (p < npix && k < KK) {
        int ox = (int)(p % OW);
        long long q = p / OW;
        int oy = (int)(q % OH);
        int n = (int)(q / OH);
        int ci = k / (R * R), rs = k - ci * R * R;
        int r = rs / R, s = rs - r * R;
        int iy = oy * stride + r - pad, ix = ox * stride + s - pad;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + ((long long)(n * CIN + ci) * H + iy) * W + ix);
      }
      ps[i] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int pl = 0; pl < kWgPix; ++pl) {
      float g = dys[pl * Cout + co];
      if (kg == 0) accb += g;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        int kq = kg + j * kg_count;  // 4-wide group index
        if (kq * 4 < KKp) {
          float4 pv = *reinterpret_cast<const float4*>(ps + pl * KKp + kq * 4);
          acc[j][0] = fmaf(g, pv.x, acc[j][0]);
          acc[j][1] = fmaf(g, pv.y, acc[j][1]);
          acc[j][2] = fmaf(g, pv.z, acc[j][2]);
          acc[j][3] = fmaf(g, pv.w, acc[j][3]);
        }
      }
    }
  }
  float* po = part + (long long)blockIdx.x * Cout * (KKp + 1);
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    int kq = kg + j * kg_count;
    if (kq * 4 < KKp) {
#pragma unroll
      for (int e = 0; e < 4; ++e) po[co * (KKp + 1) + kq * 4 + e] = acc[j][e];
    }
  }
  if (kg == 0) po[co * (KKp + 1) + KKp] = accb;
}

__global__ void conv_small_wgrad_stage2(const float* __restrict__ part, float* __restrict__ dw,
                                        float* __restrict__ db, int Cout, int KK, int KKp, int blocks) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int total = Cout * (KK + 1);
  if (i >= total) return;
  int co = i / (KK + 1), k = i - co * (KK + 1);
  int kp = k < KK ? k : KKp;
  float s = 0.f;
  for (int b = 0; b < blocks; ++b) s += part[(long long)b * Cout * (KKp + 1) + co * (KKp + 1) + kp];
  if (k < KK) dw[co * KK + k] = s;
  else if (db) db[co] = s;
}

int wgrad_blocks(long long tiles) {
  long long b = 2LL * kNumSMs;
  return (int)(tiles < b ? tiles : b);
}
bool supported(int Cin, int R, int Cout, int stride, int pad, int H, int W) {
  bool combo = (Cin == 1 && R == 5) || (Cin == 3 && R == 5) || (Cin == 1 && R == 3);
  return combo && (Cout == 32 || Cout == 64) && stride == 2 && pad == R / 2 && H % 2 == 0 && W % 2 == 0 &&
         H >= 2 && W >= 2;
}

}  // namespace

extern "C" int b200np_conv_small_fwd(const float* x, const float* w, const float* bias, float* y, int N, int Cin,
                                     int H, int W, int Cout, int R, int stride, int pad, int relu, void* stream) {
  if (!x || !w || !bias || !y || N <= 0) return B200NP_E_BADARG;
  if (!supported(Cin, R, Cout, stride, pad, H, W)) return B200NP_E_UNSUPPORTED;
  if (!aligned16(y)) return B200NP_E_BADARG;
  int OH = H / stride, OW = W / stride;
  dim3 grid((OW + kTile - 1) / kTile, (OH + kTile - 1) / kTile, N);
  int PH = (kTile - 1) * 2 + R;
  size_t smem = (size_t)(((Cin * PH * (PH + 1) + 3) & ~3) + Cin * R * R * Cout) * sizeof(float);
  cudaStream_t st = as_stream(stream);
#define LAUNCH(CI, RR)                                                                              \
  conv_small_fwd_kernel<CI, RR><<<grid, 256, smem, st>>>(x, w, bias, y, H, W, Cout, stride, pad, relu)
  if (Cin == 1 && R == 5) LAUNCH(1, 5);
  else if (Cin == 3 && R == 5) LAUNCH(3, 5);
  else LAUNCH(1, 3);
#undef LAUNCH
  return launch_status();
}

extern "C" size_t b200np_conv_small_wgrad_workspace(int N, int Cin, int H, int W, int Cout, int R, int stride,
                                                    int pad) {
  (void)pad;
  if (N <= 0 || stride <= 0) return 0;
  long long npix = (long long)N * (H / stride) * (W / stride);
  long long tiles = ceil_div(npix, kWgPix);
  int KKp = (Cin * R * R + 15) / 16 * 16;
  return (size_t)wgrad_blocks(tiles) * Cout * (KKp + 1) * sizeof(float);
}

extern "C" int b200np_conv_small_wgrad(const float* x, const float* dy, float* dw, float* db, int N, int Cin, int H,
                                       int W, int Cout, int R, int stride, int pad, void* ws, size_t ws_bytes,
                                       void* stream) {
  if (!x || !dy || !dw || N <= 0) return B200NP_E_BADARG;
  if (!supported(Cin, R, Cout, stride, pad, H, W)) return B200NP_E_UNSUPPORTED;
  if (!ws || ws_bytes < b200np_conv_small_wgrad_workspace(N, Cin, H, W, Cout, R, stride, pad))
    return B200NP_E_WORKSPACE;
  long long npix = (long long)N * (H / stride) * (W / stride);
  long long tiles = ceil_div(npix, kWgPix);
  int blocks = wgrad_blocks(tiles);
  int KK = Cin * R * R, KKp = (KK + 15) / 16 * 16;
  size_t smem = (size_t)(kWgPix * Cout + kWgPix * KKp) * sizeof(float);
  cudaStream_t st = as_stream(stream);
#define LAUNCH(CI, RR)                                                                             \
  conv_small_wgrad_stage1<CI, RR><<<blocks, 256, smem, st>>>(x, dy, (float*)ws, N, H, W, Cout, stride, pad, tiles)
  if (Cin == 1 && R == 5) LAUNCH(1, 5);
  else if (Cin == 3 && R == 5) LAUNCH(3, 5);
  else LAUNCH(1, 3);
#undef LAUNCH
  int total = Cout * (KK + 1);
  conv_small_wgrad_stage2<<<(total + 127) / 128, 128, 0, st>>>((const float*)ws, dw, db, Cout, KK, KKp, blocks);
  return launch_status(2);
}
