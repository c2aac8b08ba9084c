// Small-Cin direct convolution: the 5x5 stride-2 stem with 1 or 3 input channels and the 3x3 stride-2
// first conv of encoder_w0.  K = Cin*R*R is 9..75 -- far too thin for a tensor-core tile and 12-32
// FLOP/B -- so these run on CUDA cores and are bound by the NHWC output stream (forward: 1 MB per
// 128x128 image) or the dY read (weight gradient).  Input is NCHW exactly as the datasets deliver it.
//
// Both kernels stage the input patch of an (8|4) x 32 output tile in shared memory and give each
// thread 4 output channels, so that 16 consecutive lanes cover the 64 channels (256 contiguous
// bytes) of one NHWC pixel: every global access of a warp is a full 128-byte line.
#include "common.cuh"

using namespace b200np;

namespace b200np {  // conv_stem_umma.cu
bool stem_umma_supported(int Cin, int R, int Cout, int precision);
int launch_stem_fwd_umma(const float* x, const float* w, const float* bias, float* y, uint32_t* relu_bits, int N, int H,
                         int W, int relu, int precision, cudaStream_t st);
size_t stem_wgrad_umma_workspace(int N, int H, int W);
int launch_stem_wgrad_umma(const float* x, const float* dy, float* dw, float* db, int N, int H, int W, int precision,
                           void* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace b200np

namespace {

constexpr int kTW = 32;  // output tile width

// ---------------------------------------------------------------------------------------------
// forward: y[n,oy,ox,:] = act(b + sum_{ci,r,s} w[:,ci,r,s] * x[n,ci,oy*2+r-pad,ox*2+s-pad])
// Persistent 128-thread CTAs walk (2 x 32)-pixel output tiles.  thread = (channel group cg = 4 couts,
// pixel-pair lane): the weights of its 4 couts stay in registers for the whole kernel, and a thread owns
// PAIRS of horizontally adjacent output pixels whose stride-2 patches overlap -- one patch row of both
// pixels is two aligned 16-byte shared-memory loads feeding 2*R*4 FMAs.  16 consecutive lanes write the
// 256 contiguous bytes of one NHWC pixel.
// ---------------------------------------------------------------------------------------------
constexpr int kFwdTH = 2, kFwdThreads = 128;
template <int CIN, int R>
__global__ void __launch_bounds__(kFwdThreads) conv_small_fwd_kernel(const float* __restrict__ x,
                                                                    const float* __restrict__ w,
                                                                    const float* __restrict__ bias,
                                                                    float* __restrict__ y, int N, int H, int W,
                                                                    int Cout, int pad, int relu, long long tiles) {
  constexpr int RR = R * R;
  constexpr int PH = (kFwdTH - 1) * 2 + R, PW = (kTW - 1) * 2 + R;
  constexpr int PWp = (PW + 1 + 3) & ~3;           // pitch multiple of 4 floats (+1: the pair window reads 8)
  constexpr int NPAIR = kFwdTH * kTW / 2;          // 32 pixel pairs per tile
  constexpr int MAXP = NPAIR * 16 / kFwdThreads;   // pairs per thread when Cout = 64 (half of them for Cout = 32)
  __shared__ __align__(16) float patch[CIN][PH][PWp];
  const int OH = H / 2, OW = W / 2;
  const int tiles_x = (OW + kTW - 1) / kTW, tiles_y = (OH + kFwdTH - 1) / kFwdTH;
  const int tid = threadIdx.x;
  const int cgs = Cout / 4;                        // 16 (Cout 64) or 8 (Cout 32) channel groups
  const int cg = tid % cgs, pl = tid / cgs, npl = kFwdThreads / cgs;
  const int ppt = NPAIR / npl;
  const float4 b4 = ldg4(bias + cg * 4);

  // This thread's 4 output channels.  One input channel (stem of the 1-channel tasks): the 25 x 4 weights
  // stay in registers for the whole kernel.  Three input channels would need 300 registers, so they are
  // kept in shared memory and re-read (25 16-byte loads per input channel and tile, against 800 FMAs).
  __shared__ __align__(16) float wsm[CIN > 1 ? CIN * RR * 64 : 4];
  float wr[RR][4];
  if (CIN > 1) {
    for (int i = tid; i < CIN * RR * Cout; i += kFwdThreads) {
      const int co = i / (CIN * RR), rem = i - co * CIN * RR;  // torch layout [co][ci][k]
      wsm[rem * Cout + co] = __ldg(w + i);                       // -> [ci][k][co]
    }
  } else {
#pragma unroll
    for (int k = 0; k < RR; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) wr[k][e] = __ldg(w + (long long)(cg * 4 + e) * RR + k);
  }

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int xt = (int)(tile % tiles_x);
    const long long q = tile / tiles_x;
    const int yt = (int)(q % tiles_y), n = (int)(q / tiles_y);
    const int oy0 = yt * kFwdTH, ox0 = xt * kTW;
    const int iy0 = oy0 * 2 - pad, ix0 = ox0 * 2 - pad;
    __syncthreads();
    for (int i = tid; i < CIN * PH * PWp; i += kFwdThreads) {
      const int ci = i / (PH * PWp), rem = i - ci * PH * PWp;
      const int py = rem / PWp, px = rem - py * PWp;
      const int iy = iy0 + py, ix = ix0 + px;
      float v = 0.f;
      if (px < PW && iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + ((long long)(n * CIN + ci) * H + iy) * W + ix);
      patch[ci][py][px] = v;
    }
    __syncthreads();
    float acc[MAXP][2][4];
#pragma unroll
    for (int i = 0; i < MAXP; ++i)
#pragma unroll
      for (int pq = 0; pq < 2; ++pq) { acc[i][pq][0] = b4.x; acc[i][pq][1] = b4.y; acc[i][pq][2] = b4.z; acc[i][pq][3] = b4.w; }
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      if (CIN > 1) {
#pragma unroll
        for (int k = 0; k < RR; ++k) {
          const float4 w4 = *reinterpret_cast<const float4*>(&wsm[(ci * RR + k) * Cout + cg * 4]);
          wr[k][0] = w4.x; wr[k][1] = w4.y; wr[k][2] = w4.z; wr[k][3] = w4.w;
        }
      }
#pragma unroll
      for (int i = 0; i < MAXP; ++i) {
        if (i < ppt) {
          const int p = pl + i * npl;              // pair index inside the tile: row p / 16, pixels 2*(p%16), +1
          const int ty = p / (kTW / 2), tp = p - ty * (kTW / 2);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float4* row = reinterpret_cast<const float4*>(&patch[ci][ty * 2 + r][tp * 4]);
            const float4 v0 = row[0], v1 = row[1];
            const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int sx = 0; sx < R; ++sx)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                acc[i][0][e] = fmaf(v[sx], wr[r * R + sx][e], acc[i][0][e]);
                acc[i][1][e] = fmaf(v[2 + sx], wr[r * R + sx][e], acc[i][1][e]);
              }
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
      if (i < ppt) {
        const int p = pl + i * npl;
        const int ty = p / (kTW / 2), tp = p - ty * (kTW / 2);
        const int oy = oy0 + ty;
#pragma unroll
        for (int pq = 0; pq < 2; ++pq) {
          const int ox = ox0 + tp * 2 + pq;
          if (oy < OH && ox < OW) {
            float4 o = make_float4(acc[i][pq][0], acc[i][pq][1], acc[i][pq][2], acc[i][pq][3]);
            if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4*>(y + ((long long)(n * OH + oy) * OW + ox) * Cout + cg * 4) = o;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient: dw[co,k] = sum_pixels dy[pix,co] * patch[pix,k];  db[co] = sum dy[pix,co]
// (db rides along as the extra "tap" k = KK whose patch value is 1).
// Persistent CTAs walk 4 x 32 output tiles; thread = (8 couts, 8 taps, pixel lane) keeps an 8x8 partial in
// registers across all its tiles (2 x 16-byte + 8 scalar shared loads per 64 FMAs); the shared reducer
// folds lanes and CTAs deterministically.
// ---------------------------------------------------------------------------------------------
constexpr int kWgTH = 4;
template <int CIN, int R>
__global__ void __launch_bounds__(256) conv_small_wgrad_stage1(const float* __restrict__ x,
                                                              const float* __restrict__ dy,
                                                              float* __restrict__ part, int N, int H, int W,
                                                              int Cout, int pad, int KG, int nlanes,
                                                              long long tiles) {
  constexpr int RR = R * R, KK = CIN * RR;
  constexpr int PH = (kWgTH - 1) * 2 + R, PW = (kTW - 1) * 2 + R;
  constexpr int PWp = PW | 1;
  extern __shared__ float smem[];
  float* patch = smem;                                  // [CIN][PH][PWp] (+1 float holding 1.0f for db)
  float* dys = smem + ((CIN * PH * PWp + 1 + 3) & ~3);   // [128][Cout]
  const int OH = H / 2, OW = W / 2;
  const int tiles_x = (OW + kTW - 1) / kTW, tiles_y = (OH + kWgTH - 1) / kWgTH;
  const int tid = threadIdx.x;
  const int cgs = Cout / 8;                              // groups of 8 output channels
  const int lane_threads = cgs * KG;
  const int lane_id = tid / lane_threads, lt = tid - lane_id * lane_threads;
  const bool active = lane_id < nlanes;
  const int cg = lt % cgs, kg = lt / cgs;
  // patch offsets of this thread's 8 taps (relative to the pixel's patch origin); tap KK -> the 1.0f slot
  const int one_slot = CIN * PH * PWp;
  int koff[8];
  bool kvar[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = kg * 8 + e;
    kvar[e] = k < KK;
    if (k < KK) {
      const int ci = k / RR, rs = k - ci * RR;
      koff[e] = (ci * PH + rs / R) * PWp + rs % R;
    } else {
      koff[e] = one_slot;  // k == KK reads the constant 1 (bias gradient); k > KK is dropped by the reducer
    }
  }
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[i][e] = 0.f;
  if (tid == 0) patch[one_slot] = 1.0f;

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int xt = (int)(tile % tiles_x);
    const long long q = tile / tiles_x;
    const int yt = (int)(q % tiles_y), n = (int)(q / tiles_y);
    const int oy0 = yt * kWgTH, ox0 = xt * kTW;
    const int iy0 = oy0 * 2 - pad, ix0 = ox0 * 2 - pad;
    __syncthreads();
    for (int i = tid; i < CIN * PH * PW; i += blockDim.x) {
      const int ci = i / (PH * PW), rem = i - ci * PH * PW;
      const int py = rem / PW, px = rem - py * PW;
      const int iy = iy0 + py, ix = ix0 + px;
      float v = 0.f;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + ((long long)(n * CIN + ci) * H + iy) * W + ix);
      patch[(ci * PH + py) * PWp + px] = v;
    }
    const int c4 = Cout / 4;
    for (int i = tid; i < kWgTH * kTW * c4; i += blockDim.x) {
      const int p = i / c4, c = i - p * c4;
      const int ty = p / kTW, tx = p - ty * kTW;
      const int oy = oy0 + ty, ox = ox0 + tx;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (oy < OH && ox < OW) v = ldg4(dy + ((long long)(n * OH + oy) * OW + ox) * Cout + c * 4);
      reinterpret_cast<float4*>(dys)[i] = v;
    }
    __syncthreads();
    if (active) {
      for (int p = lane_id; p < kWgTH * kTW; p += nlanes) {
        const int ty = p / kTW, tx = p - ty * kTW;
        const float4 g0 = reinterpret_cast<const float4*>(dys)[p * c4 + cg * 2];
        const float4 g1 = reinterpret_cast<const float4*>(dys)[p * c4 + cg * 2 + 1];
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const int base = (ty * 2) * PWp + tx * 2;
        float pv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) pv[e] = patch[kvar[e] ? base + koff[e] : koff[e]];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[i][e] = fmaf(g[i], pv[e], acc[i][e]);
      }
    }
  }
  // partial layout: [cta][lane][co][KG*8]
  if (active) {
    float* po = part + (((long long)blockIdx.x * nlanes + lane_id) * Cout) * (KG * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float* row = po + (long long)(cg * 8 + i) * (KG * 8) + kg * 8;
      *reinterpret_cast<float4*>(row) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4*>(row + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
  }
}

struct WgCfg { int KG, nlanes, threads, blocks; long long tiles; size_t smem, ws; };
WgCfg wg_cfg(int N, int Cin, int H, int W, int Cout, int R) {
  WgCfg c;
  const int KK = Cin * R * R;
  c.KG = (KK + 1 + 7) / 8;                       // groups of 8 taps (incl. the bias pseudo-tap)
  const int lane_threads = (Cout / 8) * c.KG;
  c.nlanes = 256 / lane_threads < 1 ? 1 : 256 / lane_threads;
  if (c.nlanes > kWgTH * kTW) c.nlanes = kWgTH * kTW;
  c.threads = ((c.nlanes * lane_threads + 31) / 32) * 32;
  const int OH = H / 2, OW = W / 2;
  c.tiles = (long long)N * ((OH + kWgTH - 1) / kWgTH) * ((OW + kTW - 1) / kTW);
  long long b = 2LL * kNumSMs;
  c.blocks = (int)(c.tiles < b ? c.tiles : b);
  const int PH = (kWgTH - 1) * 2 + R, PW = (kTW - 1) * 2 + R, PWp = PW | 1;
  c.smem = (size_t)(((Cin * PH * PWp + 1 + 3) & ~3) + kWgTH * kTW * Cout) * sizeof(float);
  c.ws = (size_t)c.blocks * c.nlanes * Cout * c.KG * 8 * sizeof(float);
  return c;
}

bool supported(int Cin, int R, int Cout, int stride, int pad, int H, int W) {
  bool combo = (Cin == 1 && R == 5) || (Cin == 3 && R == 5) || (Cin == 1 && R == 3);
  return combo && (Cout == 32 || Cout == 64) && stride == 2 && pad == R / 2 && H % 2 == 0 && W % 2 == 0 &&
         H >= 2 && W >= 2;
}

}  // namespace

extern "C" int b200np_conv_small_fwd(const float* x, const float* w, const float* bias, float* y, int N, int Cin,
                                     int H, int W, int Cout, int R, int stride, int pad, int relu, int precision,
                                     uint32_t* relu_bits, void* stream) {
  if (!x || !w || !bias || !y || N <= 0) return B200NP_E_BADARG;
  if (!supported(Cin, R, Cout, stride, pad, H, W)) return B200NP_E_UNSUPPORTED;
  if (!aligned16(y) || !aligned16(bias)) return B200NP_E_BADARG;
  const int OH = H / 2, OW = W / 2;
  cudaStream_t st = as_stream(stream);
  if (stem_umma_supported(Cin, R, Cout, precision))
    return launch_stem_fwd_umma(x, w, bias, y, relu_bits, N, H, W, relu, precision, st);
  if (relu_bits) return B200NP_E_UNSUPPORTED;  // only the tcgen05 stem writes the 1-bit gates
  const long long tiles = (long long)N * ((OH + kFwdTH - 1) / kFwdTH) * ((OW + kTW - 1) / kTW);
  const long long cap = 8LL * kNumSMs;
  const int grid = (int)(tiles < cap ? tiles : cap);
  if (Cin == 1 && R == 5) conv_small_fwd_kernel<1, 5><<<grid, kFwdThreads, 0, st>>>(x, w, bias, y, N, H, W, Cout, pad, relu, tiles);
  else if (Cin == 3 && R == 5) conv_small_fwd_kernel<3, 5><<<grid, kFwdThreads, 0, st>>>(x, w, bias, y, N, H, W, Cout, pad, relu, tiles);
  else conv_small_fwd_kernel<1, 3><<<grid, kFwdThreads, 0, st>>>(x, w, bias, y, N, H, W, Cout, pad, relu, tiles);
  return launch_status();
}

extern "C" size_t b200np_conv_small_wgrad_workspace(int N, int Cin, int H, int W, int Cout, int R, int stride,
                                                    int pad, int precision) {
  (void)pad;
  if (N <= 0 || stride != 2) return 0;
  if (stem_umma_supported(Cin, R, Cout, precision)) return stem_wgrad_umma_workspace(N, H, W);
  return wg_cfg(N, Cin, H, W, Cout, R).ws;
}

extern "C" int b200np_conv_small_wgrad(const float* x, const float* dy, float* dw, float* db, int N, int Cin, int H,
                                       int W, int Cout, int R, int stride, int pad, int precision, void* ws,
                                       size_t ws_bytes, void* stream) {
  if (!x || !dy || !dw || N <= 0) return B200NP_E_BADARG;
  if (!supported(Cin, R, Cout, stride, pad, H, W)) return B200NP_E_UNSUPPORTED;
  if (!aligned16(dy) || !aligned16(ws)) return B200NP_E_BADARG;
  if (stem_umma_supported(Cin, R, Cout, precision))
    return launch_stem_wgrad_umma(x, dy, dw, db, N, H, W, precision, ws, ws_bytes, as_stream(stream));
  const WgCfg c = wg_cfg(N, Cin, H, W, Cout, R);
  if (!ws || ws_bytes < c.ws) return B200NP_E_WORKSPACE;
  cudaStream_t st = as_stream(stream);
#define LAUNCH(CI, RR)                                                                                          \
  do {                                                                                                          \
    if (c.smem > 48 * 1024 &&                                                                                   \
        cudaFuncSetAttribute(conv_small_wgrad_stage1<CI, RR>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                             (int)c.smem) != cudaSuccess)                                                       \
      return B200NP_E_LAUNCH;                                                                                   \
    conv_small_wgrad_stage1<CI, RR><<<c.blocks, c.threads, c.smem, st>>>(x, dy, (float*)ws, N, H, W, Cout, pad, \
                                                                         c.KG, c.nlanes, c.tiles);              \
  } while (0)
  if (Cin == 1 && R == 5) LAUNCH(1, 5);
  else if (Cin == 3 && R == 5) LAUNCH(3, 5);
  else LAUNCH(1, 3);
#undef LAUNCH
  const int KK = Cin * R * R;
  int rc = launch_status(1);
  if (rc != B200NP_OK) return rc;
  return launch_reduce_partials((const float*)ws, dw, c.blocks * c.nlanes, Cout * c.KG * 8,
                                ReduceMap{2, KK, c.KG * 8, 0, db}, st);
}
