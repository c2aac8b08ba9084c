// Building blocks of the MMAML conv nets: networks/gated_conv_net.py:167-212 (GatedConvModel:
// 3x3 stride-2 conv -> batch-statistics BatchNorm (training=True always, no affine) -> FiLM x*(1+gamma)+beta -> ReLU,
// 32/64/128/256 channels) and networks/conv_embedding_model.py:99-184 (same stack with affine BatchNorm).
//
//   * im2col / col2im for 3x3 stride-2 pad-1 convolutions on NHWC tensors with the column order k = ci*9 + tap, i.e.
//     the flattening of torch's [Cout, Cin, 3, 3] weight: forward, weight gradient and data gradient are then plain
//     strided GEMMs on the tcgen05 kernel (b200np_gemm) reading the parameter tensors as they are -- any channel count;
//   * batch statistics per channel over all N*H*W rows (two-level, deterministic; sums and sums of squares in fp64, the
//     mean handed on as a (hi, lo) float pair -- see bn_stats_l1_kernel), the fused normalise + scale/shift + ReLU, and its backward (two more column reductions + one
//     elementwise pass).  scale/shift cover both users: FiLM (scale = 1 + gamma, shift = beta) and affine BatchNorm
//     (scale = weight, shift = bias).
#include "common.cuh"

namespace b200np {
namespace {

// ---------------------------------------------------------------------------------------------------------------
// im2col: x [N, H, W, C] -> col [N*OH*OW, C*9], col[m][ci*9 + r*3 + s] = x[n, 2*oy + r - 1, 2*ox + s - 1, ci]
// thread = (pixel m, channel ci): 9 loads coalesced over ci, 36 contiguous bytes stored.
// ---------------------------------------------------------------------------------------------------------------
__global__ void im2col3x3s2_kernel(const float* __restrict__ x, float* __restrict__ col, int N, int H, int W, int C) {
  const int OH = H / 2, OW = W / 2;
  const long long total = (long long)N * OH * OW * C;
  const long long st = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += st) {
    const int ci = (int)(i % C);
    const long long m = i / C;
    const int ox = (int)(m % OW);
    const long long q = m / OW;
    const int oy = (int)(q % OH), n = (int)(q / OH);
    float v[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int iy = 2 * oy + r - 1, ix = 2 * ox + s - 1;
        v[r * 3 + s] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(x + (((long long)n * H + iy) * W + ix) * C + ci) : 0.f;
      }
    float* o = col + (m * C + ci) * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) o[t] = v[t];
  }
}
// col2im (gather form, deterministic): dx[n, y, x, ci] = sum over the taps (r, s) and outputs (oy, ox) with
// 2*oy + r - 1 == y, 2*ox + s - 1 == x of dcol[m(n, oy, ox)][ci*9 + r*3 + s]; optional gate (mask > 0).
__global__ void col2im3x3s2_kernel(const float* __restrict__ dcol, const float* __restrict__ mask, float* __restrict__ dx,
                                   int N, int H, int W, int C) {
  const int OH = H / 2, OW = W / 2;
  const long long total = (long long)N * H * W * C;
  const long long st = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += st) {
    const int ci = (int)(i % C);
    const long long p = i / C;
    const int xx = (int)(p % W);
    const long long q = p / W;
    const int yy = (int)(q % H), n = (int)(q / H);
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ty = yy + 1 - r;
      if (ty < 0 || (ty & 1)) continue;
      const int oy = ty >> 1;
      if (oy >= OH) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int tx = xx + 1 - s;
        if (tx < 0 || (tx & 1)) continue;
        const int ox = tx >> 1;
        if (ox >= OW) continue;
        acc += __ldg(dcol + ((((long long)n * OH + oy) * OW + ox) * C + ci) * 9 + r * 3 + s);
      }
    }
    if (mask && !(__ldg(mask + i) > 0.f)) acc = 0.f;
    dx[i] = acc;
  }
}

// im2col of an NCHW image with few input channels for an R x R stride-2 convolution (the stems: 5x5 over 3 channels,
// networks/ResNet.py conv1; 3x3 over 1 channel, encoder_w0[0]): col[m][ci*R*R + r*R + s] = x[n, ci, 2*oy + r - pad, 2*ox + s - pad]
// with row pitch `ld` (>= Cin*R*R, the tail is zero-filled): K = 9..75 is thin, but as a tcgen05 GEMM over this matrix
// the stem and its weight gradient cost a fraction of the CUDA-core direct convolution.  thread = (row m, column k):
// coalesced stores, gathers served by L1 (every input pixel is read R*R/4 times).
__global__ void im2col_small_kernel(const float* __restrict__ x, float* __restrict__ col, int N, int Cin, int H, int W, int R,
                                    int pad, int ld) {
  // threadIdx.y / blockIdx.x: one output row (image n, row oy); threadIdx.x: column k.  Everything that needs a division
  // is done once per thread; the loop over ox walks one input row with stride 2 and stores 4 bytes at pitch ld, so that a
  // warp writes 128 contiguous bytes per iteration.
  const int OH = H / 2, OW = W / 2, RR = R * R, KK = Cin * RR;
  const long long row = (long long)blockIdx.x * blockDim.y + threadIdx.y;
  if (row >= (long long)N * OH) return;
  const int oy = (int)(row % OH);
  const long long n = row / OH;
  for (int k = threadIdx.x; k < ld; k += blockDim.x) {
    float* out = col + row * OW * ld + k;
    bool rowok = false;
    const float* src = x;
    int ix = 0;
    if (k < KK) {
      const int ci = k / RR, rs = k - ci * RR, r = rs / R, sx = rs - r * R;
      const int iy = 2 * oy + r - pad;
      rowok = iy >= 0 && iy < H;
      src = x + ((n * Cin + ci) * H + (rowok ? iy : 0)) * W;
      ix = sx - pad;
    }
    for (int ox = 0; ox < OW; ++ox, ix += 2)
      out[(long long)ox * ld] = (rowok && ix >= 0 && ix < W) ? __ldg(src + ix) : 0.f;
  }
}
// TAP-major variants (k = tap*C + ci) for the implicit-convolution GEMM (b200np_gemm_desc.conv_operand): the channel
// vector of a pixel is contiguous in the column matrix, so everything moves as 16-byte words.
__global__ void conv_weight_tapmajor_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout, int Cin, int fwd) {
  const int total = Cout * Cin * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i / (Cin * 9), rem = i - co * Cin * 9;
    if (fwd) {                       // i indexes dst [co][tap][ci]
      const int tap = rem / Cin, ci = rem - tap * Cin;
      dst[i] = __ldg(src + (co * Cin + ci) * 9 + tap);
    } else {                         // i indexes dst [co][ci][tap]
      const int ci = rem / 9, tap = rem - ci * 9;
      dst[i] = __ldg(src + (co * 9 + tap) * Cin + ci);
    }
  }
}
__global__ void col2im3x3s2_tapmajor_kernel(const float* __restrict__ dcol, const float* __restrict__ mask,
                                            float* __restrict__ dx, int N, int H, int W, int C) {
  const int OH = H / 2, OW = W / 2, C4 = C / 4;
  const long long total = (long long)N * H * W * C4;
  const long long st = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += st) {
    const int c4 = (int)(i % C4);
    const long long p = i / C4;
    const int xx = (int)(p % W);
    const long long q = p / W;
    const int yy = (int)(q % H), n = (int)(q / H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ty = yy + 1 - r;
      if (ty < 0 || (ty & 1)) continue;
      const int oy = ty >> 1;
      if (oy >= OH) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int tx = xx + 1 - s;
        if (tx < 0 || (tx & 1)) continue;
        const int ox = tx >> 1;
        if (ox >= OW) continue;
        const float4 v = ldg4(dcol + (((long long)n * OH + oy) * OW + ox) * (9LL * C) + (r * 3 + s) * C + 4 * c4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    if (mask) {
      const float4 mk = ldg4(mask + p * C + 4 * c4);
      acc.x = mk.x > 0.f ? acc.x : 0.f; acc.y = mk.y > 0.f ? acc.y : 0.f;
      acc.z = mk.z > 0.f ? acc.z : 0.f; acc.w = mk.w > 0.f ? acc.w : 0.f;
    }
    *reinterpret_cast<float4*>(dx + p * C + 4 * c4) = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Column statistics of a [R, C] matrix.  Level 1: block (32 channels x 8 row lanes) over a row slice -> (count, mean,
// M2) per channel, merged across the 8 lanes in shared memory; level 2: one thread per channel merges the slices in
// order (Chan et al.), writes mean / rstd and updates the running statistics like F.batch_norm(training=True) does.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kStatLanes = 8;
// Statistics in fp64: sum and sum of squares per (slice, channel), folded in slice order.  The mean is handed to the
// elementwise kernels as a (hi, lo) pair of floats (mean[c], mean[C + c]) and x - mean is formed as (x - hi) - lo: with
// fp32 Welford statistics the mean carries an error of ~1e-6 standard deviations when |mean| >> std (a conv output with
// a bias), which decides the sign of pre-activations that close to zero -- on a 61440 x 32 map a few ReLU gates differ
// from the exact result, each worth ~1e-3 of the gradient and more of the second-order gradient (measured: torch's own
// fp32 CUDA batch norm shows the same; its CPU kernel accumulates in double and does not).  DESIGN.md finding 25.
__global__ void __launch_bounds__(256) bn_stats_l1_kernel(const float* __restrict__ x, double* __restrict__ part, long long R,
                                                          int C, long long rows_per_slice) {
  __shared__ double s_s[kStatLanes][32], s_q[kStatLanes][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const long long r0 = (long long)blockIdx.y * rows_per_slice;
  const long long r1 = r0 + rows_per_slice < R ? r0 + rows_per_slice : R;
  double sum = 0.0, sq = 0.0;
  if (c < C)
    for (long long r = r0 + rl; r < r1; r += kStatLanes) {
      const double v = (double)__ldg(x + r * C + c);
      sum += v;
      sq = fma(v, v, sq);
    }
  s_s[rl][cl] = sum; s_q[rl][cl] = sq;
  __syncthreads();
  if (rl == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < kStatLanes; ++k) { sum += s_s[k][cl]; sq += s_q[k][cl]; }
    double* o = part + ((long long)blockIdx.y * C + c) * 2;
    o[0] = sum; o[1] = sq;
  }
}
__global__ void bn_stats_l2_kernel(const double* __restrict__ part, int slices, int C, double rows, float eps,
                                   float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ run_mean,
                                   float* __restrict__ run_var, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double sum = 0.0, sq = 0.0;
  for (int s = 0; s < slices; ++s) {
    sum += part[((long long)s * C + c) * 2];
    sq += part[((long long)s * C + c) * 2 + 1];
  }
  const double m = sum / rows;
  double var = sq / rows - m * m;   // biased, as used for normalisation
  if (var < 0.0) var = 0.0;
  const float hi = (float)m;
  mean[c] = hi;
  mean[C + c] = (float)(m - (double)hi);
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (run_mean) run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * hi;
  if (run_var) run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)(rows > 1.0 ? var * rows / (rows - 1.0) : var);
}

// y = relu?( (x - mean) * rstd * scale' + shift ),  scale' = scale + plus_one  (FiLM: 1 + gamma)
__global__ void bn_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                                  const float* __restrict__ scale, const float* __restrict__ shift, float plus_one,
                                  float* __restrict__ y, long long total, int C, int relu) {
  const long long st = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += st) {
    const int c = (int)(i % C);
    const float sc = (scale ? __ldg(scale + c) : 0.f) + plus_one;
    float v = ((__ldg(x + i) - __ldg(mean + c)) - __ldg(mean + C + c)) * __ldg(rstd + c) * sc + (shift ? __ldg(shift + c) : 0.f);
    y[i] = relu ? fmaxf(v, 0.f) : v;
  }
}
// backward, level 1: per slice and channel  sum g  and  sum g * xhat,  g = dy * (y > 0 | !relu)
__global__ void __launch_bounds__(256) bn_act_bwd_l1_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                            const float* __restrict__ x, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, float* __restrict__ part,
                                                            long long R, int C, long long rows_per_slice, int relu) {
  __shared__ float s_a[kStatLanes][32], s_b[kStatLanes][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const long long r0 = (long long)blockIdx.y * rows_per_slice;
  const long long r1 = r0 + rows_per_slice < R ? r0 + rows_per_slice : R;
  float a = 0.f, b = 0.f;
  if (c < C) {
    const float mu = __ldg(mean + c), mlo = __ldg(mean + C + c), rs = __ldg(rstd + c);
    for (long long r = r0 + rl; r < r1; r += kStatLanes) {
      const long long i = r * C + c;
      float g = __ldg(dy + i);
      if (relu && !(__ldg(y + i) > 0.f)) g = 0.f;
      a += g;
      b = fmaf(g, ((__ldg(x + i) - mu) - mlo) * rs, b);
    }
  }
  s_a[rl][cl] = a; s_b[rl][cl] = b;
  __syncthreads();
  if (rl == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < kStatLanes; ++k) { a += s_a[k][cl]; b += s_b[k][cl]; }
    float* o = part + ((long long)blockIdx.y * C + c) * 2;
    o[0] = a; o[1] = b;
  }
}
__global__ void bn_act_bwd_l2_kernel(const float* __restrict__ part, int slices, int C, float* __restrict__ dshift,
                                     float* __restrict__ dscale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int s = 0; s < slices; ++s) { a += part[((long long)s * C + c) * 2]; b += part[((long long)s * C + c) * 2 + 1]; }
  dshift[c] = a;
  dscale[c] = b;
}
// dx = rstd * sc * ( g - mean_r(g) - xhat * mean_r(g * xhat) ),  sc = scale + plus_one
__global__ void bn_act_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x,
                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                        const float* __restrict__ scale, float plus_one, const float* __restrict__ dshift,
                                        const float* __restrict__ dscale, float* __restrict__ dx, long long total, int C,
                                        float inv_rows, int relu) {
  const long long st = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += st) {
    const int c = (int)(i % C);
    float g = __ldg(dy + i);
    if (relu && !(__ldg(y + i) > 0.f)) g = 0.f;
    const float rs = __ldg(rstd + c);
    const float xh = ((__ldg(x + i) - __ldg(mean + c)) - __ldg(mean + C + c)) * rs;
    const float sc = (scale ? __ldg(scale + c) : 0.f) + plus_one;
    dx[i] = rs * sc * (g - __ldg(dshift + c) * inv_rows - xh * __ldg(dscale + c) * inv_rows);
  }
}

// ---- second order: the derivative of the BACKWARD above (MAML differentiates through the inner-loop gradient,
// trainer/meta_learner_reg.py:116-130).  With G = dy * gate, xh = (x - mean) * rstd, s = scale + plus_one the backward is
//   dx = s * rstd * (G - mean_r G - xh * mean_r(G xh)),   dscale = sum_r G xh,   dshift = sum_r G.
// Given cotangents vx (of dx), vs (of dscale), vt (of dshift) and the per-channel means
//   a = <G>, b = <G xh>, e = <vx>, c = <vx xh>, m = <vx G>,  S = m - e a - c b
// the gradients of  L = <vx, dx> + <vs, dscale> + <vt, dshift>  are (checked against autograd in fp64):
//   dL/ddy    = gate * ( s * rstd * (vx - e - xh c) + vs * xh + vt )
//   dL/dx     = -s * rstd^2 * ( S xh + b (vx - e - c xh) + c (G - a - b xh) ) + vs * rstd * (G - a - b xh)
//   dL/dscale = rstd * rows * S
// level 1: per slice and channel the five sums.  S is a difference of products of means (a covariance written with raw
// moments): in fp32 the cancellation cost 1e-2 of the outer gradient on the reference's own initialisation (torch's fp32
// double backward loses 4e-3 there), so the moments are accumulated and combined in fp64 -- 5 doubles per thread in a
// kernel that is bound by reading three fp32 tensors.
__global__ void __launch_bounds__(256) bn_act_bwd2_l1_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                             const float* __restrict__ x, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, const float* __restrict__ vx,
                                                             double* __restrict__ part, long long R, int C,
                                                             long long rows_per_slice, int relu) {
  __shared__ double sm[5][kStatLanes][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const long long r0 = (long long)blockIdx.y * rows_per_slice;
  const long long r1 = r0 + rows_per_slice < R ? r0 + rows_per_slice : R;
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (c < C) {
    const float mu = __ldg(mean + c), mlo = __ldg(mean + C + c), rs = __ldg(rstd + c);
    for (long long r = r0 + rl; r < r1; r += kStatLanes) {
      const long long i = r * C + c;
      float g = __ldg(dy + i);
      if (relu && !(__ldg(y + i) > 0.f)) g = 0.f;
      const double xh = (double)(((__ldg(x + i) - mu) - mlo) * rs), v = vx ? (double)__ldg(vx + i) : 0.0, gd = (double)g;
      acc[0] += gd;
      acc[1] += gd * xh;
      acc[2] += v;
      acc[3] += v * xh;
      acc[4] += v * gd;
    }
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) sm[k][rl][cl] = acc[k];
  __syncthreads();
  if (rl == 0 && c < C) {
    double* o = part + ((long long)blockIdx.y * C + c) * 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      double t = 0.0;
#pragma unroll
      for (int l = 0; l < kStatLanes; ++l) t += sm[k][l][cl];
      o[k] = t;
    }
  }
}
// level 2: slices folded in order -> per channel a, b, e, c and S (stats[c][0..4]), and dL/dscale
__global__ void bn_act_bwd2_l2_kernel(const double* __restrict__ part, int slices, int C, const float* __restrict__ rstd,
                                      double rows, float* __restrict__ stats, float* __restrict__ gscale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double t[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int s = 0; s < slices; ++s)
#pragma unroll
    for (int k = 0; k < 5; ++k) t[k] += part[((long long)s * C + c) * 5 + k];
#pragma unroll
  for (int k = 0; k < 5; ++k) t[k] /= rows;
  const double S = t[4] - t[2] * t[0] - t[3] * t[1];
  stats[c * 5 + 0] = (float)t[0];
  stats[c * 5 + 1] = (float)t[1];
  stats[c * 5 + 2] = (float)t[2];
  stats[c * 5 + 3] = (float)t[3];
  stats[c * 5 + 4] = (float)S;
  if (gscale) gscale[c] = (float)((double)__ldg(rstd + c) * rows * S);
}
__global__ void bn_act_bwd2_apply_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x,
                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                         const float* __restrict__ scale, float plus_one, const float* __restrict__ vx,
                                         const float* __restrict__ vs, const float* __restrict__ vt,
                                         const float* __restrict__ stats, float* __restrict__ gx, float* __restrict__ gdy,
                                         long long total, int C, int relu) {
  const long long st = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += st) {
    const int ch = (int)(i % C);
    const bool open = !relu || __ldg(y + i) > 0.f;
    const float g = open ? __ldg(dy + i) : 0.f;
    const float rs = __ldg(rstd + ch), xh = ((__ldg(x + i) - __ldg(mean + ch)) - __ldg(mean + C + ch)) * rs;
    const float s = (scale ? __ldg(scale + ch) : 0.f) + plus_one;
    const float v = vx ? __ldg(vx + i) : 0.f, ws = vs ? __ldg(vs + ch) : 0.f, wt = vt ? __ldg(vt + ch) : 0.f;
    const float* q = stats + ch * 5;
    const float a = __ldg(q), b = __ldg(q + 1), e = __ldg(q + 2), c = __ldg(q + 3), S = __ldg(q + 4);
    const float vr = v - e - c * xh, gr = g - a - b * xh;
    if (gdy) gdy[i] = open ? s * rs * vr + ws * xh + wt : 0.f;
    if (gx) gx[i] = -s * rs * rs * (S * xh + b * vr + c * gr) + ws * rs * gr;
  }
}

int stat_slices(long long R) {
  long long s = ceil_div(R, 512);
  if (s > 4 * kNumSMs) s = 4 * kNumSMs;
  return (int)(s < 1 ? 1 : s);
}

}  // namespace
}  // namespace b200np

using namespace b200np;

extern "C" int b200np_im2col3x3s2(const float* x, float* col, int N, int H, int W, int C, void* stream) {
  if (!x || !col || N <= 0 || C <= 0 || H < 2 || W < 2 || (H & 1) || (W & 1)) return B200NP_E_BADARG;
  const long long total = (long long)N * (H / 2) * (W / 2) * C;
  im2col3x3s2_kernel<<<ew_grid(total, 256), 256, 0, as_stream(stream)>>>(x, col, N, H, W, C);
  return launch_status();
}
extern "C" int b200np_col2im3x3s2(const float* dcol, const float* mask, float* dx, int N, int H, int W, int C,
                                  void* stream) {
  if (!dcol || !dx || N <= 0 || C <= 0 || H < 2 || W < 2 || (H & 1) || (W & 1)) return B200NP_E_BADARG;
  const long long total = (long long)N * H * W * C;
  col2im3x3s2_kernel<<<ew_grid(total, 256), 256, 0, as_stream(stream)>>>(dcol, mask, dx, N, H, W, C);
  return launch_status();
}
extern "C" int b200np_im2col_small(const float* x, float* col, int N, int Cin, int H, int W, int R, int pad, int ld,
                                   void* stream) {
  if (!x || !col || N <= 0 || Cin <= 0 || R <= 0 || H < 2 || W < 2 || (H & 1) || (W & 1) || ld < Cin * R * R)
    return B200NP_E_BADARG;
  const int tx = ld >= 128 ? 128 : (ld + 31) / 32 * 32, ty = 256 / tx;
  const long long rows = (long long)N * (H / 2);
  im2col_small_kernel<<<(unsigned)ceil_div(rows, ty), dim3(tx, ty), 0, as_stream(stream)>>>(x, col, N, Cin, H, W, R, pad, ld);
  return launch_status();
}
extern "C" int b200np_conv_weight_tapmajor(const float* src, float* dst, int Cout, int Cin, int to_tapmajor, void* stream) {
  if (!src || !dst || Cout <= 0 || Cin <= 0) return B200NP_E_BADARG;
  conv_weight_tapmajor_kernel<<<ew_grid((long long)Cout * Cin * 9, 256), 256, 0, as_stream(stream)>>>(src, dst, Cout, Cin,
                                                                                                      to_tapmajor);
  return launch_status();
}
extern "C" int b200np_col2im3x3s2_tapmajor(const float* dcol, const float* mask, float* dx, int N, int H, int W, int C,
                                           void* stream) {
  if (!dcol || !dx || N <= 0 || C <= 0 || (C & 3) || H < 2 || W < 2 || (H & 1) || (W & 1)) return B200NP_E_BADARG;
  if (!aligned16(dcol) || !aligned16(dx) || (mask && !aligned16(mask))) return B200NP_E_BADARG;
  const long long total = (long long)N * H * W * (C / 4);
  col2im3x3s2_tapmajor_kernel<<<ew_grid(total, 256), 256, 0, as_stream(stream)>>>(dcol, mask, dx, N, H, W, C);
  return launch_status();
}
extern "C" size_t b200np_bn_workspace(long long rows, int C) {
  return (size_t)stat_slices(rows) * (size_t)C * 2 * sizeof(double);
}
extern "C" int b200np_bn_act_fwd(const float* x, const float* scale, const float* shift, float plus_one, float eps,
                                 float* y, float* mean, float* rstd, float* run_mean, float* run_var, float momentum,
                                 long long rows, int C, int relu, void* ws, size_t ws_bytes, void* stream) {
  if (!x || !y || !mean || !rstd || rows <= 0 || C <= 0 || !ws) return B200NP_E_BADARG;
  if (ws_bytes < b200np_bn_workspace(rows, C)) return B200NP_E_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int slices = stat_slices(rows);
  const long long per = ceil_div(rows, slices);
  bn_stats_l1_kernel<<<dim3((C + 31) / 32, slices), 256, 0, st>>>(x, (double*)ws, rows, C, per);
  bn_stats_l2_kernel<<<(C + 127) / 128, 128, 0, st>>>((const double*)ws, slices, C, (double)rows, eps, mean, rstd, run_mean,
                                                      run_var, momentum);
  const long long total = rows * C;
  bn_act_fwd_kernel<<<ew_grid(total, 256), 256, 0, st>>>(x, mean, rstd, scale, shift, plus_one, y, total, C, relu);
  return launch_status(3);
}
extern "C" int b200np_bn_act_bwd(const float* dy, const float* y, const float* x, const float* mean, const float* rstd,
                                 const float* scale, float plus_one, float* dx, float* dscale, float* dshift,
                                 long long rows, int C, int relu, void* ws, size_t ws_bytes, void* stream) {
  if (!dy || !y || !x || !mean || !rstd || !dx || !dscale || !dshift || rows <= 0 || C <= 0 || !ws) return B200NP_E_BADARG;
  if (ws_bytes < b200np_bn_workspace(rows, C)) return B200NP_E_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int slices = stat_slices(rows);
  const long long per = ceil_div(rows, slices);
  bn_act_bwd_l1_kernel<<<dim3((C + 31) / 32, slices), 256, 0, st>>>(dy, y, x, mean, rstd, (float*)ws, rows, C, per, relu);
  bn_act_bwd_l2_kernel<<<(C + 127) / 128, 128, 0, st>>>((const float*)ws, slices, C, dshift, dscale);
  const long long total = rows * C;
  bn_act_bwd_apply_kernel<<<ew_grid(total, 256), 256, 0, st>>>(dy, y, x, mean, rstd, scale, plus_one, dshift, dscale, dx,
                                                               total, C, 1.f / (float)rows, relu);
  return launch_status(3);
}

extern "C" size_t b200np_bn_workspace2(long long rows, int C) {
  return (size_t)stat_slices(rows) * (size_t)C * 5 * sizeof(double) + (size_t)C * 5 * sizeof(float);
}
extern "C" int b200np_bn_act_bwd2(const float* dy, const float* y, const float* x, const float* mean, const float* rstd,
                                  const float* scale, float plus_one, const float* vx, const float* vs, const float* vt,
                                  float* gx, float* gdy, float* gscale, long long rows, int C, int relu, void* ws,
                                  size_t ws_bytes, void* stream) {
  if (!dy || !y || !x || !mean || !rstd || rows <= 0 || C <= 0 || !ws) return B200NP_E_BADARG;
  if (ws_bytes < b200np_bn_workspace2(rows, C)) return B200NP_E_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int slices = stat_slices(rows);
  const long long per = ceil_div(rows, slices);
  double* part = (double*)ws;
  float* stats = (float*)(part + (size_t)slices * C * 5);
  bn_act_bwd2_l1_kernel<<<dim3((C + 31) / 32, slices), 256, 0, st>>>(dy, y, x, mean, rstd, vx, part, rows, C, per, relu);
  bn_act_bwd2_l2_kernel<<<(C + 127) / 128, 128, 0, st>>>(part, slices, C, rstd, (double)rows, stats, gscale);
  int launched = 2;
  if (gx || gdy) {
    const long long total = rows * C;
    bn_act_bwd2_apply_kernel<<<ew_grid(total, 256), 256, 0, st>>>(dy, y, x, mean, rstd, scale, plus_one, vx, vs, vt, stats,
                                                                  gx, gdy, total, C, relu);
    ++launched;
  }
  return launch_status(launched);
}

// ---------------------------------------------------------------------------------------------------------------
// Bayes-by-backprop layers of the "MR" variants (networks/bbb/BBBConv.py:83-105, BBBLinear.py):
//   sigma = log1p(exp(rho)),  w = mu + eps * sigma                         (eps: the host draws it, :86 / :92)
//   kl = 0.5 * sum( 2 log(sigma / ps) - 1 + (ps / sigma)^2 + ((mu - pm) / sigma)^2 )      (calculate_kl, :32-34, with
//        the argument order kl_loss uses at :102-105: q = prior, p = posterior)
// One pass writes w and sigma and the per-block KL partial sums (deterministic: the caller sums them in order with
// b200np_reduce); the backward turns dL/dw and the scalar dL/dkl into dL/dmu and dL/drho.
// ---------------------------------------------------------------------------------------------------------------
namespace b200np {
namespace {
__global__ void __launch_bounds__(256) bbb_sample_kl_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ rho,
                                                                const float* __restrict__ eps, float pm, float ps,
                                                                float* __restrict__ w, float* __restrict__ sigma,
                                                                float* __restrict__ kl_part, long long n) {
  __shared__ float red[32];
  const long long st = (long long)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += st) {
    const float r = __ldg(rho + i), m = __ldg(mu + i);
    const float s = log1pf(expf(r));
    sigma[i] = s;
    w[i] = fmaf(__ldg(eps + i), s, m);
    const float a = ps / s, b = (m - pm) / s;
    acc += 0.5f * (2.f * logf(s / ps) - 1.f + a * a + b * b);
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) kl_part[blockIdx.x] = tot;
}
__global__ void bbb_sample_kl_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ dkl,
                                         const float* __restrict__ mu, const float* __restrict__ rho,
                                         const float* __restrict__ eps, const float* __restrict__ sigma, float pm, float ps,
                                         float* __restrict__ dmu, float* __restrict__ drho, long long n) {
  const long long st = (long long)gridDim.x * blockDim.x;
  const float gk = dkl ? __ldg(dkl) : 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += st) {
    const float s = __ldg(sigma + i), m = __ldg(mu + i), g = dw ? __ldg(dw + i) : 0.f;
    const float d = m - pm, s2 = s * s;
    dmu[i] = g + gk * d / s2;
    const float dsig = g * __ldg(eps + i) + gk * (1.f / s - (ps * ps + d * d) / (s2 * s));
    drho[i] = dsig / (1.f + expf(-__ldg(rho + i)));     // d softplus / d rho = sigmoid(rho)
  }
}
}  // namespace
}  // namespace b200np

extern "C" int b200np_bbb_kl_blocks(long long n) { return ew_grid(n, 256); }
extern "C" int b200np_bbb_sample_kl_fwd(const float* mu, const float* rho, const float* eps, float prior_mu,
                                        float prior_sigma, float* w, float* sigma, float* kl_part, long long n,
                                        void* stream) {
  if (!mu || !rho || !eps || !w || !sigma || !kl_part || n <= 0 || !(prior_sigma > 0.f)) return B200NP_E_BADARG;
  bbb_sample_kl_fwd_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(mu, rho, eps, prior_mu, prior_sigma, w, sigma,
                                                                          kl_part, n);
  return launch_status();
}
extern "C" int b200np_bbb_sample_kl_bwd(const float* dw, const float* dkl, const float* mu, const float* rho,
                                        const float* eps, const float* sigma, float prior_mu, float prior_sigma,
                                        float* dmu, float* drho, long long n, void* stream) {
  if (!mu || !rho || !eps || !sigma || !dmu || !drho || n <= 0) return B200NP_E_BADARG;
  bbb_sample_kl_bwd_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(dw, dkl, mu, rho, eps, sigma, prior_mu,
                                                                          prior_sigma, dmu, drho, n);
  return launch_status();
}
