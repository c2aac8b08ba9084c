// Pooling / flatten kernels at the tail of the CNNs.  All tensors here are tiny next to the conv
// activations (<= N x 4096 floats); they are written for exactness (first-max tie rule) and
// coalesced access on the NHWC side.
#include "common.cuh"

using namespace b200np;

// AdaptiveMaxPool2d((2,2)) + NCHW flatten.  Thread = (image, channel): scans the four windows in
// torch's order (rows, then columns; strict '>' keeps the first maximum), writes one float4.
__global__ void amp2_flat_fwd_kernel(const float* __restrict__ x, float* __restrict__ out,
                                     int32_t* __restrict__ idx, int N, int H, int W, int C) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * C) return;
  int n = (int)(i / C), c = (int)(i - (long long)n * C);
  const float* p = x + (long long)n * H * W * C + c;
  int hh = H / 2, hw = W / 2;
  float mo[4];
  int ao[4];
  if (H == 4 && W == 4) {
    // the trunk's last map (the only shape on the hot path): all 16 loads in flight before the first compare
    float v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = __ldg(p + (long long)q * C);
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int base = (o >> 1) * 8 + (o & 1) * 2;         // window origin (row 2*oh, column 2*ow)
      float m = v[base];
      int a = base;
#pragma unroll
      for (int k = 1; k < 4; ++k) {                        // torch's scan order: rows, then columns; first maximum wins
        const int q = base + (k >> 1) * 4 + (k & 1);
        if (v[q] > m) { m = v[q]; a = q; }
      }
      mo[o] = m;
      ao[o] = a;
    }
  } else {
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      int oh = o >> 1, ow = o & 1;
      float m = -INFINITY;
      int a = (oh * hh) * W + ow * hw;
      for (int h = oh * hh; h < (oh + 1) * hh; ++h)
        for (int w = ow * hw; w < (ow + 1) * hw; ++w) {
          float v = __ldg(p + ((long long)h * W + w) * C);
          if (v > m) { m = v; a = h * W + w; }
        }
      mo[o] = m;
      ao[o] = a;
    }
  }
  // the four window results of a channel are consecutive in the NCHW-flatten order: one 16-byte store each
  const long long k = (long long)n * C * 4 + c * 4;
  if (aligned16(out) && aligned16(idx)) {
    *reinterpret_cast<float4*>(out + k) = make_float4(mo[0], mo[1], mo[2], mo[3]);
    *reinterpret_cast<int4*>(idx + k) = make_int4(ao[0], ao[1], ao[2], ao[3]);
  } else {
#pragma unroll
    for (int o = 0; o < 4; ++o) { out[k + o] = mo[o]; idx[k + o] = ao[o]; }
  }
}
__global__ void amp2_flat_bwd_kernel(const float* __restrict__ dout, const int32_t* __restrict__ idx,
                                     const float* __restrict__ xs, float* __restrict__ dx, int N, int H, int W,
                                     int C) {
  long long n_el = (long long)N * H * W * C;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, st = (long long)gridDim.x * blockDim.x;
  for (; i < n_el; i += st) {
    int c = (int)(i % C);
    long long r = i / C;
    int w = (int)(r % W);
    r /= W;
    int h = (int)(r % H);
    long long n = r / H;
    int o = (h / (H / 2)) * 2 + (w / (W / 2));
    long long k = n * C * 4 + c * 4 + o;
    float g = (idx[k] == h * W + w) ? dout[k] : 0.f;
    dx[i] = (xs[i] > 0.f) ? g : 0.f;
  }
}
extern "C" int b200np_adaptive_maxpool2x2_flatten_fwd(const float* x, float* out, int32_t* idx, int N, int H,
                                                      int W, int C, void* stream) {
  if (!x || !out || !idx || N <= 0 || C <= 0 || H < 2 || W < 2 || (H & 1) || (W & 1)) return B200NP_E_BADARG;
  long long n = (long long)N * C;
  amp2_flat_fwd_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, as_stream(stream)>>>(x, out, idx, N, H, W, C);
  return launch_status();
}
extern "C" int b200np_adaptive_maxpool2x2_flatten_bwd(const float* dout, const int32_t* idx, const float* x_saved,
                                                      float* dx, int N, int H, int W, int C, void* stream) {
  if (!dout || !idx || !x_saved || !dx || N <= 0 || C <= 0 || H < 2 || W < 2 || (H & 1) || (W & 1))
    return B200NP_E_BADARG;
  long long n = (long long)N * H * W * C;
  amp2_flat_bwd_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(dout, idx, x_saved, dx, N, H, W, C);
  return launch_status();
}

// NHWC <-> NCHW-order flatten.
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ out, int N, int H, int W,
                                    int C) {
  long long n_el = (long long)N * H * W * C;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, st = (long long)gridDim.x * blockDim.x;
  for (; i < n_el; i += st) {  // i indexes the NHWC side (coalesced reads)
    int c = (int)(i % C);
    long long r = i / C;
    int hw = (int)(r % ((long long)H * W));
    long long n = r / ((long long)H * W);
    out[(n * C + c) * H * W + hw] = x[i];
  }
}
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ dout, const float* __restrict__ xs,
                                    float* __restrict__ dx, int N, int H, int W, int C) {
  long long n_el = (long long)N * H * W * C;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, st = (long long)gridDim.x * blockDim.x;
  for (; i < n_el; i += st) {
    int c = (int)(i % C);
    long long r = i / C;
    int hw = (int)(r % ((long long)H * W));
    long long n = r / ((long long)H * W);
    float g = dout[(n * C + c) * H * W + hw];
    dx[i] = (!xs || xs[i] > 0.f) ? g : 0.f;
  }
}
extern "C" int b200np_nhwc_to_nchw_flat(const float* x, float* out, int N, int H, int W, int C, void* stream) {
  if (!x || !out || N <= 0 || H <= 0 || W <= 0 || C <= 0) return B200NP_E_BADARG;
  long long n = (long long)N * H * W * C;
  nhwc_to_nchw_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(x, out, N, H, W, C);
  return launch_status();
}
extern "C" int b200np_nchw_flat_to_nhwc(const float* dout, const float* x_saved, float* dx, int N, int H, int W,
                                        int C, void* stream) {
  if (!dout || !dx || N <= 0 || H <= 0 || W <= 0 || C <= 0) return B200NP_E_BADARG;
  long long n = (long long)N * H * W * C;
  nchw_to_nhwc_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(dout, x_saved, dx, N, H, W, C);
  return launch_status();
}

// MaxPool2d(2,2) on NHWC; first maximum in (0,0),(0,1),(1,0),(1,1) order.
__global__ void maxpool2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int8_t* __restrict__ idx,
                                    int N, int H, int W, int C) {
  int OH = H / 2, OW = W / 2;
  long long n_el = (long long)N * OH * OW * C;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, st = (long long)gridDim.x * blockDim.x;
  for (; i < n_el; i += st) {
    int c = (int)(i % C);
    long long r = i / C;
    int ow = (int)(r % OW);
    r /= OW;
    int oh = (int)(r % OH);
    long long n = r / OH;
    const float* p = x + ((n * H + oh * 2) * W + ow * 2) * C + c;
    float m = p[0];
    int a = 0;
    float v = p[C];
    if (v > m) { m = v; a = 1; }
    v = p[(long long)W * C];
    if (v > m) { m = v; a = 2; }
    v = p[(long long)W * C + C];
    if (v > m) { m = v; a = 3; }
    y[i] = m;
    idx[i] = (int8_t)a;
  }
}
__global__ void maxpool2_bwd_kernel(const float* __restrict__ dy, const int8_t* __restrict__ idx,
                                    const float* __restrict__ xs, float* __restrict__ dx, int N, int H, int W,
                                    int C) {
  int OH = H / 2, OW = W / 2;
  long long n_el = (long long)N * H * W * C;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, st = (long long)gridDim.x * blockDim.x;
  for (; i < n_el; i += st) {
    int c = (int)(i % C);
    long long r = i / C;
    int w = (int)(r % W);
    r /= W;
    int h = (int)(r % H);
    long long n = r / H;
    long long k = ((n * OH + h / 2) * OW + w / 2) * C + c;
    float g = (idx[k] == (h & 1) * 2 + (w & 1)) ? dy[k] : 0.f;
    dx[i] = (xs[i] > 0.f) ? g : 0.f;
  }
}
// four channels per thread (C % 4 == 0, 16-byte aligned tensors): 16-byte loads / stores, one packed index word, and
// a quarter of the index arithmetic -- the scalar forms moved 1.6 TB/s
__global__ void maxpool2_fwd_vec4_kernel(const float* __restrict__ x, float* __restrict__ y, int8_t* __restrict__ idx,
                                         int N, int H, int W, int C) {
  const int OH = H / 2, OW = W / 2, C4 = C / 4;
  const long long n_el = (long long)N * OH * OW * C4, st = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += st) {
    const int c = (int)(i % C4) * 4;
    long long r = i / C4;
    const int ow = (int)(r % OW);
    r /= OW;
    const int oh = (int)(r % OH);
    const long long n = r / OH;
    const float* p = x + ((n * H + oh * 2) * W + ow * 2) * C + c;
    const float4 v0 = ldg4(p), v1 = ldg4(p + C), v2 = ldg4(p + (long long)W * C), v3 = ldg4(p + (long long)W * C + C);
    float m[4] = {v0.x, v0.y, v0.z, v0.w};
    int a[4] = {0, 0, 0, 0};
    const float c1[4] = {v1.x, v1.y, v1.z, v1.w}, c2[4] = {v2.x, v2.y, v2.z, v2.w}, c3[4] = {v3.x, v3.y, v3.z, v3.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (c1[e] > m[e]) { m[e] = c1[e]; a[e] = 1; }
      if (c2[e] > m[e]) { m[e] = c2[e]; a[e] = 2; }
      if (c3[e] > m[e]) { m[e] = c3[e]; a[e] = 3; }
    }
    *reinterpret_cast<float4*>(y + i * 4) = make_float4(m[0], m[1], m[2], m[3]);
    *reinterpret_cast<char4*>(idx + i * 4) = make_char4((signed char)a[0], (signed char)a[1], (signed char)a[2], (signed char)a[3]);
  }
}
__global__ void maxpool2_bwd_vec4_kernel(const float* __restrict__ dy, const int8_t* __restrict__ idx,
                                         const float* __restrict__ xs, float* __restrict__ dx, int N, int H, int W, int C) {
  const int OH = H / 2, OW = W / 2, C4 = C / 4;
  const long long n_el = (long long)N * H * W * C4, st = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += st) {
    const int c = (int)(i % C4) * 4;
    long long r = i / C4;
    const int w = (int)(r % W);
    r /= W;
    const int h = (int)(r % H);
    const long long n = r / H;
    const long long k = ((n * OH + h / 2) * OW + w / 2) * C + c;
    const int me = (h & 1) * 2 + (w & 1);
    const char4 a = *reinterpret_cast<const char4*>(idx + k);
    const float4 g = ldg4(dy + k), xv = ldg4(xs + i * 4);
    float4 o;
    o.x = (a.x == me && xv.x > 0.f) ? g.x : 0.f;
    o.y = (a.y == me && xv.y > 0.f) ? g.y : 0.f;
    o.z = (a.z == me && xv.z > 0.f) ? g.z : 0.f;
    o.w = (a.w == me && xv.w > 0.f) ? g.w : 0.f;
    *reinterpret_cast<float4*>(dx + i * 4) = o;
  }
}
extern "C" int b200np_maxpool2x2_fwd(const float* x, float* y, int8_t* idx, int N, int H, int W, int C,
                                     void* stream) {
  if (!x || !y || !idx || N <= 0 || C <= 0 || H < 2 || W < 2 || (H & 1) || (W & 1)) return B200NP_E_BADARG;
  long long n = (long long)N * (H / 2) * (W / 2) * C;
  if ((C & 3) == 0 && aligned16(x) && aligned16(y) && (reinterpret_cast<uintptr_t>(idx) & 3u) == 0)
    maxpool2_fwd_vec4_kernel<<<ew_grid(n / 4, 256), 256, 0, as_stream(stream)>>>(x, y, idx, N, H, W, C);
  else
    maxpool2_fwd_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(x, y, idx, N, H, W, C);
  return launch_status();
}
extern "C" int b200np_maxpool2x2_bwd(const float* dy, const int8_t* idx, const float* x_saved, float* dx, int N,
                                     int H, int W, int C, void* stream) {
  if (!dy || !idx || !x_saved || !dx || N <= 0 || C <= 0 || H < 2 || W < 2 || (H & 1) || (W & 1))
    return B200NP_E_BADARG;
  long long n = (long long)N * H * W * C;
  if ((C & 3) == 0 && aligned16(dy) && aligned16(x_saved) && aligned16(dx) && (reinterpret_cast<uintptr_t>(idx) & 3u) == 0)
    maxpool2_bwd_vec4_kernel<<<ew_grid(n / 4, 256), 256, 0, as_stream(stream)>>>(dy, idx, x_saved, dx, N, H, W, C);
  else
    maxpool2_bwd_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(dy, idx, x_saved, dx, N, H, W, C);
  return launch_status();
}
