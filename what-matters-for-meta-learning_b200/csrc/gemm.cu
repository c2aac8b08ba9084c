// Strided, grouped fp32 GEMM with a fused epilogue (bias, activation, beta*C, row_scale*addend).
// One kernel serves every dense layer of the path in all three autograd roles:
//   forward   Y = act(X W^T + b)        A = X  (k contiguous), B = W   (k contiguous)
//   dgrad     dX = dZ W                 A = dZ (k contiguous), B = W   (n contiguous)
//   wgrad     dW = dZ^T X               A = dZ (m contiguous), B = X   (n contiguous)
// and the FAVOR+ feature projections.  CUDA-core fp32 (exact); 128x64x16 tiles, 8x8 per thread.
#include "common.cuh"

using namespace b200np;

namespace b200np {
int launch_gemm_umma(const b200np_gemm_desc& d, int a_vec, int b_vec, cudaStream_t st);  // gemm_umma.cu
int gemm_umma_splits(const b200np_gemm_desc& d);
}

namespace {

constexpr int BM = 128, BN = 64, BK = 16, APAD = 4;

struct GemmArgs {
  b200np_gemm_desc d;
  int a_vec, b_vec;  // float4 loads legal (pointer and stride alignment) for A / B
};

__global__ void __launch_bounds__(128) gemm_kernel(const GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM + APAD];
  __shared__ __align__(16) float Bs[BK][BN];
  const b200np_gemm_desc& d = g.d;
  const int grp = blockIdx.z;
  const float* __restrict__ A = d.A[grp];
  const float* __restrict__ B = d.B[grp];
  float* __restrict__ C = d.C[grp];
  const float* __restrict__ bias = d.bias[grp];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int M = d.M, N = d.N, K = d.K;
  const bool a_kc = d.a_cs == 1;  // A is k-contiguous
  const bool b_kc = d.b_rs == 1;  // B is k-contiguous

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ar[16], br[8];
  auto fetch = [&](int k0) {
    if (a_kc) {  // thread = row m0+tid, 16 consecutive k
      const int m = m0 + tid;
      const float* p = A + (long long)m * d.a_rs + k0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (m < M && g.a_vec && k0 + q * 4 + 3 < K) {
          float4 v = ldg4(p + q * 4);
          ar[q * 4] = v.x; ar[q * 4 + 1] = v.y; ar[q * 4 + 2] = v.z; ar[q * 4 + 3] = v.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) ar[q * 4 + e] = (m < M && k0 + q * 4 + e < K) ? __ldg(p + q * 4 + e) : 0.f;
        }
      }
    } else {  // A m-contiguous: thread = (k = tid/8, 16 consecutive m starting at (tid%8)*16)
      const int k = k0 + (tid >> 3), mb = m0 + (tid & 7) * 16;
      const float* p = A + (long long)k * d.a_cs + mb;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (k < K && g.a_vec && mb + q * 4 + 3 < M) {
          float4 v = ldg4(p + q * 4);
          ar[q * 4] = v.x; ar[q * 4 + 1] = v.y; ar[q * 4 + 2] = v.z; ar[q * 4 + 3] = v.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) ar[q * 4 + e] = (k < K && mb + q * 4 + e < M) ? __ldg(p + q * 4 + e) : 0.f;
        }
      }
    }
    if (b_kc) {  // thread = (n = tid/2, 8 consecutive k at (tid%2)*8)
      const int n = n0 + (tid >> 1), kb = k0 + (tid & 1) * 8;
      const float* p = B + (long long)n * d.b_cs + kb;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (n < N && g.b_vec && kb + q * 4 + 3 < K) {
          float4 v = ldg4(p + q * 4);
          br[q * 4] = v.x; br[q * 4 + 1] = v.y; br[q * 4 + 2] = v.z; br[q * 4 + 3] = v.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) br[q * 4 + e] = (n < N && kb + q * 4 + e < K) ? __ldg(p + q * 4 + e) : 0.f;
        }
      }
    } else {  // B n-contiguous: thread = (k = tid/8, 8 consecutive n at (tid%8)*8)
      const int k = k0 + (tid >> 3), nb = n0 + (tid & 7) * 8;
      const float* p = B + (long long)k * d.b_rs + nb;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (k < K && g.b_vec && nb + q * 4 + 3 < N) {
          float4 v = ldg4(p + q * 4);
          br[q * 4] = v.x; br[q * 4 + 1] = v.y; br[q * 4 + 2] = v.z; br[q * 4 + 3] = v.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) br[q * 4 + e] = (k < K && nb + q * 4 + e < N) ? __ldg(p + q * 4 + e) : 0.f;
        }
      }
    }
  };

  const int tm = tid >> 3, tn = tid & 7;
  if (K > 0) fetch(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    __syncthreads();
    if (a_kc) {
#pragma unroll
      for (int e = 0; e < 16; ++e) As[e][tid] = ar[e];
    } else {
      const int k = tid >> 3, mb = (tid & 7) * 16;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(&As[k][mb + q * 4]) = make_float4(ar[q * 4], ar[q * 4 + 1], ar[q * 4 + 2], ar[q * 4 + 3]);
    }
    if (b_kc) {
      const int n = tid >> 1, kb = (tid & 1) * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) Bs[kb + e][n] = br[e];
    } else {
      const int k = tid >> 3, nb = (tid & 7) * 8;
      *reinterpret_cast<float4*>(&Bs[k][nb]) = make_float4(br[0], br[1], br[2], br[3]);
      *reinterpret_cast<float4*>(&Bs[k][nb + 4]) = make_float4(br[4], br[5], br[6], br[7]);
    }
    __syncthreads();
    if (k0 + BK < K) fetch(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][tm * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][tm * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tn * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][tn * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + tm * 8 + i;
    if (m >= M) break;
    const float rs = (d.row_scale && grp == 0) ? __ldg(d.row_scale + m) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tn * 8 + j;
      if (n >= N) break;
      float v = d.alpha * acc[i][j];
      if (bias) v += __ldg(bias + n);
      float* cp = C + (long long)m * d.ldc + n;
      if (d.beta != 0.f) v = fmaf(d.beta, *cp, v);
      if (d.row_scale && grp == 0) v = fmaf(rs, __ldg(d.addend + (long long)m * d.ld_add + n), v);
      if (d.act == B200NP_ACT_RELU) v = fmaxf(v, 0.f);
      else if (d.act == B200NP_ACT_TANH) v = tanhf(v);
      *cp = v;
    }
  }
}

// Few outputs, long reduction (the 256 -> 2 output layer and its weight gradient, the label embedding's weight gradient):
// the tiled kernel would run a handful of CTAs through K/16 dependent iterations (~20 us).  One warp per output element,
// lanes stride over k, shuffle reduction; same epilogue.  Deterministic (fixed lane assignment and reduction tree).
__global__ void __launch_bounds__(256) gemm_skinny_kernel(const GemmArgs g) {
  const b200np_gemm_desc& d = g.d;
  const int grp = blockIdx.y;
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (o >= d.M * d.N) return;
  const int m = o / d.N, n = o - m * d.N;
  const float* __restrict__ a = d.A[grp] + (long long)m * d.a_rs;
  const float* __restrict__ b = d.B[grp] + (long long)n * d.b_cs;
  float s0 = 0.f, s1 = 0.f;
  int k = lane;
  for (; k + 32 < d.K; k += 64) {
    s0 = fmaf(__ldg(a + (long long)k * d.a_cs), __ldg(b + (long long)k * d.b_rs), s0);
    s1 = fmaf(__ldg(a + (long long)(k + 32) * d.a_cs), __ldg(b + (long long)(k + 32) * d.b_rs), s1);
  }
  if (k < d.K) s0 = fmaf(__ldg(a + (long long)k * d.a_cs), __ldg(b + (long long)k * d.b_rs), s0);
  const float acc = warp_sum(s0 + s1);
  if (lane) return;
  float v = d.alpha * acc;
  if (d.bias[grp]) v += __ldg(d.bias[grp] + n);
  float* cp = d.C[grp] + (long long)m * d.ldc + n;
  if (d.beta != 0.f) v = fmaf(d.beta, *cp, v);
  if (d.row_scale && grp == 0) v = fmaf(__ldg(d.row_scale + m), __ldg(d.addend + (long long)m * d.ld_add + n), v);
  if (d.act == B200NP_ACT_RELU) v = fmaxf(v, 0.f);
  else if (d.act == B200NP_ACT_TANH) v = tanhf(v);
  *cp = v;
}

}  // namespace

extern "C" size_t b200np_gemm_workspace(const b200np_gemm_desc* dp) {
  if (!dp || dp->M <= 0 || dp->N <= 0) return 0;
  const int s = gemm_umma_splits(*dp);
  return s > 1 ? (size_t)s * dp->M * dp->N * sizeof(float) : 0;
}

extern "C" int b200np_gemm(const b200np_gemm_desc* dp, void* stream) {
  if (!dp) return B200NP_E_BADARG;
  const b200np_gemm_desc& d = *dp;
  if (d.groups < 1 || d.groups > 8 || d.M < 0 || d.N < 0 || d.K < 0) return B200NP_E_BADARG;
  if (d.M == 0 || d.N == 0) return B200NP_OK;
  if (!((d.a_rs == 1) ^ (d.a_cs == 1)) && !(d.a_rs == 1 && d.a_cs == 1)) return B200NP_E_BADARG;
  if (!((d.b_rs == 1) ^ (d.b_cs == 1)) && !(d.b_rs == 1 && d.b_cs == 1)) return B200NP_E_BADARG;
  if (d.row_scale && !d.addend) return B200NP_E_BADARG;
  GemmArgs g;
  g.d = d;
  g.a_vec = 1;
  g.b_vec = 1;
  const long long a_ld = d.a_cs == 1 ? d.a_rs : d.a_cs;
  const long long b_ld = d.b_rs == 1 ? d.b_cs : d.b_rs;
  if (a_ld % 4) g.a_vec = 0;
  if (b_ld % 4) g.b_vec = 0;
  for (int i = 0; i < d.groups; ++i) {
    if (!d.A[i] || !d.B[i] || !d.C[i]) return B200NP_E_BADARG;
    if (!aligned16(d.A[i])) g.a_vec = 0;
    if (!aligned16(d.B[i])) g.b_vec = 0;
  }
  for (int i = d.groups; i < 8; ++i) { g.d.A[i] = nullptr; g.d.B[i] = nullptr; g.d.C[i] = nullptr; g.d.bias[i] = nullptr; }
  {  // tensor-core path for everything that fills a tile; CUDA cores for the tiny layers and fp32 mode
    int rc = launch_gemm_umma(g.d, g.a_vec, g.b_vec, as_stream(stream));
    if (rc != B200NP_E_UNSUPPORTED) return rc;
  }
  if (d.conv_operand) return B200NP_E_UNSUPPORTED;   // the implicit convolution exists on the tensor-core path only
  if (d.sum_groups) {  // CUDA-core path: one accumulating launch per K-slice, epilogue on the last
    for (int i = 0; i < d.groups; ++i) {
      GemmArgs gi = g;
      const bool last = i == d.groups - 1;
      gi.d.groups = 1;
      gi.d.sum_groups = 0;
      gi.d.A[0] = d.A[i]; gi.d.B[0] = d.B[i]; gi.d.C[0] = d.C[0];
      gi.d.bias[0] = last ? d.bias[0] : nullptr;
      gi.d.act = last ? d.act : B200NP_ACT_NONE;
      gi.d.row_scale = last ? d.row_scale : nullptr;
      gi.d.beta = i == 0 ? d.beta : 1.f;
      dim3 grid((d.M + BM - 1) / BM, (d.N + BN - 1) / BN, 1);
      gemm_kernel<<<grid, 128, 0, as_stream(stream)>>>(gi);
      int rc = launch_status();
      if (rc != B200NP_OK) return rc;
    }
    return B200NP_OK;
  }
  if ((long long)d.M * d.N <= 4096 && d.K >= 64) {
    gemm_skinny_kernel<<<dim3((unsigned)((d.M * d.N + 7) / 8), d.groups), 256, 0, as_stream(stream)>>>(g);
    return launch_status();
  }
  dim3 grid((d.M + BM - 1) / BM, (d.N + BN - 1) / BN, d.groups);
  gemm_kernel<<<grid, 128, 0, as_stream(stream)>>>(g);
  return launch_status();
}
