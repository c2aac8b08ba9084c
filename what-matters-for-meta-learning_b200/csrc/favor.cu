// FAVOR+ attention core (networks/fast_attention.py:74-99,151-156; math in SURVEY.md appendix A).
//
// The feature pre-activations U = c*xq*P^T and W = c*xk*P^T are produced by the GEMM kernel.
// Everything after them is fused here, one CTA per (task, head):
//   forward : Q' = rho(exp(U - s - m) + eps), K' = rho(exp(W - t - g) + eps) built chunk-wise in
//             shared memory, A = Q'K'^T (nt x nc) accumulated in registers, out = A v / rowsum(A).
//             Re-associated form: the [M x d] "context" matrix of the reference never exists.
//   backward: dA, dv, then chunk-wise dU = (dA K') * Eq, dW = (dA^T Q') * Ek with the row sums
//             that feed the diag / max terms, and the row-argmax routing for queries.
// Layouts: xq/xk/v rows are (t, i, h); out / d_out are [T, nt, d, H] (index e*H + h).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

using namespace b200np;

namespace {

constexpr int FC = 128;       // features per chunk
constexpr int PITCH = FC + 1; // odd pitch: conflict-free row-strided reads
constexpr float kEps = 1e-4f; // fast_attention.py:74

// Feature split over a thread-block cluster: the chunk loops below are latency-bound (load chunk -> exp -> barrier ->
// small products -> barrier, 12 chunks for M = 1419), so the S CTAs of a cluster take M/S features of ONE (task, head)
// each and combine their partial tiles / row sums through distributed shared memory, in rank order (deterministic).
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the float at the same shared-memory offset as `p` in CTA `rank` of this cluster
__device__ __forceinline__ float ld_cluster(const float* p, uint32_t rank) {
  uint32_t la = static_cast<uint32_t>(__cvta_generic_to_shared(p)), ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}

// one warp per row
__global__ void favor_rowstats_kernel(const float* __restrict__ x, const float* __restrict__ U,
                                      float* __restrict__ diag, float* __restrict__ rowmax,
                                      int32_t* __restrict__ argmax, long long R, int d, int M, long long ldu) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  float ss = 0.f;
  for (int k = lane; k < d; k += 32) {
    float v = x[r * d + k];
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  float m = -INFINITY;
  int a = 0x7fffffff;
  for (int f = lane; f < M; f += 32) {
    float v = U[r * ldu + f];
    if (v > m) { m = v; a = f; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float om = __shfl_xor_sync(0xffffffffu, m, o);
    int oa = __shfl_xor_sync(0xffffffffu, a, o);
    if (om > m || (om == m && oa < a)) { m = om; a = oa; }
  }
  if (lane == 0) {
    const float dn = powf((float)d, -0.25f);
    diag[r] = ss / 2.0f * (dn * dn);  // fast_attention.py:86-89
    rowmax[r] = m;
    if (argmax) argmax[r] = a;
  }
}

__global__ void reduce_kernel(const float* __restrict__ x, long long n, float* __restrict__ out, int op) {
  __shared__ float red[32];
  float m = op == 0 ? -INFINITY : 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) m = op == 0 ? fmaxf(m, x[i]) : m + x[i];
  m = op == 0 ? warp_max(m) : warp_sum(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : (op == 0 ? -INFINITY : 0.f);
    t = op == 0 ? warp_max(t) : warp_sum(t);
    if (threadIdx.x == 0) out[0] = t;
  }
}

__global__ void __launch_bounds__(256) favor_attn_fwd_kernel(
    const float* __restrict__ U, const float* __restrict__ W, const float* __restrict__ sq,
    const float* __restrict__ mq, const float* __restrict__ tk, const float* __restrict__ g,
    const float* __restrict__ v, float* __restrict__ out, float* __restrict__ A, float* __restrict__ Dn,
    float* __restrict__ ties, int H, int nt, int nc, int d, int M, long long ldu, int S) {
  extern __shared__ float sm[];
  float* Qs = sm;                    // [nt][PITCH]
  float* Ks = Qs + nt * PITCH;       // [nc][PITCH]
  float* As = Ks + nc * PITCH;       // [nt*nc]
  float* Ds = As + nt * nc;          // [nt]
  float* s1 = Ds + nt;               // [nt] sq
  float* s2 = s1 + nt;               // [nt] mq
  float* s3 = s2 + nt;               // [nc] tk
  const int tid = threadIdx.x;
  const uint32_t rank = S > 1 ? cluster_rank() : 0u;
  const int pair = blockIdx.x / S;
  const int t = pair / H, h = pair % H;
  const int nchunk = (M + FC - 1) / FC, per = (nchunk + S - 1) / S;
  const int f_lo = (int)rank * per * FC, f_hi = min(M, ((int)rank + 1) * per * FC);   // this CTA's features
  const float gmax = __ldg(g);
  const float rho = rsqrtf((float)M);
  for (int i = tid; i < nt; i += 256) {
    long long r = ((long long)t * nt + i) * H + h;
    s1[i] = sq[r];
    s2[i] = mq[r];
  }
  for (int j = tid; j < nc; j += 256) s3[j] = tk[((long long)t * nc + j) * H + h];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int tie_local = 0;
  __syncthreads();
  for (int f0 = f_lo; f0 < f_hi; f0 += FC) {
    const int fc = f_hi - f0 < FC ? f_hi - f0 : FC;
    // the chunk's U / W values are fetched eight at a time before the exponentials: one exposed load latency per
    // batch instead of one per element (this loop was the kernel's critical path)
    for (int base = tid; base < (nt + nc) * FC; base += 256 * 8) {
      float raw[8];
#pragma unroll
      for (int u8 = 0; u8 < 8; ++u8) {
        const int idx = base + u8 * 256;
        const int row = idx / FC, f = idx - row * FC;
        raw[u8] = 0.f;
        if (idx < (nt + nc) * FC && f < fc) {
          const long long r = row < nt ? ((long long)t * nt + row) * H + h : ((long long)t * nc + (row - nt)) * H + h;
          raw[u8] = (row < nt ? U : W)[r * ldu + f0 + f];
        }
      }
#pragma unroll
      for (int u8 = 0; u8 < 8; ++u8) {
        const int idx = base + u8 * 256;
        if (idx >= (nt + nc) * FC) break;
        const int row = idx / FC, f = idx - row * FC;
        float val = 0.f;
        if (f < fc) {
          if (row < nt) {
            val = rho * (expf(raw[u8] - s1[row] - s2[row]) + kEps);
          } else {
            val = rho * (expf(raw[u8] - s3[row - nt] - gmax) + kEps);
            tie_local += (raw[u8] == gmax);
          }
        }
        if (row < nt) Qs[row * PITCH + f] = val;
        else Ks[(row - nt) * PITCH + f] = val;
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int p = tid + q * 256;
      if (p < nt * nc) {
        const int i = p / nc, j = p - i * nc;
        const float* qp = Qs + i * PITCH;
        const float* kp = Ks + j * PITCH;
        float a = acc[q];
#pragma unroll 8
        for (int f = 0; f < FC; ++f) a = fmaf(qp[f], kp[f], a);
        acc[q] = a;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int p = tid + q * 256;
    if (p < nt * nc) As[p] = acc[q];
  }
  // tie count: integer-valued float adds are exact -> deterministic
  {
    int tl = tie_local;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tl += __shfl_xor_sync(0xffffffffu, tl, o);
    if ((tid & 31) == 0 && tl) atomicAdd(ties, (float)tl);
  }
  if (S > 1) {
    cluster_sync();                    // every rank's partial tile is in its shared memory
    float full[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {      // every rank forms the whole tile, summing the partials in rank order
      const int p = tid + q * 256;
      full[q] = 0.f;
      if (p < nt * nc)
        for (int r = 0; r < S; ++r) full[q] += ld_cluster(As + p, (uint32_t)r);
    }
    cluster_sync();                    // all remote reads are done: As may be overwritten (and CTAs may exit)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int p = tid + q * 256;
      if (p < nt * nc) As[p] = full[q];
    }
  }
  __syncthreads();
  for (int i = tid; i < nt; i += 256) {
    float s = 0.f;
    for (int j = 0; j < nc; ++j) s += As[i * nc + j];
    Ds[i] = s;
    if (rank == 0) Dn[((long long)t * H + h) * nt + i] = s;
  }
  if (rank == 0)
    for (int p = tid; p < nt * nc; p += 256) A[((long long)t * H + h) * nt * nc + p] = As[p];
  __syncthreads();
  // out rows i = rank, rank + S, ...: the ranks share the A v product
  const int ni = (nt - (int)rank + S - 1) / S;
  for (int idx = tid; idx < ni * d; idx += 256) {
    const int ii = idx / d, e = idx - ii * d, i = (int)rank + ii * S;
    float o = 0.f;
    for (int j = 0; j < nc; ++j) o = fmaf(As[i * nc + j], v[(((long long)t * nc + j) * H + h) * d + e], o);
    out[(((long long)t * nt + i) * d + e) * H + h] = o / Ds[i];
  }
}

__global__ void __launch_bounds__(256) favor_attn_bwd_kernel(
    const float* __restrict__ d_out, const float* __restrict__ U, const float* __restrict__ W,
    const float* __restrict__ sq, const float* __restrict__ mq, const int32_t* __restrict__ amq,
    const float* __restrict__ tk, const float* __restrict__ g, const float* __restrict__ v,
    const float* __restrict__ out, const float* __restrict__ A, const float* __restrict__ Dn,
    float* __restrict__ dU, float* __restrict__ dW, float* __restrict__ dv, float* __restrict__ ds_c2,
    float* __restrict__ dt_c2, float* __restrict__ dg_part, int H, int nt, int nc, int d, int M, long long ldu, int S) {
  extern __shared__ float sm[];
  const int dp = d + 1;
  float* Eq = sm;                     // [nt][PITCH]
  float* Ek = Eq + nt * PITCH;        // [nc][PITCH]
  float* dOs = Ek + nc * PITCH;       // [nt][d+1]
  float* vs = dOs + nt * dp;          // [nc][d+1]
  float* dAs = vs + nc * dp;          // [nt*nc]
  float* An = dAs + nt * nc;          // [nt*nc]  A / D
  float* rv = An + nt * nc;           // [nt] dO . out
  float* Ds = rv + nt;                // [nt]
  float* s1 = Ds + nt;                // [nt]
  float* s2 = s1 + nt;                // [nt]
  float* s3 = s2 + nt;                // [nc]
  float* rq = s3 + nc;                // [nt][4] row-sum partials of dU
  float* rk = rq + nt * 4;            // [nc][4] row-sum partials of dW
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // S CTAs (one cluster) per (task, head): every rank builds dA (cheap, needed by all), takes 1/S of the feature chunks
  // and of the dv rows; rank 0 folds the row-sum partials of all ranks (in rank order) into ds / dt / dg
  const uint32_t rank = S > 1 ? cluster_rank() : 0u;
  const int pair = blockIdx.x / S;
  const int t = pair / H, h = pair % H;
  const int nchunk = (M + FC - 1) / FC, per = (nchunk + S - 1) / S;
  const int f_lo = (int)rank * per * FC, f_hi = min(M, ((int)rank + 1) * per * FC);
  const float gmax = __ldg(g);
  const float rho = rsqrtf((float)M);
  const float rho_eps = rho * kEps;
  const float dn = powf((float)d, -0.25f);
  const float c2 = dn * dn;

  for (int idx = tid; idx < nt * d; idx += 256) {
    const int i = idx / d, e = idx - i * d;
    dOs[i * dp + e] = d_out[(((long long)t * nt + i) * d + e) * H + h];
  }
  for (int idx = tid; idx < nc * d; idx += 256) {
    const int j = idx / d, e = idx - j * d;
    vs[j * dp + e] = v[(((long long)t * nc + j) * H + h) * d + e];
  }
  for (int i = tid; i < nt; i += 256) {
    const long long r = ((long long)t * nt + i) * H + h;
    s1[i] = sq[r];
    s2[i] = mq[r];
    Ds[i] = Dn[((long long)t * H + h) * nt + i];
  }
  for (int j = tid; j < nc; j += 256) s3[j] = tk[((long long)t * nc + j) * H + h];
  for (int k = tid; k < (nt + nc) * 4; k += 256) rq[k] = 0.f;
  __syncthreads();
  // r_i = dO_i . out_i  (one warp per row)
  for (int i = warp; i < nt; i += 8) {
    float s = 0.f;
    for (int e = lane; e < d; e += 32)
      s = fmaf(dOs[i * dp + e], out[(((long long)t * nt + i) * d + e) * H + h], s);
    s = warp_sum(s);
    if (lane == 0) rv[i] = s;
  }
  __syncthreads();
  for (int p = tid; p < nt * nc; p += 256) {
    const int i = p / nc, j = p - i * nc;
    float gij = 0.f;
    for (int e = 0; e < d; ++e) gij = fmaf(dOs[i * dp + e], vs[j * dp + e], gij);
    dAs[p] = (gij - rv[i]) / Ds[i];
    An[p] = A[((long long)t * H + h) * nt * nc + p] / Ds[i];
  }
  __syncthreads();
  {
    const int nj = (nc - (int)rank + S - 1) / S;          // dv rows j = rank, rank + S, ...
    for (int idx = tid; idx < nj * d; idx += 256) {
      const int jj = idx / d, e = idx - jj * d, j = (int)rank + jj * S;
      float s = 0.f;
      for (int i = 0; i < nt; ++i) s = fmaf(An[i * nc + j], dOs[i * dp + e], s);
      dv[(((long long)t * nc + j) * H + h) * d + e] = s;
    }
  }

  // chunk loop: thread = (feature f = tid%128, row parity tid/128); each warp owns one quarter
  // of the chunk's features for the rows it visits -> deterministic row-sum partials.
  const int f = tid & (FC - 1), half = tid >> 7, wq = (tid >> 5) & 3;
  for (int f0 = f_lo; f0 < f_hi; f0 += FC) {
    const int fc = f_hi - f0 < FC ? f_hi - f0 : FC;
    __syncthreads();
    for (int base = tid; base < (nt + nc) * FC; base += 256 * 8) {   // loads batched as in the forward kernel
      float raw[8];
#pragma unroll
      for (int u8 = 0; u8 < 8; ++u8) {
        const int idx = base + u8 * 256;
        const int row = idx / FC, ff = idx - row * FC;
        raw[u8] = 0.f;
        if (idx < (nt + nc) * FC && ff < fc) {
          const long long r = row < nt ? ((long long)t * nt + row) * H + h : ((long long)t * nc + (row - nt)) * H + h;
          raw[u8] = (row < nt ? U : W)[r * ldu + f0 + ff];
        }
      }
#pragma unroll
      for (int u8 = 0; u8 < 8; ++u8) {
        const int idx = base + u8 * 256;
        if (idx >= (nt + nc) * FC) break;
        const int row = idx / FC, ff = idx - row * FC;
        float val = 0.f;
        if (ff < fc) val = row < nt ? rho * expf(raw[u8] - s1[row] - s2[row]) : rho * expf(raw[u8] - s3[row - nt] - gmax);
        if (row < nt) Eq[row * PITCH + ff] = val;
        else Ek[(row - nt) * PITCH + ff] = val;
      }
    }
    __syncthreads();
    for (int i = half; i < nt; i += 2) {
      float dq = 0.f;
      for (int j = 0; j < nc; ++j) dq = fmaf(dAs[i * nc + j], Ek[j * PITCH + f] + rho_eps, dq);
      float du = f < fc ? dq * Eq[i * PITCH + f] : 0.f;
      if (f < fc) dU[(((long long)t * nt + i) * H + h) * ldu + f0 + f] = du;
      du = warp_sum(du);
      if (lane == 0) rq[i * 4 + wq] += du;
    }
    for (int j = half; j < nc; j += 2) {
      float dk = 0.f;
      for (int i = 0; i < nt; ++i) dk = fmaf(dAs[i * nc + j], Eq[i * PITCH + f] + rho_eps, dk);
      float dw = f < fc ? dk * Ek[j * PITCH + f] : 0.f;
      if (f < fc) dW[(((long long)t * nc + j) * H + h) * ldu + f0 + f] = dw;
      dw = warp_sum(dw);
      if (lane == 0) rk[j * 4 + wq] += dw;
    }
  }
  __syncthreads();
  if (S > 1) cluster_sync();     // every rank's row-sum partials (shared memory) and dU / dW chunks (global) are complete
  if (rank == 0) {
    for (int i = tid; i < nt; i += 256) {
      const long long r = ((long long)t * nt + i) * H + h;
      float acc = 0.f;
      for (int q = 0; q < S; ++q)
        for (int w = 0; w < 4; ++w) acc += S > 1 ? ld_cluster(rq + i * 4 + w, (uint32_t)q) : rq[i * 4 + w];
      const float ds = -acc;
      ds_c2[r] = c2 * ds;
      dU[r * ldu + amq[r]] += ds;  // dm_i = ds_i routed to the first row-argmax (the element may be another rank's)
    }
    if (tid == 0) {
      float dg = 0.f;
      for (int j = 0; j < nc; ++j) {
        float acc = 0.f;
        for (int q = 0; q < S; ++q)
          for (int w = 0; w < 4; ++w) acc += S > 1 ? ld_cluster(rk + j * 4 + w, (uint32_t)q) : rk[j * 4 + w];
        const float dt = -acc;
        dt_c2[((long long)t * nc + j) * H + h] = c2 * dt;
        dg += dt;
      }
      dg_part[pair] = dg;
    }
  }
  if (S > 1) cluster_sync();     // rank 0 has read the other ranks' shared memory
}

__global__ void favor_key_fixup_kernel(float* __restrict__ dW, const float* __restrict__ W,
                                       const float* __restrict__ g, const float* __restrict__ dg,
                                       const float* __restrict__ ties, long long R, int M, long long ldu) {
  const float gmax = __ldg(g);
  const float share = __ldg(dg) / __ldg(ties);
  long long n = R * M;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, st = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += st) {
    long long r = i / M;
    int f = (int)(i - r * M);
    if (W[r * ldu + f] == gmax) dW[r * ldu + f] += share;
  }
}

// B200NP_FAVOR_SPLIT = 1 | 2 | 4 | 8 (default 4): CTAs per (task, head) of the forward and backward kernels
int favor_split() {
  static const int v = [] {
    const char* e = getenv("B200NP_FAVOR_SPLIT");
    const int s = (e && e[0]) ? atoi(e) : 4;
    return (s == 1 || s == 2 || s == 4 || s == 8) ? s : 4;
  }();
  return v;
}
size_t fwd_smem(int nt, int nc) { return (size_t)((nt + nc) * PITCH + nt * nc + 3 * nt + nc) * sizeof(float); }
size_t bwd_smem(int nt, int nc, int d) {
  return (size_t)((nt + nc) * PITCH + (nt + nc) * (d + 1) + 2 * nt * nc + 4 * nt + nc + 4 * (nt + nc)) * sizeof(float);
}
bool dims_ok(int T, int H, int nt, int nc, int d, int M, long long ldu) {
  return T > 0 && H > 0 && nt > 0 && nc > 0 && d > 0 && M > 0 && ldu >= M;
}

}  // namespace

extern "C" int b200np_favor_rowstats(const float* x, const float* U, float* diag, float* rowmax, int32_t* argmax,
                                     long long R, int d, int M, long long ldu, void* stream) {
  if (!x || !U || !diag || !rowmax || R <= 0 || d <= 0 || M <= 0 || ldu < M) return B200NP_E_BADARG;
  favor_rowstats_kernel<<<(unsigned)ceil_div(R, 8), 256, 0, as_stream(stream)>>>(x, U, diag, rowmax, argmax, R, d, M, ldu);
  return launch_status();
}

extern "C" int b200np_reduce(const float* x, long long n, float* out, int op, void* stream) {
  if (!x || !out || n <= 0 || (op != 0 && op != 1)) return B200NP_E_BADARG;
  reduce_kernel<<<1, 1024, 0, as_stream(stream)>>>(x, n, out, op);
  return launch_status();
}

extern "C" int b200np_favor_attn_fwd(const float* U, const float* W, const float* sq, const float* mq,
                                     const float* tk, const float* g, const float* v, float* out, float* A,
                                     float* Dn, float* ties, int T, int H, int nt, int nc, int d, int M,
                                     long long ldu, void* stream) {
  if (!U || !W || !sq || !mq || !tk || !g || !v || !out || !A || !Dn || !ties) return B200NP_E_BADARG;
  if (!dims_ok(T, H, nt, nc, d, M, ldu)) return B200NP_E_BADARG;
  if (nt * nc > 1024) return B200NP_E_UNSUPPORTED;   // the 4 x 256 accumulator layout of the A = Q'K'^T tile (evaluation: 36 x 25 = 900)
  size_t smem = fwd_smem(nt, nc);
  if (smem > 220 * 1024) return B200NP_E_UNSUPPORTED;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(favor_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return B200NP_E_LAUNCH;
  // S CTAs (one thread-block cluster) per (task, head), each with 1/S of the features
  int S = favor_split();
  while (S > 1 && (M + FC - 1) / FC < S) S >>= 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(T * H * S));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = S > 1 ? 1 : 0;
  if (cudaLaunchKernelEx(&cfg, favor_attn_fwd_kernel, U, W, sq, mq, tk, g, v, out, A, Dn, ties, H, nt, nc, d, M, ldu, S) !=
      cudaSuccess)
    return B200NP_E_LAUNCH;
  return launch_status();
}

extern "C" int b200np_favor_attn_bwd(const float* d_out, const float* U, const float* W, const float* sq,
                                     const float* mq, const int32_t* amq, const float* tk, const float* g,
                                     const float* v, const float* out, const float* A, const float* Dn, float* dU,
                                     float* dW, float* dv, float* ds_c2, float* dt_c2, float* dg_part, int T, int H,
                                     int nt, int nc, int d, int M, long long ldu, void* stream) {
  if (!d_out || !U || !W || !sq || !mq || !amq || !tk || !g || !v || !out || !A || !Dn || !dU || !dW || !dv ||
      !ds_c2 || !dt_c2 || !dg_part)
    return B200NP_E_BADARG;
  if (!dims_ok(T, H, nt, nc, d, M, ldu)) return B200NP_E_BADARG;
  if (nt * nc > 1024) return B200NP_E_UNSUPPORTED;   // the 4 x 256 accumulator layout of the A = Q'K'^T tile (evaluation: 36 x 25 = 900)
  size_t smem = bwd_smem(nt, nc, d);
  if (smem > 220 * 1024) return B200NP_E_UNSUPPORTED;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(favor_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return B200NP_E_LAUNCH;
  int S = favor_split();
  while (S > 1 && (M + FC - 1) / FC < S) S >>= 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(T * H * S));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = S > 1 ? 1 : 0;
  if (cudaLaunchKernelEx(&cfg, favor_attn_bwd_kernel, d_out, U, W, sq, mq, amq, tk, g, v, out, A, Dn, dU, dW, dv, ds_c2,
                         dt_c2, dg_part, H, nt, nc, d, M, ldu, S) != cudaSuccess)
    return B200NP_E_LAUNCH;
  return launch_status();
}

extern "C" int b200np_favor_key_fixup(float* dW, const float* W, const float* g, const float* dg, const float* ties,
                                      long long R, int M, long long ldu, void* stream) {
  if (!dW || !W || !g || !dg || !ties || R <= 0 || M <= 0 || ldu < M) return B200NP_E_BADARG;
  favor_key_fixup_kernel<<<ew_grid(R * M, 256), 256, 0, as_stream(stream)>>>(dW, W, g, dg, ties, R, M, ldu);
  return launch_status();
}
