// "Tap convolution": the one implicit-GEMM formulation behind every channel-dense conv on the path.
//
//   dst[pix(n,oy,ox), co] = epilogue( sum_{t < ntaps} sum_{ci < Cin}
//                                       src_t[n, oy*in_s + dy_t, ox*in_s + dx_t, ci] * w_t[co, ci] )
//
// A tap names a source tensor (NHWC, zero outside its bounds), a spatial offset and a
// [Cout x Cin] weight slab.  With the right tap table this is
//   * forward RxR conv, any stride                       (taps = filter positions)
//   * conv2 + 1x1 stride-2 skip projection fused         (one extra tap on a second source)
//   * data gradient of a stride-1 conv                   (taps = mirrored filter positions)
//   * data gradient of a stride-2 conv (+ skip gradient) (one launch per input-pixel parity class,
//                                                         each with its dense subset of taps; the
//                                                         destination is written with stride 2)
// GEMM view: M = N*OH*OW pixels, N = Cout, K = ntaps*Cin.  Both operands are K-major
// (channels contiguous), which is what the tcgen05 path wants as well.
#pragma once
#include "common.cuh"

namespace b200np {

constexpr int kMaxTaps = 12;

struct Tap {
  int8_t src;   // 0 or 1
  int8_t dy;    // input row  = oy*in_s[src] + dy
  int8_t dx;    // input col  = ox*in_s[src] + dx
  int8_t slab;  // weight slab index inside w[src]
};

struct TapConvArgs {
  const float* src[2];
  int srcH[2], srcW[2], in_s[2];
  const float* w[2];   // [slabs][Cout][Cin] per source
  int Cin, Cout;
  const float* bias;   // nullable
  const float* bias2;  // nullable
  float* dst;
  const float* mask;   // nullable; same geometry as dst: result *= (mask > 0)
  const uint32_t* mask_bits;  // nullable, takes precedence over `mask`: the same gate as 1 bit per element,
                              // [dst pixel][Cout / 32] words, bit j of word h = channel 32 h + j (8 B instead of
                              // 256 B read per pixel); written by the forward kernels' ReLU epilogues
  uint32_t* relu_bits; // nullable output (tcgen05 kernels, Cout == 64): the gates (dst > 0) in the mask_bits format
  int N, OH, OW;       // pixel grid of this launch
  int dstH, dstW, dst_s, dst_oy, dst_ox;  // dst pixel = (oy*dst_s + dst_oy, ox*dst_s + dst_ox)
  int ntaps;
  Tap taps[kMaxTaps];
  int act;
};

struct TapWgradArgs {
  const float* src;    // [N, srcH, srcW, Cin]
  int srcH, srcW, in_s, Cin;
  const float* src2;   // optional second source for taps with src == 1 (fused skip projection), 64 channels
  int src2H, src2W, in_s2;
  const float* dy;     // [N, OH, OW, Cout]
  int N, OH, OW, Cout;
  int ntaps;
  Tap taps[kMaxTaps];  // dy/dx only
  float* part;         // [chunks][ntaps][Cout][Cin]
  float* part_db;      // nullable, tcgen05 kernel only: [chunks][Cout] column sums of dy (the bias gradient) --
                       // the kernel stages every dy element anyway, so db costs no second pass over dy
  int chunks;
  long long pix_per_chunk;  // multiple of 16
};

int launch_tapconv_simt(const TapConvArgs& a, cudaStream_t st);
int launch_tapwgrad_simt(const TapWgradArgs& a, cudaStream_t st);
// tcgen05 path (tapconv_umma.cu); returns B200NP_E_UNSUPPORTED when the shape does not fit
int launch_tapconv_umma(const TapConvArgs& a, int precision, cudaStream_t st);
int launch_tapwgrad_umma(const TapWgradArgs& a, int precision, cudaStream_t st);
// halo formulation of the 3x3 stride-1 (+ skip) weight gradient (tapwgrad_halo.cu): chunk count for the workspace
// (0: shape does not fit) and the launch, which sets a.chunks
int tapwgrad_halo_chunks(int N, int OH, int OW);
int tapwgrad_halo_s2_chunks(int N, int OH, int OW);   // stride-2 variant (2 x 16 pixel tiles)
int launch_tapwgrad_halo(TapWgradArgs& a, int precision, cudaStream_t st);
// Several outputs from one staged input ("classes"): tap t accumulates into output class tap_cls[t] (taps sorted
// by class), class c is written at dst pixel (oy*dst_s + oy_c, ox*dst_s + ox_c).  The four input-parity classes
// of a stride-2 data gradient read the same dY halo, so one launch stages it once for all four.
struct TapClasses {
  int ncls;              // 2..4
  int tap_cls[kMaxTaps];
  int oy[4], ox[4];
};
// halo-tile tcgen05 path for unit-input-stride stencils (tapconv_halo.cu); bp = tensor-core weight images
int launch_tapconv_halo(const TapConvArgs& a, const float* bp0, int nslabs0, const float* bp1, int nslabs1,
                        int precision, cudaStream_t st, const TapClasses* cls = nullptr);

}  // namespace b200np
