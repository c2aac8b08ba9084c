// TMA (cp.async.bulk.tensor) building blocks: host-side tensor-map encoding without linking libcuda, and the
// device-side tile load / store instructions (sm_100a inline PTX; SASS: UTMALDG / UTMASTG).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "umma.cuh"

namespace b200np {
namespace tma {

// A CUtensorMap travels to the kernel BY VALUE as a __grid_constant__ parameter (64-byte aligned, 128 bytes): no
// device allocation, no global descriptor cache, and a captured CUDA graph keeps its own copy with the launch.
struct alignas(64) Map {
  CUtensorMap m;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (the library links cudart only).
inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// fp32 tensor of rank 4, innermost dimension first.  `strides` are the byte strides of dimensions 1..3 (dimension 0
// is dense).  box[0] * 4 must not exceed the swizzle span (128 B for SWIZZLE_128B*).  Out-of-bounds box elements
// read as zero -- which is exactly a convolution's zero padding.  Returns false if the driver refuses the map.
inline bool encode_f32_4d(Map* out, const float* base, const uint64_t dims[4], const uint64_t strides[3],
                          const uint32_t box[4], CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t gd[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gs[3] = {strides[0], strides[1], strides[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return fn(&out->m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gd, gs, bx, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Tile load global -> shared, completion counted in bytes on an mbarrier.  `issue`: instruction-level predicate (the
// instruction takes uniform-register operands; see umma.cuh on why the issuing warp runs its loop with all lanes).
__device__ __forceinline__ void load_4d(uint32_t smem_dst, const Map* map, uint32_t bar, int c0, int c1, int c2, int c3,
                                        uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %7, 0;\n\t"
      "@q cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}"
      ::"r"(smem_dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(issue)
      : "memory");
}
// Tile store shared -> global (bulk async group).
__device__ __forceinline__ void store_4d(const Map* map, uint32_t smem_src, int c0, int c1, int c2, int c3, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %6, 0;\n\t"
      "@q cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n\t}"
      ::"l"(map), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest `N` bulk groups of this thread have finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void store_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void prefetch_map(const Map* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

}  // namespace tma
}  // namespace b200np
