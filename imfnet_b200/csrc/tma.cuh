// imfnet_b200 -- tensor-map (TMA) helpers: host-side descriptor encoding and the device-side tile::gather4 / tile store
// instructions used by the sparse-convolution kernels.  A gather4 copies four rows of a 2-D tensor, chosen by four
// independent row indices, into four consecutive 128-byte rows of a (swizzled) shared-memory tile; rows whose index
// is outside the tensor are filled with zeros, which is how absent neighbours (-1 in a kernel map) are materialised.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tma {

// 2-D map over a row-major matrix of 16-bit elements: `rows` x `cols`, row stride `ld` elements; box = box_cols x box_rows,
// 128-byte swizzle (box_cols * 2 bytes must be <= 128).  Returns a CUresult.
// cuTensorMapEncodeTiled is resolved through the runtime (cudaGetDriverEntryPoint) so the library carries no link-time
// dependency on libcuda.so and still loads on a machine without a driver (CPU-side ABI tests).
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline encode_tiled_fn encode_tiled() {
  static encode_tiled_fn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<encode_tiled_fn>(p);
  }();
  return fn;
}

inline int encode_2d_u16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                         uint32_t box_rows) {
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {ld * 2};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  encode_tiled_fn fn = encode_tiled();
  if (!fn) return -1;
  return (int)fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

#ifdef __CUDACC__
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// rows r0..r3 of the mapped matrix, columns [col, col + box_cols) -> 4 x (box_cols*2) bytes at smem_dst; completes `bar` by bytes
__device__ __forceinline__ void gather4(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int col, int r0, int r1, int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
// shared tile (box of the map) -> global at (col, row); rows/cols outside the tensor are clipped
__device__ __forceinline__ void store_2d(const CUtensorMap* map, uint32_t smem_src, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(smem_src), "r"(col), "r"(row)
               : "memory");
}
__device__ __forceinline__ void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void store_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
#endif

}  // namespace tma
