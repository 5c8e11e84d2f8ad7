// imfnet_b200 -- self-test of the TMA tile::gather4 path the sparse convolution relies on (tools/tma_selftest.py):
// gathers 128 rows of an fp16 matrix by index into a 128-byte-swizzled shared tile and copies the raw tile out, and
// stores a shared tile back with a tiled TMA store, so the host can check layout, zero fill of absent rows and clipping.
#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"

namespace {

__global__ void __launch_bounds__(32) k_probe_gather4(const __grid_constant__ CUtensorMap map, const __grid_constant__ CUtensorMap omap,
                                                      const int* __restrict__ idx, int col, int out_row, unsigned char* __restrict__ raw,
                                                      int* err) {
  __shared__ __align__(1024) unsigned char tile[128 * 128];
  __shared__ __align__(8) uint64_t bar;
  const int lane = threadIdx.x;
  if (lane == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  __syncwarp();
  const int4 r = reinterpret_cast<const int4*>(idx)[lane];
  if (lane == 0) tc::mbar_arrive_expect_tx(&bar, 128 * 128);
  __syncwarp();
  tma::gather4(tc::smem_u32(tile) + lane * 512, &map, tc::smem_u32(&bar), col, r.x, r.y, r.z, r.w);
  tc::mbar_wait(&bar, 0u, err, 1);
  for (int i = lane; i < 128 * 128 / 16; i += 32) reinterpret_cast<int4*>(raw)[i] = reinterpret_cast<const int4*>(tile)[i];
  // round trip: the same tile stored as 128 consecutive rows starting at out_row of the output matrix
  tc::fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    tma::store_2d(&omap, tc::smem_u32(tile), col, out_row);
    tma::store_commit();
    tma::store_wait<0>();
  }
}

}  // namespace

// X: fp16 [n_rows, ld] (ld halves); idx: device int32[128] row indices (negative or >= n_rows -> zero rows);
// raw: device 16 KB, receives the shared tile as laid out by the TMA; O: fp16 [o_rows, ld] receives the tile at rows
// [out_row, out_row+128) (clipped), columns [col, col+64).  box_rows: second box dimension of the gather map (probe).
extern "C" int imf_debug_gather4(const void* X, int32_t ld, int32_t n_rows, const int32_t* idx, int32_t col, int32_t box_rows,
                                 void* raw, void* O, int32_t o_rows, int32_t out_row, int32_t* err, cudaStream_t stream) {
  IMF_CHECK_ARG(X && idx && raw && O && ld % 8 == 0 && col % 64 == 0 && col + 64 <= ld);
  CUtensorMap map, omap;
  int rc = tma::encode_2d_u16(&map, X, (uint64_t)n_rows, (uint64_t)ld, (uint64_t)ld, 64, (uint32_t)box_rows);
  if (rc) { imf_set_error("cuTensorMapEncodeTiled (gather) failed: %d", rc); return IMF_ERR_CUDA; }
  rc = tma::encode_2d_u16(&omap, O, (uint64_t)o_rows, (uint64_t)ld, (uint64_t)ld, 64, 128);
  if (rc) { imf_set_error("cuTensorMapEncodeTiled (store) failed: %d", rc); return IMF_ERR_CUDA; }
  k_probe_gather4<<<1, 32, 0, stream>>>(map, omap, idx, col, out_row, reinterpret_cast<unsigned char*>(raw), err);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

// ---- throughput probe: `nwarps` warps each issue `iters` rounds of one gather4 per lane (4 rows x 128 B) into a private
// 16 KB tile and wait for each round; out[0] = cycles for the whole loop (max over warps), per CTA 0.
namespace {
__global__ void __launch_bounds__(256) k_probe_gather4_rate(const __grid_constant__ CUtensorMap map, const int* __restrict__ idx, int n_idx,
                                                            int iters, int depth, long long* __restrict__ out, int* err) {
  extern __shared__ unsigned char dsm[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dsm) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[8][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  if (threadIdx.x == 0) {
    for (int w = 0; w < 8; ++w) for (int d = 0; d < 4; ++d) tc::mbar_init(&bars[w][d], 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  // each warp keeps `depth` (<= 4) rounds in flight, tiles of 16 KB each (nwarps * depth * 16 KB <= 192 KB)
  for (int it = 0; it < iters + depth; ++it) {
    if (it >= depth) tc::mbar_wait(&bars[warp][(it - depth) % depth], (uint32_t)((it - depth) / depth) & 1u, err, 1);
    if (it < iters) {
      const int d = it % depth;
      const int o = ((blockIdx.x * nwarps + warp) * iters + it) * 128 + lane * 4;
      const int4 r = *reinterpret_cast<const int4*>(idx + (o % n_idx));
      if (lane == 0) tc::mbar_arrive_expect_tx(&bars[warp][d], 128 * 128);
      __syncwarp();
      tma::gather4(tc::smem_u32(base + (warp * depth + d) * 16384) + lane * 512, &map, tc::smem_u32(&bars[warp][d]), 0, r.x, r.y, r.z, r.w);
    }
  }
  const long long t1 = clock64();
  if (lane == 0 && blockIdx.x == 0) out[warp] = t1 - t0;
}
}  // namespace

extern "C" int imf_debug_gather4_rate(const void* X, int32_t ld, int32_t n_rows, const int32_t* idx, int32_t n_idx, int32_t nwarps,
                                      int32_t iters, int32_t depth, int32_t nctas, long long* out, int32_t* err, cudaStream_t stream) {
  IMF_CHECK_ARG(X && idx && out && nwarps >= 1 && nwarps <= 8 && depth >= 1 && depth <= 4 && nwarps * depth <= 12 && n_idx % 128 == 0);
  CUtensorMap map;
  int rc = tma::encode_2d_u16(&map, X, (uint64_t)n_rows, (uint64_t)ld, (uint64_t)ld, 64, 1);
  if (rc) { imf_set_error("cuTensorMapEncodeTiled failed: %d", rc); return IMF_ERR_CUDA; }
  const size_t smem = (size_t)nwarps * depth * 16384 + 1024;
  IMF_CHECK_CUDA(cudaFuncSetAttribute(k_probe_gather4_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_probe_gather4_rate<<<nctas, nwarps * 32, smem, stream>>>(map, idx, n_idx, iters, depth, out, err);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}
