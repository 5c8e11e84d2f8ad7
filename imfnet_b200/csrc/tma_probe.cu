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
