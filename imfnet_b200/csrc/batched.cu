// imfnet_b200 -- device-side batch bookkeeping for the batched captured plan (imfnet_b200/batched.py).
//
// ResUNet2.transformer (/root/reference/model/resunet.py:237-273) splits the stride-8 rows by batch index on the HOST
// (a python loop with two synchronisations per item) and runs the fusion module once per item, each item against its own
// image.  In a captured CUDA graph no size may travel to the host, so the split is done here:
//   * imf_batch_segments_n : seg[b] = first stride-8 row of item b (rows are batch-sorted: the coarser coordinate sets keep
//                            first-occurrence order of batch-sorted input, oracle/sparse_ops.py::stride_coords),
//                            cnt[b] = its row count, clipped to the per-item capacity of the plan; flags in *err when an item
//                            overflows that capacity or when rows carry a batch index >= num_batches;
// The fusion module itself then runs ONCE over all rows (imf_attention_fusion_fwd_batched, dense.cu): everything but the attention
// core is row-wise, and the attention kernel (flash_fusion.cu) takes seg / cnt as its per-item row ranges.

#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_batch_segments_n(const int4* __restrict__ coords, const int* __restrict__ n_ptr, int n_max, int B,
                                                          int cap_item, int* __restrict__ seg, int* __restrict__ cnt, int* err) {
  __shared__ int seg_s[257];
  int n = n_max;
  if (n_ptr) { const int v = *n_ptr; n = v < n_max ? v : n_max; }
  if (n < 0) n = 0;
  const int b = threadIdx.x;
  if (b <= B) {
    int lo = 0, hi = n;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (coords[mid].x < b) lo = mid + 1; else hi = mid;
    }
    seg_s[b] = lo;
    seg[b] = lo;
  }
  __syncthreads();
  if (b < B) {
    const int c = seg_s[b + 1] - seg_s[b];
    cnt[b] = c < cap_item ? c : cap_item;
    if (c > cap_item && err) atomicOr(err, 0x20000);
  }
  if (b == B && seg_s[B] != n && err) atomicOr(err, 0x40000);
}

}  // namespace

extern "C" int imf_batch_segments_n(const int32_t* coords, const int32_t* n_dev, int32_t n_max, int32_t num_batches, int32_t cap_item,
                                    int32_t* seg, int32_t* cnt, int32_t* err, cudaStream_t stream) {
  IMF_CHECK_ARG(seg != nullptr && cnt != nullptr && num_batches >= 1 && num_batches <= 255 && n_max >= 0 && cap_item >= 0);
  IMF_CHECK_ARG(coords != nullptr || n_max == 0);
  k_batch_segments_n<<<1, 256, 0, stream>>>(reinterpret_cast<const int4*>(coords), n_dev, n_max, num_batches, cap_item, seg, cnt, err);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}
