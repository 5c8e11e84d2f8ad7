// imfnet_b200 -- device-side batch bookkeeping for the batched captured plan (imfnet_b200/batched.py).
//
// ResUNet2.transformer (/root/reference/model/resunet.py:237-273) splits the stride-8 rows by batch index on the HOST
// (a python loop with two synchronisations per item) and runs the fusion module once per item, each item against its own
// image.  In a captured CUDA graph no size may travel to the host, so the split is done here:
//   * imf_batch_segments_n : seg[b] = first stride-8 row of item b (rows are batch-sorted: the coarser coordinate sets keep
//                            first-occurrence order of batch-sorted input, oracle/sparse_ops.py::stride_coords),
//                            cnt[b] = its row count, clipped to the per-item capacity of the plan; flags in *err when an item
//                            overflows that capacity or when rows carry a batch index >= num_batches;
//   * imf_h2_unpack_seg    : rows [seg[b], seg[b] + cnt[b]) of an h2 matrix -> fp32 rows 0.. of the item's private buffer
//                            (the query matrix of imf_attention_fusion_fwd_m, which takes cnt[b] as its device-side M);
//   * imf_h2_pack_seg      : the item's fused fp32 rows back into rows [seg[b], ...) of the level's h2 matrix.
// Nothing here is arithmetic on features beyond the exact fp16 hi/lo split of csrc/sparse_conv_h2.cu::k_h2_pack.
#include <cuda_fp16.h>

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

__global__ void __launch_bounds__(256) k_h2_unpack_seg(const __half* __restrict__ H, int ldh, const int* __restrict__ seg_b,
                                                       const int* __restrict__ cnt_b, int cap, int C, int KC, float* __restrict__ X,
                                                       int ldx) {
  const int n = min(*cnt_b, cap);
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)n * C) return;
  const int row = (int)(idx / C), c = (int)(idx % C);
  const __half* p = H + (size_t)(*seg_b + row) * ldh + (c / KC) * 2 * KC + (c % KC);
  X[(size_t)row * ldx + c] = __half2float(p[0]) + __half2float(p[KC]);
}

__global__ void __launch_bounds__(256) k_h2_pack_seg(const float* __restrict__ X, int ldx, const int* __restrict__ seg_b,
                                                     const int* __restrict__ cnt_b, int cap, int C, int KC, __half* __restrict__ H, int ldh,
                                                     int* err) {
  const int n = min(*cnt_b, cap);
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)n * C) return;
  const int row = (int)(idx / C), c = (int)(idx % C);
  const float x = X[(size_t)row * ldx + c];
  const __half h = __float2half_rn(x);
  __half* p = H + (size_t)(*seg_b + row) * ldh + (c / KC) * 2 * KC + (c % KC);
  p[0] = h;
  p[KC] = __float2half_rn(x - __half2float(h));
  if (fabsf(x) > 60000.f && err) atomicOr(err, 0x10000);
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

extern "C" int imf_h2_unpack_seg(const void* H, int32_t ldh, const int32_t* seg_b_dev, const int32_t* cnt_b_dev, int32_t cap, int32_t C,
                                 int32_t KC, float* X, int32_t ldx, cudaStream_t stream) {
  IMF_CHECK_ARG(cap >= 0 && C > 0 && (KC == 32 || KC == 64) && C % KC == 0 && ldx >= C && ldh >= 2 * C);
  if (cap == 0) return IMF_OK;
  IMF_CHECK_ARG(X != nullptr && H != nullptr && seg_b_dev != nullptr && cnt_b_dev != nullptr);
  const long long total = (long long)cap * C;
  k_h2_unpack_seg<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const __half*>(H), ldh, seg_b_dev, cnt_b_dev, cap, C,
                                                                     KC, X, ldx);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

extern "C" int imf_h2_pack_seg(const float* X, int32_t ldx, const int32_t* seg_b_dev, const int32_t* cnt_b_dev, int32_t cap, int32_t C,
                               int32_t KC, void* H, int32_t ldh, int32_t* err, cudaStream_t stream) {
  IMF_CHECK_ARG(cap >= 0 && C > 0 && (KC == 32 || KC == 64) && C % KC == 0 && ldx >= C && ldh >= 2 * C);
  if (cap == 0) return IMF_OK;
  IMF_CHECK_ARG(X != nullptr && H != nullptr && seg_b_dev != nullptr && cnt_b_dev != nullptr);
  const long long total = (long long)cap * C;
  k_h2_pack_seg<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(X, ldx, seg_b_dev, cnt_b_dev, cap, C, KC, reinterpret_cast<__half*>(H), ldh,
                                                                   err);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}
