// imfnet_b200 -- nearest-neighbour search in descriptor space (BASELINE config 3: 5000-keypoint L2 matching).
//
// Replaces the reference's per-query KD-tree loop (/root/reference/util/uio.py:245-258, two calls per fragment pair in
// scripts/evaluation_3dmatch.py:207-217) and its chunked torch variant (lib/eval.py:18-48 + lib/metrics.py:22-29):
//   nn[i] = argmin_j || A[i] - B[j] ||^2        (first index wins exact ties, like numpy/torch argmin)
// Brute force in fp32, (a-b)^2 summed in channel order (no |a|^2+|b|^2-2ab cancellation), one thread per query row, B streamed
// through shared memory in chunks; the chunks of B are spread over grid.y and merged with a 64-bit atomicMin on
// (distance bits << 32 | index): non-negative floats order like their bit patterns, so the result is exact and deterministic.
#include "common.cuh"

namespace {

constexpr int kQ = 128;        // queries per CTA (one per thread)
constexpr int kTB = 64;        // rows of B per shared-memory tile

template <int C>
__global__ void __launch_bounds__(kQ) k_nn_search(const float* __restrict__ A, int lda, int na, const float* __restrict__ B, int ldb, int nb,
                                                  int rows_per_chunk, unsigned long long* __restrict__ best) {
  __shared__ float bs[kTB][C];
  const int i = blockIdx.x * kQ + threadIdx.x;
  float a[C];
#pragma unroll
  for (int c = 0; c < C; ++c) a[c] = (i < na) ? __ldg(A + (size_t)i * lda + c) : 0.f;
  const int j_begin = blockIdx.y * rows_per_chunk, j_end = min(nb, j_begin + rows_per_chunk);
  float bd = INFINITY;
  int bj = -1;
  for (int j0 = j_begin; j0 < j_end; j0 += kTB) {
    __syncthreads();
    for (int t = threadIdx.x; t < kTB * C; t += kQ) {
      const int r = t / C, c = t % C;
      bs[r][c] = (j0 + r < j_end) ? __ldg(B + (size_t)(j0 + r) * ldb + c) : 0.f;
    }
    __syncthreads();
    const int lim = min(kTB, j_end - j0);
    for (int r = 0; r < lim; ++r) {
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) { const float t = a[c] - bs[r][c]; d = fmaf(t, t, d); }
      if (d < bd) { bd = d; bj = j0 + r; }
    }
  }
  if (i < na && bj >= 0) {
    const unsigned long long key = ((unsigned long long)__float_as_uint(bd) << 32) | (unsigned)bj;
    atomicMin(best + i, key);
  }
}

__global__ void k_nn_unpack(const unsigned long long* __restrict__ best, int na, int* __restrict__ idx, float* __restrict__ d2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= na) return;
  const unsigned long long k = best[i];
  idx[i] = (k == 0xFFFFFFFFFFFFFFFFull) ? -1 : (int)(unsigned)(k & 0xFFFFFFFFu);
  if (d2) d2[i] = __uint_as_float((unsigned)(k >> 32));
}

}  // namespace

extern "C" size_t imf_nn_search_workspace_bytes(int32_t na) { return (size_t)(na > 0 ? na : 1) * sizeof(unsigned long long); }

// idx[i] = argmin_j ||A[i,:C] - B[j,:C]||^2 (int32, -1 when nb == 0), d2 (optional) the squared distance.  C in {16, 32, 64}.
extern "C" int imf_nn_search(const float* A, int32_t lda, int32_t na, const float* B, int32_t ldb, int32_t nb, int32_t C, int32_t* idx,
                             float* d2, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  IMF_CHECK_ARG(na >= 0 && nb >= 0 && (C == 16 || C == 32 || C == 64) && lda >= C && ldb >= C);
  if (na == 0) return IMF_OK;
  IMF_CHECK_ARG(A != nullptr && idx != nullptr && workspace != nullptr && workspace_bytes >= imf_nn_search_workspace_bytes(na));
  IMF_CHECK_ARG(nb == 0 || B != nullptr);
  unsigned long long* best = reinterpret_cast<unsigned long long*>(workspace);
  IMF_CHECK_CUDA(cudaMemsetAsync(best, 0xFF, (size_t)na * sizeof(unsigned long long), stream));
  if (nb > 0) {
    const int qblocks = (na + kQ - 1) / kQ;
    int chunks = (2 * imf_sm_count() + qblocks - 1) / qblocks;                 // about two CTAs per SM
    const int max_chunks = (nb + kTB - 1) / kTB;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    const int rows_per_chunk = ((nb + chunks - 1) / chunks + kTB - 1) / kTB * kTB;
    dim3 grid(qblocks, (nb + rows_per_chunk - 1) / rows_per_chunk);
    if (C == 16) k_nn_search<16><<<grid, kQ, 0, stream>>>(A, lda, na, B, ldb, nb, rows_per_chunk, best);
    else if (C == 32) k_nn_search<32><<<grid, kQ, 0, stream>>>(A, lda, na, B, ldb, nb, rows_per_chunk, best);
    else k_nn_search<64><<<grid, kQ, 0, stream>>>(A, lda, na, B, ldb, nb, rows_per_chunk, best);
    IMF_CHECK_LAUNCH();
  }
  k_nn_unpack<<<(na + 255) / 256, 256, 0, stream>>>(best, na, idx, d2);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}
