// imfnet_b200 -- dense GEMM on the 5th-gen tensor cores (tcgen05.mma kind::tf32, fp32 accumulators in TMEM), 3xTF32.
//
//   C[M,N] = alpha * A[M,K] . B[N,K]^T (+ bias[n]) (+ R[m,n])          (A, B row-major = K-major operands)
//   GEGLU mode:  B has 2*N rows;  C[m,n] = (a_n + bias[n]) * gelu(a_{n+N} + bias[n+N])
// Serves the attention-fusion module of the IMFNet descriptor path:
//   to_q / to_kv / QK^T / PV / to_out / FFN     /root/reference/model/attention_fusion.py:79-95, 57-59
//
// CTA = 288 threads, tile 128 x BN:
//   warps 0-7  producers: global -> registers -> split hi/lo -> swizzled (SW128, K-major) shared tiles of 32 K-columns;
//              afterwards the epilogue: TMEM -> registers -> bias / GEGLU / residual -> global
//   warp  8    allocates TMEM and issues tcgen05.mma (one lane); 12 MMAs (4 K-slices x 3 split products) per stage
// Stages form a 3-deep ring guarded by full/empty mbarriers; tcgen05.commit releases a stage when its MMAs are done.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kBM = 128, kBK = 32, kNS = 3;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// 4 consecutive K elements of row `p` starting at column k, zero-filled past K; vector path when aligned.
__device__ __forceinline__ float4 load4(const float* __restrict__ p, int k, int K, bool vec_ok) {
  if (vec_ok && k + 3 < K) return __ldg(reinterpret_cast<const float4*>(p + k));
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k < K) v.x = __ldg(p + k);
  if (k + 1 < K) v.y = __ldg(p + k + 1);
  if (k + 2 < K) v.z = __ldg(p + k + 2);
  if (k + 3 < K) v.w = __ldg(p + k + 3);
  return v;
}

template <int BN, bool GEGLU>
__global__ void __launch_bounds__(288, 1) k_tc_gemm(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                    float* __restrict__ C, int ldc, int M, int N, int K, float alpha,
                                                    const float* __restrict__ bias, const float* __restrict__ R, int ldr,
                                                    int chunks_per_split, int* err, const int* __restrict__ m_ptr) {
  constexpr int A_BYTES = kBM * 128, B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[kNS], empty_bar[kNS], acc_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * kBM;
  const int M_layout = M;                                  // split-K partial tiles are laid out with the host-side row count
  if (m_ptr) { const int v = *m_ptr; M = v < M ? v : M; }  // device-side row count: whole tiles beyond it exit before any barrier
  if (m0 >= M) return;
  const int n0 = blockIdx.x * (GEGLU ? BN / 2 : BN);      // first OUTPUT column of this tile
  // split-K: blockIdx.z owns K chunks [kc_begin, kc_end) and writes its partial tile to C + z*M*ldc (ldc == N there)
  const int nk_total = (K + kBK - 1) / kBK;
  const int kc_begin = blockIdx.z * chunks_per_split;
  const int nk = min(nk_total, kc_begin + chunks_per_split) - kc_begin;
  C += (size_t)blockIdx.z * (size_t)M_layout * ldc;

  if (tid == 0) {
    for (int s = 0; s < kNS; ++s) { tc::mbar_init(&full_bar[s], 256); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&acc_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 8) { tc::tmem_alloc(&tmem_base_s, BN); tc::tmem_relinquish(); }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_d = tmem_base_s;

  if (warp < 8) {
    // ------------------------------ producers ------------------------------
    const bool a_vec = (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const bool b_vec = (ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % kNS;
      const uint32_t ph = (uint32_t)(kc / kNS) & 1u;
      tc::mbar_wait(&empty_bar[s], ph ^ 1u, err, 1);
      unsigned char* st = smem + s * STAGE_BYTES;
      const int k0 = (kc_begin + kc) * kBK;
      // A tile: 128 rows x 8 chunks = 1024 chunks -> 4 per thread
      float4 va[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int id = it * 256 + tid, r = id >> 3, c = id & 7;
        const int m = m0 + r;
        va[it] = (m < M) ? load4(A + (size_t)m * lda, k0 + c * 4, K, a_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // B tile: BN rows x 8 chunks -> BN/32 per thread
      float4 vb[BN / 32];
#pragma unroll
      for (int it = 0; it < BN / 32; ++it) {
        const int id = it * 256 + tid, r = id >> 3, c = id & 7;
        int n;
        bool ok;
        if (GEGLU) {
          const int half = BN / 2;
          const int j = (r < half) ? r : r - half;
          ok = (n0 + j) < N;
          n = (r < half) ? (n0 + j) : (N + n0 + j);
        } else {
          n = n0 + r;
          ok = n < N;
        }
        vb[it] = ok ? load4(B + (size_t)n * ldb, k0 + c * 4, K, b_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int id = it * 256 + tid, r = id >> 3, c = id & 7;
        float4 hi, lo;
        tc::split_tf32(va[it], hi, lo);
        const uint32_t off = tc::sw128_offset(r, c);
        *reinterpret_cast<float4*>(st + off) = hi;
        *reinterpret_cast<float4*>(st + A_BYTES + off) = lo;
      }
#pragma unroll
      for (int it = 0; it < BN / 32; ++it) {
        const int id = it * 256 + tid, r = id >> 3, c = id & 7;
        float4 hi, lo;
        tc::split_tf32(vb[it], hi, lo);
        const uint32_t off = tc::sw128_offset(r, c);
        *reinterpret_cast<float4*>(st + 2 * A_BYTES + off) = hi;
        *reinterpret_cast<float4*>(st + 2 * A_BYTES + B_BYTES + off) = lo;
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(&full_bar[s]);
    }
    // ------------------------------ epilogue ------------------------------
    tc::mbar_wait(&acc_bar, 0u, err, 3);
    tc::tc_fence_after_sync();
    const int lane_base = (warp & 3) * 32;
    const int m = m0 + lane_base + lane;
    if (!GEGLU) {
      const int col_base = (warp >> 2) * (BN / 2);
#pragma unroll 1
      for (int cb = 0; cb < BN / 2; cb += 16) {
        float v[16];
        tc::tmem_ld16(tmem_d + ((uint32_t)lane_base << 16) + (uint32_t)(col_base + cb), v);
        if (m < M) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int n = n0 + col_base + cb + i;
            if (n < N) {
              float x = alpha * v[i];
              if (bias) x += __ldg(bias + n);
              if (R) x += R[(size_t)m * ldr + n];
              C[(size_t)m * ldc + n] = x;
            }
          }
        }
      }
    } else {
      // columns [0, BN/2) = value half, [BN/2, BN) = gate half; each warp-group half handles BN/4 output columns
      const int col_base = (warp >> 2) * (BN / 4);
#pragma unroll 1
      for (int cb = 0; cb < BN / 4; cb += 16) {
        float v[16], g[16];
        tc::tmem_ld16(tmem_d + ((uint32_t)lane_base << 16) + (uint32_t)(col_base + cb), v);
        tc::tmem_ld16(tmem_d + ((uint32_t)lane_base << 16) + (uint32_t)(BN / 2 + col_base + cb), g);
        if (m < M) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int n = n0 + col_base + cb + i;
            if (n < N) {
              float x = alpha * v[i], gt = alpha * g[i];
              if (bias) { x += __ldg(bias + n); gt += __ldg(bias + N + n); }
              x = x * gelu_erf(gt);
              if (R) x += R[(size_t)m * ldr + n];
              C[(size_t)m * ldc + n] = x;
            }
          }
        }
      }
    }
  } else {
    // ------------------------------ MMA issuer ------------------------------
    constexpr uint32_t idesc = tc::idesc_tf32(kBM, BN);
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % kNS;
      const uint32_t ph = (uint32_t)(kc / kNS) & 1u;
      tc::mbar_wait(&full_bar[s], ph, err, 2);
      tc::tc_fence_after_sync();
      // warp-uniform operands + elect.sync instead of `if (lane == 0)`, so that ptxas emits bare UTC*MMA instructions instead of an
      // ELECT / R2UR / BRA.U.ANY loop around each (see sparse_conv_g4.cu)
      const uint32_t a_hi = __shfl_sync(0xffffffffu, tc::smem_u32(smem + s * STAGE_BYTES), 0), a_lo = a_hi + A_BYTES;
      const uint32_t tmem_d_u = __shfl_sync(0xffffffffu, tmem_d, 0);
      if (tc::elect_one()) {
#define tmem_d tmem_d_u
        const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t o = ks * 32;
          tc::mma_tf32(tmem_d, tc::smem_desc_sw128(a_lo + o), tc::smem_desc_sw128(b_hi + o), idesc, (kc | ks) ? 1u : 0u);
          tc::mma_tf32(tmem_d, tc::smem_desc_sw128(a_hi + o), tc::smem_desc_sw128(b_lo + o), idesc, 1u);
          tc::mma_tf32(tmem_d, tc::smem_desc_sw128(a_hi + o), tc::smem_desc_sw128(b_hi + o), idesc, 1u);
        }
        tc::mma_commit(&empty_bar[s]);
        if (kc == nk - 1) tc::mma_commit(&acc_bar);
      }
#undef tmem_d
      __syncwarp();
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem_d, BN);
}

template <int BN, bool GEGLU>
int launch(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K, float alpha,
           const float* bias, const float* R, int ldr, int splits, int chunks_per_split, int* err, const int* m_ptr, cudaStream_t stream) {
  constexpr int STAGE_BYTES = 2 * kBM * 128 + 2 * BN * 128;
  const size_t smem = (size_t)kNS * STAGE_BYTES + 1024;
  IMF_CHECK_CUDA(cudaFuncSetAttribute(k_tc_gemm<BN, GEGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int out_per_tile = GEGLU ? BN / 2 : BN;
  dim3 grid((N + out_per_tile - 1) / out_per_tile, (M + kBM - 1) / kBM, splits);
  k_tc_gemm<BN, GEGLU><<<grid, 288, smem, stream>>>(A, lda, B, ldb, C, ldc, M, N, K, alpha, bias, R, ldr, chunks_per_split, err, m_ptr);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

// Sum split-K partials: C[m,n] = sum_z P[z][m][n] (+ bias[n]) (+ R[m,n]).
__global__ void __launch_bounds__(256) k_splitk_reduce(const float* __restrict__ P, int splits, int M, int N, const float* __restrict__ bias,
                                                       const float* __restrict__ R, int ldr, float* __restrict__ C, int ldc,
                                                       const int* __restrict__ m_ptr) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)M * N) return;
  const int m = (int)(idx / N), n = (int)(idx % N);
  if (m_ptr && m >= *m_ptr) return;
  float v = 0.f;
  for (int z = 0; z < splits; ++z) v += P[(size_t)z * M * N + idx];
  if (bias) v += __ldg(bias + n);
  if (R) v += R[(size_t)m * ldr + n];
  C[(size_t)m * ldc + n] = v;
}

}  // namespace

extern "C" size_t imf_tc_gemm_workspace_bytes(int32_t M, int32_t N, int32_t K) {
  // room for up to 16 split-K partial tiles
  return (size_t)16 * (size_t)(M > 0 ? M : 1) * (size_t)(N > 0 ? N : 1) * sizeof(float);
}

// C[M,N] = alpha * A . B^T (+bias) (+R); geglu != 0: B holds 2N rows and C = (.)_n * gelu((.)_{n+N}).
// workspace (optional, imf_tc_gemm_workspace_bytes) enables split-K when the tile grid alone would leave most SMs idle.
// `err` is an optional device int that receives a non-zero code if an in-kernel wait times out (the kernel then traps).
extern "C" int imf_tc_gemm(const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc, int32_t M, int32_t N,
                           int32_t K, float alpha, const float* bias, const float* R, int32_t ldr, int32_t geglu, void* workspace,
                           size_t workspace_bytes, int32_t* err, cudaStream_t stream) {
  return imf_tc_gemm_m(A, lda, B, ldb, C, ldc, M, nullptr, N, K, alpha, bias, R, ldr, geglu, workspace, workspace_bytes, err, stream);
}

// Same with an optional device-side row count: only min(*m_dev, M) rows are computed (M sizes the launch and the workspace).
extern "C" int imf_tc_gemm_m(const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc, int32_t M,
                             const int32_t* m_dev, int32_t N, int32_t K, float alpha, const float* bias, const float* R, int32_t ldr,
                             int32_t geglu, void* workspace, size_t workspace_bytes, int32_t* err, cudaStream_t stream) {
  IMF_CHECK_ARG(M >= 0 && N >= 0 && K >= 1 && lda >= K && ldb >= K && ldc >= N);
  if (M == 0 || N == 0) return IMF_OK;
  IMF_CHECK_ARG(A != nullptr && B != nullptr && C != nullptr);
  const int nk = (K + kBK - 1) / kBK;
  if (geglu) return launch<128, true>(A, lda, B, ldb, C, ldc, M, N, K, alpha, bias, R, ldr, 1, nk, err, m_dev, stream);
  // narrow outputs or few row tiles: smaller BN puts more CTAs on the SMs
  const int sms = imf_sm_count();
  const long long tiles128 = (long long)((M + 127) / 128) * ((N + 127) / 128);
  const bool bn64 = (N <= 64 || tiles128 < sms / 2);
  const long long tiles = bn64 ? (long long)((M + 127) / 128) * ((N + 63) / 64) : tiles128;
  int splits = 1;
  if (workspace != nullptr && tiles < sms / 2 && nk >= 16) {
    splits = (int)((sms + tiles - 1) / tiles);
    if (splits > 16) splits = 16;
    if (splits > nk / 4) splits = nk / 4;
    if (workspace_bytes < (size_t)splits * M * N * sizeof(float)) splits = 1;
  }
  if (splits <= 1) {
    if (bn64) return launch<64, false>(A, lda, B, ldb, C, ldc, M, N, K, alpha, bias, R, ldr, 1, nk, err, m_dev, stream);
    return launch<128, false>(A, lda, B, ldb, C, ldc, M, N, K, alpha, bias, R, ldr, 1, nk, err, m_dev, stream);
  }
  const int cps = (nk + splits - 1) / splits;
  splits = (nk + cps - 1) / cps;                       // no empty split
  float* P = reinterpret_cast<float*>(workspace);
  int rc;
  if (bn64) rc = launch<64, false>(A, lda, B, ldb, P, N, M, N, K, alpha, nullptr, nullptr, 0, splits, cps, err, m_dev, stream);
  else rc = launch<128, false>(A, lda, B, ldb, P, N, M, N, K, alpha, nullptr, nullptr, 0, splits, cps, err, m_dev, stream);
  if (rc) return rc;
  const long long total = (long long)M * N;
  k_splitk_reduce<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(P, splits, M, N, bias, R, ldr, C, ldc, m_dev);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}
