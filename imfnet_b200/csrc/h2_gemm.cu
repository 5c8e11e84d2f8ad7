// imfnet_b200 -- dense GEMM on fp16 hi/lo ("h2") operands: the projections and the GEGLU feed-forward of the fusion module
//   to_q / to_kv / to_out / net.0 (+ GEGLU) / net.2          /root/reference/model/attention_fusion.py:48-63, 79-95
// for ALL stride-8 tokens of a batch in one launch each (imf_attention_fusion_fwd_batched, dense.cu).
//
//   C[M, N] = alpha * A[M, K] . W[N, K]^T (+ bias[n]) (+ R[m, n])            A: h2 matrix (chunk width 64), W: packed h2 slabs
// The 3xTF32 kernel these launches replace (tc_gemm.cu) stages its operands through registers (global -> split -> st.shared) and
// reached ~10 % of the tensor peak at these shapes (520 us per batch of 10 fragments); here both operands are already split in HBM, so
// the loads are pure TMA (2-D tiled loads of A, bulk copies of the pre-packed weight slabs), the three split products take two
// tcgen05.mma per K step (D[:, 0:2BN] += a_hi . [Whi ; Wlo]^T, D[:, 0:BN] += a_lo . Whi^T, as in the convolution kernel), and the
// accumulators are double-buffered in TMEM so that the epilogue of one tile runs under the main loop of the next.
//
// Persistent grid; CTA = 6 warps: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5 epilogue (one TMEM lane quadrant each).
// Tiles are (128 rows) x (BN = 128 columns), N-tile fastest, so the CTAs working at the same time share their A rows in the L2.
// Epilogues: fp32 out (+ bias, + fp32 residual), h2 out (+ bias), GEGLU h2 out (value / gate columns interleaved per tile by the
// weight packing: tile t = [value 64t..64t+63 | gate 64t..64t+63]).
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"

namespace {

constexpr int kBM = 128, kBN = 128;
constexpr int kImg = kBM * 128;                 // 16 KB: 128 rows x 64 halves
constexpr int kABytes = 2 * kImg;               // hi + lo image of one 64-channel chunk
constexpr int kWBytes = 2 * kBN * 128;          // [Whi ; Wlo] slab of one chunk: 256 rows x 128 bytes
constexpr int kStage = kABytes + kWBytes;       // 64 KB
constexpr int kNS = 3;
constexpr int kThreads = 192;
constexpr int kAccCols = 2 * kBN;               // D1 | D2

__host__ __device__ constexpr uint32_t hg_idesc(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void hg_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void hg_tma_load(uint32_t dst, const CUtensorMap* map, uint32_t bar, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(col), "r"(row)
               : "memory");
}
__device__ __forceinline__ float hg_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

struct __align__(16) HHalf8 { __half2 a, b, c, d; };

// 16 floats -> fp16 hi / lo (8 + 8 per 16-byte vector); returns max |x|
__device__ __forceinline__ float hg_split16(const float* x, HHalf8* hi, HHalf8* lo) {
  __half2 h[8], l[8];
  float m = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float x0 = x[2 * i], x1 = x[2 * i + 1];
    m = fmaxf(m, fmaxf(fabsf(x0), fabsf(x1)));
    h[i] = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h[i]);
    l[i] = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  }
  hi[0] = HHalf8{h[0], h[1], h[2], h[3]};
  hi[1] = HHalf8{h[4], h[5], h[6], h[7]};
  lo[0] = HHalf8{l[0], l[1], l[2], l[3]};
  lo[1] = HHalf8{l[4], l[5], l[6], l[7]};
  return m;
}

// MODE 0: fp32 out  C[m, n] = alpha * acc + bias[n] + R[m, n]
// MODE 1: h2 out    C[m, n] = alpha * acc + bias[n]                      (chunk width 64, ldc in halves)
// MODE 2: GEGLU h2  C[m, 64 t + j] = (alpha * acc[j] + bias[64 t + j]) * gelu(alpha * acc[64 + j] + bias[N / 2 + 64 t + j])
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
k_h2_gemm(const __grid_constant__ CUtensorMap tmA, const unsigned char* __restrict__ Wp, const int* __restrict__ m_ptr, int M_max, int N,
          int nchunks, float alpha, const float* __restrict__ bias, const float* __restrict__ R, int ldr, void* __restrict__ Cout, int ldc,
          int* err) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full[kNS], empty[kNS], acc_full[2], acc_free[2];
  __shared__ uint32_t tmem_base_s;

  int M = M_max;
  if (m_ptr) { const int v = *m_ptr; M = v < M_max ? v : M_max; }
  if (M <= 0) return;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int ntn = N / kBN;
  const int tiles = ((M + kBM - 1) / kBM) * ntn;
  if ((int)blockIdx.x >= tiles) return;

  if (tid == 0) {
    for (int s = 0; s < kNS; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&acc_full[b], 1); tc::mbar_init(&acc_free[b], 4); }
    tc::fence_barrier_init();
    tma::prefetch_map(&tmA);
  }
  if (warp == 1) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_d = tmem_base_s;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int zt = t % ntn, m0 = (t / ntn) * kBM;
        for (int c = 0; c < nchunks; ++c, ++it) {
          const uint32_t s = it % kNS;
          tc::mbar_wait(&empty[s], ((it / kNS) & 1u) ^ 1u, err, 1);
          unsigned char* st = smem + s * kStage;
          tc::mbar_arrive_expect_tx(&full[s], kStage);
          hg_tma_load(tc::smem_u32(st), &tmA, tc::smem_u32(&full[s]), c * 128, m0);
          hg_tma_load(tc::smem_u32(st + kImg), &tmA, tc::smem_u32(&full[s]), c * 128 + 64, m0);
          tc::bulk_g2s(st + kABytes, Wp + ((size_t)c * ntn + zt) * kWBytes, kWBytes, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (warp-uniform: bare UTCHMMA) ===========================
    constexpr uint32_t idesc2 = hg_idesc(kBM, 2 * kBN), idesc1 = hg_idesc(kBM, kBN);
    const uint32_t s0 = __shfl_sync(0xffffffffu, tc::smem_u32(smem), 0);
    const uint32_t td = __shfl_sync(0xffffffffu, tmem_d, 0);
    uint32_t it = 0, tl = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++tl) {
      const uint32_t buf = tl & 1u;
      tc::mbar_wait(&acc_free[buf], ((tl >> 1) & 1u) ^ 1u, err, 2);          // the epilogue has drained this accumulator
      tc::tc_fence_after_sync();
      const uint32_t d = td + buf * kAccCols;
      for (int c = 0; c < nchunks; ++c, ++it) {
        const uint32_t s = it % kNS;
        tc::mbar_wait(&full[s], (it / kNS) & 1u, err, 3);
        tc::tc_fence_after_sync();
        const uint32_t a0 = s0 + s * kStage, w0 = a0 + kABytes;
        const uint64_t da = tc::smem_desc_sw128(a0), dw = tc::smem_desc_sw128(w0);
        const uint32_t first = (c == 0) ? 0u : 1u;
        if (tc::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            hg_mma(d, da + (uint64_t)(ks * 2), dw + (uint64_t)(ks * 2), idesc2, (ks == 0) ? first : 1u);
            hg_mma(d, da + (uint64_t)(kImg / 16 + ks * 2), dw + (uint64_t)(ks * 2), idesc1, 1u);
          }
          tc::mma_commit(&empty[s]);
          if (c == nchunks - 1) tc::mma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== epilogue (4 warps, TMEM lane quadrant = warp % 4) ===========================
    const int q = warp & 3;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t tl = 0;
    bool big = false;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++tl) {
      const uint32_t buf = tl & 1u;
      const int zt = t % ntn, m0 = (t / ntn) * kBM;
      const int m = m0 + q * 32 + lane;
      tc::mbar_wait(&acc_full[buf], (tl >> 1) & 1u, err, 4);
      tc::tc_fence_after_sync();
      const uint32_t d = tmem_d + lane_addr + buf * kAccCols;
      constexpr int NCOL = (MODE == 2) ? kBN / 2 : kBN;
#pragma unroll 1
      for (int cb = 0; cb < NCOL; cb += 16) {
        uint32_t t1[16], t2[16];
        float a[16];
        tc::tmem_ld16_issue(d + (uint32_t)cb, t1);
        tc::tmem_ld16_issue(d + (uint32_t)(kBN + cb), t2);
        if (MODE == 2) {
          uint32_t g1[16], g2[16];
          tc::tmem_ld16_issue(d + (uint32_t)(kBN / 2 + cb), g1);
          tc::tmem_ld16_issue(d + (uint32_t)(kBN + kBN / 2 + cb), g2);
          tc::tmem_ld_wait();
          const int nv = zt * (kBN / 2) + cb, ng = N / 2 + nv;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float x = fmaf(alpha, __uint_as_float(t1[i]) + __uint_as_float(t2[i]), bias ? __ldg(bias + nv + i) : 0.f);
            const float g = fmaf(alpha, __uint_as_float(g1[i]) + __uint_as_float(g2[i]), bias ? __ldg(bias + ng + i) : 0.f);
            a[i] = x * hg_gelu(g);
          }
        } else {
          tc::tmem_ld_wait();
          const int n0 = zt * kBN + cb;
#pragma unroll
          for (int i = 0; i < 16; ++i) a[i] = fmaf(alpha, __uint_as_float(t1[i]) + __uint_as_float(t2[i]), bias ? __ldg(bias + n0 + i) : 0.f);
        }
        if (m < M) {
          if (MODE == 0) {
            const int n0 = zt * kBN + cb;
            float* cp = reinterpret_cast<float*>(Cout) + (size_t)m * ldc + n0;
            if (R) {
              const float4* rp = reinterpret_cast<const float4*>(R + (size_t)m * ldr + n0);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 r = __ldg(rp + i);
                a[4 * i] += r.x; a[4 * i + 1] += r.y; a[4 * i + 2] += r.z; a[4 * i + 3] += r.w;
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(cp)[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
          } else {
            const int n0 = (MODE == 2 ? zt * (kBN / 2) : zt * kBN) + cb;          // output channel of a[0]
            HHalf8 hi[2], lo[2];
            big |= !(hg_split16(a, hi, lo) <= 60000.f);
            __half* yp = reinterpret_cast<__half*>(Cout) + (size_t)m * ldc + (n0 >> 6) * 128 + (n0 & 63);
            tc::st_global_16(yp, hi[0]);
            tc::st_global_16(yp + 8, hi[1]);
            tc::st_global_16(yp + 64, lo[0]);
            tc::st_global_16(yp + 72, lo[1]);
          }
        }
      }
      tc::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_free[buf]);
    }
    if (big && err) atomicOr(err, 0x10000);
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_d, 512);
}

template <int MODE>
int hg_launch(const CUtensorMap& tmA, const void* Wp, const int* m_dev, int M_max, int N, int K, float alpha, const float* bias, const float* R,
              int ldr, void* C, int ldc, int* err, cudaStream_t stream) {
  const size_t smem = (size_t)kNS * kStage + 1024;
  IMF_CHECK_CUDA(imf_set_max_smem_once(reinterpret_cast<const void*>(&k_h2_gemm<MODE>), (int)smem));
  const int tiles = ((M_max + kBM - 1) / kBM) * (N / kBN);
  const int grid = tiles < imf_sm_count() ? tiles : imf_sm_count();
  k_h2_gemm<MODE><<<grid, kThreads, smem, stream>>>(tmA, reinterpret_cast<const unsigned char*>(Wp), m_dev, M_max, N, K / 64, alpha, bias, R, ldr,
                                                    C, ldc, err);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

}  // namespace

// C = alpha * A . W^T (+ bias) (+ R).  A: h2 matrix [M_max, K] (chunk width 64, lda halves, 16-byte aligned; min(*m_dev, M_max) rows are
// computed); W packed with imf_sparse_conv_h2_pack(W^T as [1, K, N], kernel_volume 1, kc_in 64, wmul) -- fold 1 / wmul into alpha.
// K % 64 == 0, N % 128 == 0.  mode 0: fp32 C (ldc floats) + optional fp32 residual R; mode 1: h2 C (ldc halves, chunk width 64);
// mode 2: GEGLU h2 C [M, N / 2] with the value / gate columns interleaved per 128-column tile by the caller's packing and
// bias = the un-permuted [N] vector.  err (optional): watchdog codes, bit 16 = an h2 output left the fp16 range.
extern "C" int imf_h2_gemm(const void* A, int32_t lda, int32_t M_max, const int32_t* m_dev, const void* Wpacked, int32_t N, int32_t K,
                           float alpha, const float* bias, const float* R, int32_t ldr, int32_t mode, void* C, int32_t ldc, int32_t* err,
                           cudaStream_t stream) {
  IMF_CHECK_ARG(M_max >= 0 && N > 0 && K > 0 && K % 64 == 0 && N % kBN == 0 && lda >= 2 * K && lda % 8 == 0 && mode >= 0 && mode <= 2);
  if (M_max == 0) return IMF_OK;
  IMF_CHECK_ARG(A != nullptr && Wpacked != nullptr && C != nullptr && ((uintptr_t)A % 16) == 0 && ((uintptr_t)Wpacked % 16) == 0);
  IMF_CHECK_ARG(mode == 0 ? (ldc >= N && ldc % 4 == 0 && ((uintptr_t)C % 16) == 0 && (R == nullptr || (ldr >= N && ldr % 4 == 0 && ((uintptr_t)R % 16) == 0)))
                          : (R == nullptr && ldc % 8 == 0 && ldc >= 2 * (mode == 2 ? N / 2 : N) && ((uintptr_t)C % 16) == 0));
  CUtensorMap tmA;
  const int rc = tma::encode_2d_u16(&tmA, A, (uint64_t)M_max, (uint64_t)(2 * K), (uint64_t)lda, 64, kBM);
  if (rc) { imf_set_error("cuTensorMapEncodeTiled (h2 gemm) failed: %d", rc); return IMF_ERR_CUDA; }
  if (mode == 0) return hg_launch<0>(tmA, Wpacked, m_dev, M_max, N, K, alpha, bias, R, ldr, C, ldc, err, stream);
  if (mode == 1) return hg_launch<1>(tmA, Wpacked, m_dev, M_max, N, K, alpha, bias, R, ldr, C, ldc, err, stream);
  return hg_launch<2>(tmA, Wpacked, m_dev, M_max, N, K, alpha, bias, R, ldr, C, ldc, err, stream);
}
