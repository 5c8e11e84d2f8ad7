// imfnet_b200 -- nearest-neighbour search in descriptor space on the tensor cores (BASELINE config 3:
// 5000-keypoint L2 feature matching): nn[i] = argmin_j || A[i] - B[j] ||^2, first index wins exact ties.
//   /root/reference/util/uio.py:245-258 (two calls per fragment pair, scripts/evaluation_3dmatch.py:207-217), lib/eval.py:18-48.
//
// The distance matrix is a dense contraction (|a|^2 + |b|^2 - 2 a.b): the 5000 x 5000 x 32 product runs on tcgen05 with fp16 hi/lo
// operands (three split products in one MMA: A rows = [a_hi | a_lo], B rows = [b_hi | b_hi] and [b_lo | 0], N = 256), but an argmin
// taken on those fp32-class values could differ from the exact one where two candidates are (nearly) equidistant.  So the tensor cores
// only FILTER:
//   pass 1: m~[i] = min_j (|b_j|^2 - 2 S~[i, j])                                    (S~ = tensor-core product, error <= tau / 2)
//   pass 2: every j with |b_j|^2 - 2 S~[i, j] <= m~[i] + tau[i] is re-evaluated EXACTLY like the brute-force kernel (matching.cu:
//           (a - b)^2 summed in channel order in fp32) and merged with a 64-bit atomicMin on (distance bits << 32 | j).
// The true minimiser always passes the filter (tau bounds the error of S~ both ways), so the result equals the brute-force kernel's
// bit for bit, ties included, while only one or two candidates per query are evaluated in fp32.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"

namespace {

constexpr int kT = 128;                  // queries per CTA tile = candidates per block
constexpr int kImg = kT * 128;           // 16 KB: 128 rows x 64 halves
constexpr int kBBytes = 2 * kImg;        // [b_hi | b_hi] rows then [b_lo | 0] rows
constexpr int kNS = 4;                   // B ring depth
constexpr int kThreads = 192;

__host__ __device__ constexpr uint32_t nn_idesc(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void nn_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void nn_tma_load(uint32_t dst, const CUtensorMap* map, uint32_t bar, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(col), "r"(row)
               : "memory");
}
// monotone map float -> unsigned (so that atomicMin on the unsigned orders like the floats, negatives included)
__device__ __forceinline__ unsigned nn_ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float nn_unord(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// rows of X [n, C] (C <= 32, zero padded to 32) -> P [npad, 64 halves]: what = 0: [hi | lo]; 1: [hi | hi]; 2: [lo | 0]; norms (fp32, what == 1)
// and the maximum norm (atomicMax on the bits of a non-negative float)
__global__ void __launch_bounds__(256) k_nn_pack(const float* __restrict__ X, int ldx, int n, int npad, int C, int what, __half* __restrict__ P,
                                                 float* __restrict__ norm2, unsigned* __restrict__ max_norm2) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= npad) return;
  const float v = (row < n && lane < C) ? __ldg(X + (size_t)row * ldx + lane) : 0.f;
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn(v - __half2float(h));
  __half* p = P + (size_t)row * 64;
  if (what == 0) { p[lane] = h; p[32 + lane] = l; }
  else if (what == 1) { p[lane] = h; p[32 + lane] = h; }
  else { p[lane] = l; p[32 + lane] = __float2half_rn(0.f); }
  if (norm2) {
    float s = v * v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      norm2[row] = row < n ? s : INFINITY;          // padding candidates can never win
      if (row < n && max_norm2) atomicMax(max_norm2, __float_as_uint(s));
    }
  }
}

// PASS 1: rowmin[i] = min_j (nb[j] - 2 S[i,j]) as ordered bits.  PASS 2: exact re-evaluation of the candidates within tau of rowmin.
// grid = (query tiles, candidate chunks); CTA: warp 0 TMA producer, warp 1 MMA, warps 2-5 epilogue (TMEM lane quadrant = warp % 4).
template <int PASS, int C>
__global__ void __launch_bounds__(kThreads, 1)
k_nn_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2, int na, int nb,
        int blocks_per_chunk, const float* __restrict__ nb2, const float* __restrict__ na2, const unsigned* __restrict__ max_nb2,
        unsigned* __restrict__ rowmin, const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
        unsigned long long* __restrict__ best, int* err) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* a_s = smem;                       // 16 KB
  unsigned char* ring = smem + kImg;               // kNS x 32 KB
  __shared__ __align__(8) uint64_t a_full, full[kNS], empty[kNS], s_full[2], s_free[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int i0 = blockIdx.x * kT;
  const int nblk_all = (nb + kT - 1) / kT;
  const int b_begin = blockIdx.y * blocks_per_chunk;
  const int nblk = min(nblk_all, b_begin + blocks_per_chunk) - b_begin;
  if (nblk <= 0) return;
  if (tid == 0) {
    tc::mbar_init(&a_full, 1);
    for (int s = 0; s < kNS; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&s_full[s], 1); tc::mbar_init(&s_free[s], 4); }
    tc::fence_barrier_init();
  }
  if (warp == 1) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_d = tmem_base_s;             // two S buffers of 256 columns: [a.(b_hi|b_hi) | a.(b_lo|0)]

  if (warp == 0) {
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(&a_full, kImg);
      nn_tma_load(tc::smem_u32(a_s), &tmA, tc::smem_u32(&a_full), 0, i0);
      for (int b = 0; b < nblk; ++b) {
        const uint32_t s = b % kNS;
        tc::mbar_wait(&empty[s], ((b / kNS) & 1u) ^ 1u, err, 1);
        tc::mbar_arrive_expect_tx(&full[s], kBBytes);
        nn_tma_load(tc::smem_u32(ring + s * kBBytes), &tmB1, tc::smem_u32(&full[s]), 0, (b_begin + b) * kT);
        nn_tma_load(tc::smem_u32(ring + s * kBBytes + kImg), &tmB2, tc::smem_u32(&full[s]), 0, (b_begin + b) * kT);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = nn_idesc(kT, 2 * kT);
    const uint32_t a0 = __shfl_sync(0xffffffffu, tc::smem_u32(a_s), 0), r0 = __shfl_sync(0xffffffffu, tc::smem_u32(ring), 0);
    const uint32_t td = __shfl_sync(0xffffffffu, tmem_d, 0);
    tc::mbar_wait(&a_full, 0u, err, 2);
    for (int b = 0; b < nblk; ++b) {
      const uint32_t s = b % kNS, sb = b & 1u;
      tc::mbar_wait(&full[s], (b / kNS) & 1u, err, 3);
      tc::mbar_wait(&s_free[sb], ((b >> 1) & 1u) ^ 1u, err, 4);
      tc::tc_fence_after_sync();
      const uint64_t da = tc::smem_desc_sw128(a0), db = tc::smem_desc_sw128(r0 + s * kBBytes);
      if (tc::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) nn_mma(td + sb * 256u, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc, ks ? 1u : 0u);
        tc::mma_commit(&empty[s]);
        tc::mma_commit(&s_full[sb]);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int i = i0 + q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float best_d = INFINITY;
    float thr = 0.f;
    float a[C];
    if (PASS == 2) {
      // tau: the tensor-core product is exact to ~2^-20 |a||b| (dropped lo.lo term + fp32 accumulation of 64 terms); both the filter value
      // and the exact re-evaluation round |b|^2 and the sum to fp32 (~2^-22 of the magnitudes).  2^-16 of the magnitudes is a safe margin.
      const float bm = __uint_as_float(*max_nb2);
      const float an = i < na ? na2[i] : 0.f;
      thr = (i < na ? nn_unord(rowmin[i]) : 0.f) + 1.52587890625e-05f * (2.f * sqrtf(an * bm) + bm + an);
#pragma unroll
      for (int c = 0; c < C; ++c) a[c] = (i < na) ? __ldg(A + (size_t)i * lda + c) : 0.f;
    }
    unsigned long long best_key = 0xFFFFFFFFFFFFFFFFull;
    for (int b = 0; b < nblk; ++b) {
      const uint32_t sb = b & 1u;
      tc::mbar_wait(&s_full[sb], (b >> 1) & 1u, err, 5);
      tc::tc_fence_after_sync();
      const int j0 = (b_begin + b) * kT;
#pragma unroll 1
      for (int cb = 0; cb < kT; cb += 16) {
        uint32_t t1[16], t2[16];
        tc::tmem_ld16_issue(tmem_d + lane_addr + sb * 256u + (uint32_t)cb, t1);
        tc::tmem_ld16_issue(tmem_d + lane_addr + sb * 256u + (uint32_t)(kT + cb), t2);
        tc::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int j = j0 + cb + e;
          const float v = __ldg(nb2 + j) - 2.f * (__uint_as_float(t1[e]) + __uint_as_float(t2[e]));          // (+inf for padding candidates)
          if (PASS == 1) {
            best_d = fminf(best_d, v);
          } else if (v <= thr && j < nb && i < na) {
            float d = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) { const float t = a[c] - __ldg(Bm + (size_t)j * ldb + c); d = fmaf(t, t, d); }
            const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j;
            best_key = key < best_key ? key : best_key;
          }
        }
      }
      tc::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&s_free[sb]);
    }
    if (i < na) {
      if (PASS == 1) atomicMin(rowmin + i, nn_ord(best_d));
      else if (best_key != 0xFFFFFFFFFFFFFFFFull) atomicMin(best + i, best_key);
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_d, 512);
}

__global__ void k_nn_tc_unpack(const unsigned long long* __restrict__ best, int na, int* __restrict__ idx, float* __restrict__ d2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= na) return;
  const unsigned long long k = best[i];
  idx[i] = (k == 0xFFFFFFFFFFFFFFFFull) ? -1 : (int)(unsigned)(k & 0xFFFFFFFFu);
  if (d2) d2[i] = __uint_as_float((unsigned)(k >> 32));
}

inline size_t r256(size_t b) { return (b + 255) / 256 * 256; }
inline int pad128(int n) { return (n + kT - 1) / kT * kT; }

struct NnLayout { size_t best, rowmin, na2, nb2, maxn, Ap, B1, B2, total; };
inline NnLayout nn_layout(int na, int nb) {
  NnLayout L;
  const int nap = pad128(na > 0 ? na : 1), nbp = pad128(nb > 0 ? nb : 1);
  size_t off = 0;
  L.best = off;   off += r256((size_t)nap * 8);
  L.rowmin = off; off += r256((size_t)nap * 4);
  L.na2 = off;    off += r256((size_t)nap * 4);
  L.nb2 = off;    off += r256((size_t)nbp * 4);
  L.maxn = off;   off += 256;
  L.Ap = off;     off += r256((size_t)nap * 128);
  L.B1 = off;     off += r256((size_t)nbp * 128);
  L.B2 = off;     off += r256((size_t)nbp * 128);
  L.total = off;
  return L;
}

template <int C>
int nn_tc_run(const CUtensorMap& tmA, const CUtensorMap& tmB1, const CUtensorMap& tmB2, int na, int nb, const float* nb2, const float* na2,
              const unsigned* maxn, unsigned* rowmin, const float* A, int lda, const float* B, int ldb, unsigned long long* best, cudaStream_t stream) {
  const size_t smem = (size_t)kImg + (size_t)kNS * kBBytes + 1024;
  IMF_CHECK_CUDA(imf_set_max_smem_once(reinterpret_cast<const void*>(&k_nn_tc<1, C>), (int)smem));
  IMF_CHECK_CUDA(imf_set_max_smem_once(reinterpret_cast<const void*>(&k_nn_tc<2, C>), (int)smem));
  const int qt = (na + kT - 1) / kT, nblk = (nb + kT - 1) / kT;
  int chunks = (2 * imf_sm_count() + qt - 1) / qt;
  if (chunks > nblk) chunks = nblk;
  if (chunks < 1) chunks = 1;
  const int bpc = (nblk + chunks - 1) / chunks;
  dim3 grid(qt, (nblk + bpc - 1) / bpc);
  k_nn_tc<1, C><<<grid, kThreads, smem, stream>>>(tmA, tmB1, tmB2, na, nb, bpc, nb2, na2, maxn, rowmin, A, lda, B, ldb, best, nullptr);
  IMF_CHECK_LAUNCH();
  k_nn_tc<2, C><<<grid, kThreads, smem, stream>>>(tmA, tmB1, tmB2, na, nb, bpc, nb2, na2, maxn, rowmin, A, lda, B, ldb, best, nullptr);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

}  // namespace

extern "C" size_t imf_nn_search_tc_workspace_bytes(int32_t na, int32_t nb) { return nn_layout(na, nb).total; }

// idx[i] = argmin_j ||A[i,:C] - B[j,:C]||^2 (int32, -1 when nb == 0; first index wins exact ties), d2 (optional) the squared distance:
// the same results as imf_nn_search, bit for bit, with the distance matrix on the tensor cores.  C in {16, 32}; workspace 256-byte aligned.
extern "C" int imf_nn_search_tc(const float* A, int32_t lda, int32_t na, const float* B, int32_t ldb, int32_t nb, int32_t C, int32_t* idx, float* d2,
                                void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  IMF_CHECK_ARG(na >= 0 && nb >= 0 && (C == 16 || C == 32) && lda >= C && ldb >= C);
  if (na == 0) return IMF_OK;
  const NnLayout L = nn_layout(na, nb);
  IMF_CHECK_ARG(A != nullptr && idx != nullptr && workspace != nullptr && workspace_bytes >= L.total && ((uintptr_t)workspace % 256) == 0);
  IMF_CHECK_ARG(nb == 0 || B != nullptr);
  char* ws = reinterpret_cast<char*>(workspace);
  unsigned long long* best = reinterpret_cast<unsigned long long*>(ws + L.best);
  unsigned* rowmin = reinterpret_cast<unsigned*>(ws + L.rowmin);
  float* na2 = reinterpret_cast<float*>(ws + L.na2);
  float* nb2 = reinterpret_cast<float*>(ws + L.nb2);
  unsigned* maxn = reinterpret_cast<unsigned*>(ws + L.maxn);
  __half* Ap = reinterpret_cast<__half*>(ws + L.Ap);
  __half* B1 = reinterpret_cast<__half*>(ws + L.B1);
  __half* B2 = reinterpret_cast<__half*>(ws + L.B2);
  const int nap = pad128(na), nbp = pad128(nb > 0 ? nb : 1);
  IMF_CHECK_CUDA(cudaMemsetAsync(best, 0xFF, (size_t)na * 8, stream));
  IMF_CHECK_CUDA(cudaMemsetAsync(rowmin, 0xFF, (size_t)na * 4, stream));
  IMF_CHECK_CUDA(cudaMemsetAsync(maxn, 0, 4, stream));
  if (nb > 0) {
    k_nn_pack<<<(nap + 7) / 8, 256, 0, stream>>>(A, lda, na, nap, C, 0, Ap, na2, nullptr);
    IMF_CHECK_LAUNCH();
    k_nn_pack<<<(nbp + 7) / 8, 256, 0, stream>>>(B, ldb, nb, nbp, C, 1, B1, nb2, maxn);
    IMF_CHECK_LAUNCH();
    k_nn_pack<<<(nbp + 7) / 8, 256, 0, stream>>>(B, ldb, nb, nbp, C, 2, B2, nullptr, nullptr);
    IMF_CHECK_LAUNCH();
    CUtensorMap tmA, tmB1, tmB2;
    int rc = tma::encode_2d_u16(&tmA, Ap, (uint64_t)nap, 64, 64, 64, kT);
    if (!rc) rc = tma::encode_2d_u16(&tmB1, B1, (uint64_t)nbp, 64, 64, 64, kT);
    if (!rc) rc = tma::encode_2d_u16(&tmB2, B2, (uint64_t)nbp, 64, 64, 64, kT);
    if (rc) { imf_set_error("cuTensorMapEncodeTiled (nn search) failed: %d", rc); return IMF_ERR_CUDA; }
    int r2;
    if (C == 16) r2 = nn_tc_run<16>(tmA, tmB1, tmB2, na, nb, nb2, na2, maxn, rowmin, A, lda, B, ldb, best, stream);
    else r2 = nn_tc_run<32>(tmA, tmB1, tmB2, na, nb, nb2, na2, maxn, rowmin, A, lda, B, ldb, best, stream);
    if (r2) return r2;
  }
  k_nn_tc_unpack<<<(na + 255) / 256, 256, 0, stream>>>(best, na, idx, d2);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}
