// imfnet_b200 -- attention of the fusion module as ONE tcgen05 kernel: S = q k^T, softmax, o = softmax(S) v
//   /root/reference/model/attention_fusion.py:84-93 (einsum 'b i d, b j d -> b i j', softmax(dim=-1), einsum 'b i j, b j d -> b i d'),
// one cross head of 128 channels, M point tokens (queries) against L image tokens (keys / values).
//
// Operands are fp16 hi/lo pairs ("h2", see sparse_conv_h2.cu): q, K and V^T are pre-split once, P = exp(S - m) is split on the fly;
// every product is accumulated as hi.hi + hi.lo + lo.hi with two tcgen05.mma (kind::f16) per K step by concatenating [hi ; lo]
// of the B operand along N, exactly as the sparse convolution does.  All operand tiles are dense, so they are fetched with tiled
// TMA loads (cp.async.bulk.tensor.2d, 128-byte swizzle) -- the [M, L] score matrix never exists in memory.
//
// CTA = one 128-query tile x one slice of the L tokens (flash-decoding style split, so 9 query tiles still fill the GPU):
//   warp 0   TMA producer (Q once, then K / V blocks of 64 tokens into a 2-stage ring)
//   warp 1   MMA issuer + TMEM owner: S (128 x 64, two halves) and O (128 x 128, two halves) live in TMEM
//   warps 2-5 softmax: one query row per thread; pass 1 finds the row maximum of the slice (S only), pass 2 recomputes S, writes
//            P = exp(S - m) (hi/lo) into shared memory for the P.V MMAs and sums the row -- no rescaling of O is ever needed.
// The slices are merged by k_flash_combine (log-sum-exp weights), which also produces the fp32 [M, 128] output.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"

namespace {

constexpr int kD = 128;                 // head dimension
constexpr int kTQ = 128;                // queries per CTA
constexpr int kTB = 64;                 // tokens per block
constexpr int kQImg = kTQ * 128;        // 16 KB: 128 rows x 64 halves
constexpr int kKImg = kTB * 128;        //  8 KB:  64 rows x 64 halves
constexpr int kVImg = kD * 128;         // 16 KB: 128 dims x 64 tokens
constexpr int kQBytes = 4 * kQImg;      // hi0 lo0 hi1 lo1
constexpr int kKBytes = 4 * kKImg;
constexpr int kVBytes = 2 * kVImg;      // Vhi Vlo
constexpr int kStage = kKBytes + kVBytes;
constexpr int kPBytes = 2 * kQImg;      // Phi Plo (128 rows x 64 tokens)
constexpr int kSmem = kQBytes + 2 * kStage + kPBytes;
constexpr int kThreads = 192;

__host__ __device__ constexpr uint32_t ff_idesc(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void ff_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void ff_tma_load(uint32_t dst, const CUtensorMap* map, uint32_t bar, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(col), "r"(row)
               : "memory");
}

struct __align__(16) FHalf8 { __half2 a, b, c, d; };

// K [L, 128] fp32 -> h2 [Lpad, 256 halves] (chunk width 64), V^T [128, ldv] fp32 -> h2 [128, 2*Lpad] over the token axis; padding = 0
__global__ void __launch_bounds__(256) k_flash_pack_kv(const float* __restrict__ K, const float* __restrict__ Vt, int ldv, int L, int Lpad,
                                                       __half* __restrict__ Kh, __half* __restrict__ Vh) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long nk = (long long)Lpad * kD;
  if (idx < nk) {
    const int t = (int)(idx / kD), c = (int)(idx % kD);
    const float v = t < L ? K[(size_t)t * kD + c] : 0.f;
    const __half h = __float2half_rn(v);
    __half* p = Kh + (size_t)t * (2 * kD) + (c >> 6) * 128 + (c & 63);
    p[0] = h;
    p[64] = __float2half_rn(v - __half2float(h));
  } else if (idx < 2 * nk) {
    const long long j = idx - nk;
    const int d = (int)(j / Lpad), t = (int)(j % Lpad);
    const float v = t < L ? Vt[(size_t)d * ldv + t] : 0.f;
    const __half h = __float2half_rn(v);
    __half* p = Vh + (size_t)d * (2 * Lpad) + (t >> 6) * 128 + (t & 63);
    p[0] = h;
    p[64] = __float2half_rn(v - __half2float(h));
  }
}

__global__ void __launch_bounds__(kThreads, 1)
k_flash_fusion(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
               const int* __restrict__ m_ptr, int M_max, int L, int blocks_per_split, float* __restrict__ Opart, float* __restrict__ ml,
               int* err) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* q_s = smem;
  unsigned char* ring = smem + kQBytes;
  unsigned char* p_s = ring + 2 * kStage;
  __shared__ __align__(8) uint64_t q_full, full[2], empty[2], s_ready, s_free, p_ready, p_free, o_done;
  __shared__ uint32_t tmem_base_s;

  int M = M_max;
  if (m_ptr) { const int v = *m_ptr; M = v < M_max ? v : M_max; }
  const int m0 = blockIdx.x * kTQ;
  if (m0 >= M) return;
  const int split = blockIdx.y;
  const int nblocks_all = (L + kTB - 1) / kTB;
  const int b_begin = min(nblocks_all, split * blocks_per_split);
  const int nb = min(nblocks_all, b_begin + blocks_per_split) - b_begin;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    tc::mbar_init(&q_full, 1);
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(&s_ready, 1);
    tc::mbar_init(&s_free, 128);
    tc::mbar_init(&p_ready, 128);
    tc::mbar_init(&p_free, 1);
    tc::mbar_init(&o_done, 1);
    tc::fence_barrier_init();
    tma::prefetch_map(&tmQ);
    tma::prefetch_map(&tmK);
    tma::prefetch_map(&tmV);
  }
  if (warp == 1) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_s = tmem_base_s;            // S: columns [0,128)  (S1 | S2)
  const uint32_t tmem_o = tmem_base_s + 128u;     // O: columns [128,384) (O1 | O2)

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0 && nb > 0) {      // (nothing may be in flight when the CTA exits: an empty slice loads nothing)
      tc::mbar_arrive_expect_tx(&q_full, kQBytes);
      for (int i = 0; i < 4; ++i) ff_tma_load(tc::smem_u32(q_s + i * kQImg), &tmQ, tc::smem_u32(&q_full), i * 64, m0);
      int it = 0;
      for (int pass = 0; pass < 2; ++pass) {
        for (int b = 0; b < nb; ++b, ++it) {
          const int st = it & 1;
          tc::mbar_wait(&empty[st], (((uint32_t)(it >> 1)) & 1u) ^ 1u, err, 1);
          unsigned char* kd = ring + st * kStage;
          const int t0 = (b_begin + b) * kTB;
          tc::mbar_arrive_expect_tx(&full[st], pass == 0 ? kKBytes : kStage);
          for (int i = 0; i < 4; ++i) ff_tma_load(tc::smem_u32(kd + i * kKImg), &tmK, tc::smem_u32(&full[st]), i * 64, t0);
          if (pass == 1) {
            ff_tma_load(tc::smem_u32(kd + kKBytes), &tmV, tc::smem_u32(&full[st]), (b_begin + b) * 128, 0);
            ff_tma_load(tc::smem_u32(kd + kKBytes + kVImg), &tmV, tc::smem_u32(&full[st]), (b_begin + b) * 128 + 64, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t id_s2 = ff_idesc(kTQ, 2 * kTB), id_s1 = ff_idesc(kTQ, kTB);       // N = 128 / 64
    constexpr uint32_t id_o2 = ff_idesc(kTQ, 2 * kD), id_o1 = ff_idesc(kTQ, kD);          // N = 256 / 128
    if (nb > 0) tc::mbar_wait(&q_full, 0u, err, 2);
    int it = 0;
    for (int pass = 0; pass < 2; ++pass) {
      for (int b = 0; b < nb; ++b, ++it) {
        const int st = it & 1;
        tc::mbar_wait(&full[st], ((uint32_t)(it >> 1)) & 1u, err, 3);
        tc::mbar_wait(&s_free, ((uint32_t)it & 1u) ^ 1u, err, 4);            // softmax warps have read the previous S
        tc::tc_fence_after_sync();
        // issued from `if (lane == 0)` every tcgen05.mma is wrapped by ptxas in an ELECT / R2UR / BRA.U.ANY loop (the same finding as
        // in sparse_conv_g4.cu); warp-uniform operands + elect.sync give bare UTCHMMAs
        const uint32_t q0 = __shfl_sync(0xffffffffu, tc::smem_u32(q_s), 0), k0 = __shfl_sync(0xffffffffu, tc::smem_u32(ring + st * kStage), 0);
        const uint32_t tmem_s_u = __shfl_sync(0xffffffffu, tmem_s, 0);
        if (tc::elect_one()) {
#define tmem_s tmem_s_u
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint32_t qhi = q0 + (2 * c) * kQImg, qlo = qhi + kQImg, kb = k0 + (2 * c) * kKImg;     // [Khi_c ; Klo_c] = 128 rows
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t o = ks * 32;
              ff_mma(tmem_s, tc::smem_desc_sw128(qhi + o), tc::smem_desc_sw128(kb + o), id_s2, (c | ks) ? 1u : 0u);
              ff_mma(tmem_s, tc::smem_desc_sw128(qlo + o), tc::smem_desc_sw128(kb + o), id_s1, 1u);
            }
          }
          tc::mma_commit(&s_ready);
          if (pass == 0) tc::mma_commit(&empty[st]);
        }
#undef tmem_s
        __syncwarp();
        if (pass == 1) {
          tc::mbar_wait(&p_ready, (uint32_t)b & 1u, err, 5);                  // P of this block is in shared memory
          tc::tc_fence_after_sync();
          const uint32_t p0 = __shfl_sync(0xffffffffu, tc::smem_u32(p_s), 0);
          const uint32_t v0 = __shfl_sync(0xffffffffu, tc::smem_u32(ring + st * kStage + kKBytes), 0);       // [Vhi ; Vlo] = 256 rows
          const uint32_t tmem_o_u = __shfl_sync(0xffffffffu, tmem_o, 0);
          if (tc::elect_one()) {
#define tmem_o tmem_o_u
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t o = ks * 32;
              ff_mma(tmem_o, tc::smem_desc_sw128(p0 + o), tc::smem_desc_sw128(v0 + o), id_o2, (b | ks) ? 1u : 0u);
              ff_mma(tmem_o, tc::smem_desc_sw128(p0 + kQImg + o), tc::smem_desc_sw128(v0 + o), id_o1, 1u);
            }
            tc::mma_commit(&empty[st]);
            tc::mma_commit(&p_free);
          }
#undef tmem_o
          __syncwarp();
        }
      }
    }
    if (tc::elect_one()) tc::mma_commit(&o_done);      // same thread as the MMAs above (elect.sync is deterministic per mask)
    __syncwarp();
  } else {
    // =========================== softmax (4 warps, one query row per thread) ===========================
    const int q = warp & 3;                         // TMEM lane quadrant of this warp
    const int r = q * 32 + lane;                    // row inside the tile
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float row_max = -INFINITY, row_sum = 0.f;
    int it = 0;
    for (int pass = 0; pass < 2; ++pass) {
      for (int b = 0; b < nb; ++b, ++it) {
        tc::mbar_wait(&s_ready, (uint32_t)it & 1u, err, 6);
        tc::tc_fence_after_sync();
        const int t0 = (b_begin + b) * kTB;
        if (pass == 1) tc::mbar_wait(&p_free, ((uint32_t)b & 1u) ^ 1u, err, 7);    // the previous block's P.V MMAs are done with p_s
#pragma unroll 1
        for (int cb = 0; cb < kTB; cb += 16) {
          float s1[16], s2[16];
          tc::tmem_ld16(tmem_s + lane_addr + (uint32_t)cb, s1);
          tc::tmem_ld16(tmem_s + lane_addr + (uint32_t)(kTB + cb), s2);
          if (pass == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (t0 + cb + i < L) row_max = fmaxf(row_max, s1[i] + s2[i]);
          } else {
            __half2 hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float p0 = (t0 + cb + 2 * i < L) ? expf(s1[2 * i] + s2[2 * i] - row_max) : 0.f;
              float p1 = (t0 + cb + 2 * i + 1 < L) ? expf(s1[2 * i + 1] + s2[2 * i + 1] - row_max) : 0.f;
              row_sum += p0 + p1;
              const __half h0 = __float2half_rn(p0), h1 = __float2half_rn(p1);
              hi[i] = __halves2half2(h0, h1);
              lo[i] = __halves2half2(__float2half_rn(p0 - __half2float(h0)), __float2half_rn(p1 - __half2float(h1)));
            }
            const int ch = cb >> 3;                 // 16-byte chunk index of token cb inside the 64-token (128-byte) row
            *reinterpret_cast<FHalf8*>(p_s + tc::sw128_offset(r, ch)) = FHalf8{hi[0], hi[1], hi[2], hi[3]};
            *reinterpret_cast<FHalf8*>(p_s + tc::sw128_offset(r, ch + 1)) = FHalf8{hi[4], hi[5], hi[6], hi[7]};
            *reinterpret_cast<FHalf8*>(p_s + kQImg + tc::sw128_offset(r, ch)) = FHalf8{lo[0], lo[1], lo[2], lo[3]};
            *reinterpret_cast<FHalf8*>(p_s + kQImg + tc::sw128_offset(r, ch + 1)) = FHalf8{lo[4], lo[5], lo[6], lo[7]};
          }
        }
        tc::tc_fence_before_sync();
        tc::mbar_arrive(&s_free);                   // S may be overwritten by the next block's MMAs
        if (pass == 1) {
          tc::fence_proxy_async();
          tc::mbar_arrive(&p_ready);
        }
      }
    }
    // ---- partial result of this slice: O (un-normalised), row maximum and row sum ----
    tc::mbar_wait(&o_done, 0u, err, 8);
    tc::tc_fence_after_sync();
    const int row = m0 + r;
    float* op = Opart + ((size_t)split * M_max + row) * kD;
#pragma unroll 1
    for (int cb = 0; cb < kD; cb += 16) {          // tcgen05.ld is warp-collective: every lane loads, valid rows store
      float a[16];
      if (nb > 0) {
        float a2[16];
        tc::tmem_ld16(tmem_o + lane_addr + (uint32_t)cb, a);
        tc::tmem_ld16(tmem_o + lane_addr + (uint32_t)(kD + cb), a2);
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] += a2[i];
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = 0.f;
      }
      if (row < M) {
#pragma unroll
        for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(op + cb)[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
      }
    }
    if (row < M) {
      ml[((size_t)split * M_max + row) * 2] = row_max;
      ml[((size_t)split * M_max + row) * 2 + 1] = row_sum;
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base_s, 512);
}

// out[row, :] = sum_s w_s O_s / sum_s w_s l_s,  w_s = exp(m_s - max_s m_s); one thread per (row, 4 channels)
__global__ void __launch_bounds__(256) k_flash_combine(const float* __restrict__ Opart, const float* __restrict__ ml, int nsplit,
                                                       const int* __restrict__ m_ptr, int M_max, float* __restrict__ out, int ldo) {
  int M = M_max;
  if (m_ptr) { const int v = *m_ptr; M = v < M_max ? v : M_max; }
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)M * (kD / 4)) return;
  const int row = (int)(idx / (kD / 4)), c = (int)(idx % (kD / 4)) * 4;
  float mx = -INFINITY;
  for (int s = 0; s < nsplit; ++s) mx = fmaxf(mx, ml[((size_t)s * M_max + row) * 2]);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float den = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float m = ml[((size_t)s * M_max + row) * 2];
    if (m == -INFINITY) continue;                  // empty slice
    const float w = expf(m - mx);
    den = fmaf(w, ml[((size_t)s * M_max + row) * 2 + 1], den);
    const float4 o = *reinterpret_cast<const float4*>(Opart + ((size_t)s * M_max + row) * kD + c);
    acc.x = fmaf(w, o.x, acc.x); acc.y = fmaf(w, o.y, acc.y); acc.z = fmaf(w, o.z, acc.z); acc.w = fmaf(w, o.w, acc.w);
  }
  const float inv = 1.0f / den;
  *reinterpret_cast<float4*>(out + (size_t)row * ldo + c) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
}

inline int ff_lpad(int L) { return (L + kTB - 1) / kTB * kTB; }
inline int ff_nsplit(int M_max, int L) {
  const int tiles = (M_max + kTQ - 1) / kTQ, nblocks = (L + kTB - 1) / kTB;
  int ns = (2 * imf_sm_count() + tiles - 1) / tiles;          // about two waves of CTAs when every tile is active
  if (ns > nblocks) ns = nblocks;
  if (ns > 32) ns = 32;
  return ns < 1 ? 1 : ns;
}

}  // namespace

// sizes of the h2 copies of K and V^T appended to a kv buffer, and of the flash workspace (partials of every slice)
size_t imf_flash_kv_h2_bytes(int L) { return (size_t)2 * ff_lpad(L) * kD * 2 * sizeof(__half); }
size_t imf_flash_workspace_bytes(int M_max, int L) {
  return (size_t)ff_nsplit(M_max, L) * (size_t)(M_max > 0 ? M_max : 1) * (kD + 2) * sizeof(float);
}

// K [L,128], V^T [128, ldv] (fp32) -> kvh2 = { Kh2 [Lpad, 256 halves], Vth2 [128, 2*Lpad halves] }
int imf_flash_pack_kv(const float* K, const float* Vt, int ldv, int L, void* kvh2, cudaStream_t stream) {
  const int Lpad = ff_lpad(L);
  __half* Kh = reinterpret_cast<__half*>(kvh2);
  __half* Vh = Kh + (size_t)Lpad * 2 * kD;
  const long long total = 2LL * Lpad * kD;
  k_flash_pack_kv<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(K, Vt, ldv, L, Lpad, Kh, Vh);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

// o[M, 128] (fp32, row stride ldo) = softmax(q k^T) v; qh2 = h2 matrix of the (already scaled) queries [M_max, 256 halves]
int imf_flash_attention(const void* qh2, int M_max, const int* m_dev, const void* kvh2, int L, float* o, int ldo, void* workspace,
                        size_t workspace_bytes, int* err, cudaStream_t stream) {
  IMF_CHECK_ARG(qh2 && kvh2 && o && workspace && M_max > 0 && L > 0 && workspace_bytes >= imf_flash_workspace_bytes(M_max, L));
  const int Lpad = ff_lpad(L);
  const __half* Kh = reinterpret_cast<const __half*>(kvh2);
  const __half* Vh = Kh + (size_t)Lpad * 2 * kD;
  CUtensorMap tmQ, tmK, tmV;
  int rc = tma::encode_2d_u16(&tmQ, qh2, (uint64_t)M_max, 2 * kD, 2 * kD, 64, kTQ);
  if (!rc) rc = tma::encode_2d_u16(&tmK, Kh, (uint64_t)Lpad, 2 * kD, 2 * kD, 64, kTB);
  if (!rc) rc = tma::encode_2d_u16(&tmV, Vh, (uint64_t)kD, (uint64_t)2 * Lpad, (uint64_t)2 * Lpad, 64, kD);
  if (rc) { imf_set_error("cuTensorMapEncodeTiled (flash attention) failed: %d", rc); return IMF_ERR_CUDA; }
  const int nsplit = ff_nsplit(M_max, L);
  const int nblocks = (L + kTB - 1) / kTB;
  const int bps = (nblocks + nsplit - 1) / nsplit;
  float* Opart = reinterpret_cast<float*>(workspace);
  float* ml = Opart + (size_t)nsplit * M_max * kD;
  IMF_CHECK_CUDA(imf_set_max_smem_once(reinterpret_cast<const void*>(&k_flash_fusion), kSmem + 1024));
  dim3 grid((M_max + kTQ - 1) / kTQ, nsplit);
  k_flash_fusion<<<grid, kThreads, kSmem + 1024, stream>>>(tmQ, tmK, tmV, m_dev, M_max, L, bps, Opart, ml, err);
  IMF_CHECK_LAUNCH();
  const long long total = (long long)M_max * (kD / 4);
  k_flash_combine<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(Opart, ml, nsplit, m_dev, M_max, o, ldo);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}
