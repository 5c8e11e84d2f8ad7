// imfnet_b200 -- the network's tail in ONE kernel:  descriptors = L2norm( final( relu( conv1_tr([decoder | skip]) ) ) + bias )
//   out = self.conv1_tr(out); out = MEF.relu(out); out = self.final(out); out.F / torch.norm(out.F, p=2, dim=1, keepdim=True)
//   /root/reference/model/resunet.py:216-233 (conv1_tr and final are kernel_size-1 convolutions: per-voxel matrix products)
//
// The two-launch form (two one-offset runs of the convolution kernel + imf_h2_unpack_l2norm) moved the 64-channel hidden layer and the
// logits through HBM as h2 matrices: 236 us per 500 k voxels for 256 MB of compulsory traffic (384 B in, 128 B out per voxel).  Here a
// 128-row tile goes   TMA load (h2 rows) -> MMA 1 (3 split products, TMEM) -> ReLU + hi/lo split into a shared-memory operand tile ->
// MMA 2 -> bias, L2 norm over the row a thread owns -> fp32 tile in shared memory -> TMA store,
// so HBM sees the input once and the descriptors once.
//
// Persistent grid, CTA = 10 warps: warp 0 TMA producer (weights once, then the row tiles), warp 1 MMA issuer + TMEM owner,
// warps 2-5 / 6-9 two epilogue groups (one TMEM lane quadrant per warp, thread = row).  The k-th tile of a CTA belongs to group k & 1,
// which has its own input stage, accumulators (D1: 128 columns, D2: 64) and hidden tile, so the two groups' epilogues overlap each
// other's loads and MMAs.  Every hand-over is an mbarrier; the order in which the MMA warp issues (MMA1 g0, MMA1 g1, MMA2 g0, MMA2 g1)
// makes "accumulator / hidden tile free again" follow from the barriers that are already there (see the comments at the waits).
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"

namespace {

constexpr int kBM = 128;
constexpr int kImg = kBM * 128;                 // 16 KB: 128 rows x 128 bytes
constexpr int kMaxCh = 3;                       // input chunks of 32 channels (one [hi32 | lo32] line each)
constexpr int kC1 = 64, kC2 = 32;
constexpr int kW1Slab = 2 * kC1 * 128;          // per input chunk: rows [0,64) = [Whi | Whi], [64,128) = [Wlo | 0]
constexpr int kW2Bytes = 2 * kC2 * 128;         // rows [0,32) = Whi, [32,64) = Wlo (64 halves of K each)
constexpr int kAStage = kMaxCh * kImg;          // 48 KB
constexpr int kHBytes = 2 * kImg;               // hidden tile: hi image, lo image (re-used as the output staging tile)
constexpr int kThreads = 320;
constexpr int kD1Cols = 2 * kC1, kD2Cols = 2 * kC2;
constexpr int kSmem = 2 * kAStage + kMaxCh * kW1Slab + kW2Bytes + 2 * kHBytes + 1024;

__host__ __device__ constexpr uint32_t tf_idesc(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tf_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tf_tma_load(uint32_t dst, const CUtensorMap* map, uint32_t bar, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(col), "r"(row)
               : "memory");
}

struct __align__(16) THalf8 { __half2 a, b, c, d; };

__device__ __forceinline__ float tf_split16(const float* x, THalf8* hi, THalf8* lo) {
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
  hi[0] = THalf8{h[0], h[1], h[2], h[3]};
  hi[1] = THalf8{h[4], h[5], h[6], h[7]};
  lo[0] = THalf8{l[0], l[1], l[2], l[3]};
  lo[1] = THalf8{l[4], l[5], l[6], l[7]};
  return m;
}

__global__ void __launch_bounds__(kThreads, 1)
k_tail_fused(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmO, const unsigned char* __restrict__ W1p,
             const unsigned char* __restrict__ W2p, const int* __restrict__ n_ptr, int n_max, int nch, const float* __restrict__ scale1,
             const float* __restrict__ shift1, const float* __restrict__ scale2, const float* __restrict__ bias2, int normalize, int* err) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* a_st = smem;                                   // 2 x kAStage
  unsigned char* w1_s = a_st + 2 * kAStage;                     // kMaxCh x kW1Slab
  unsigned char* w2_s = w1_s + kMaxCh * kW1Slab;                // kW2Bytes
  unsigned char* h_s = w2_s + kW2Bytes;                         // 2 x kHBytes (1024-aligned: every size above is a multiple of 1024)
  __shared__ __align__(8) uint64_t a_full[2], a_empty[2], d1_full[2], h_full[2], d2_full[2], w_full;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float sc1_s[kC1], sh1_s[kC1], sc2_s[kC2], b2_s[kC2];

  int n = n_max;
  if (n_ptr) { const int v = *n_ptr; n = v < n_max ? v : n_max; }
  if (n <= 0) return;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int tiles = (n + kBM - 1) / kBM;
  const int bx = blockIdx.x, gx = gridDim.x;
  if (bx >= tiles) return;
  const int cnt = (tiles - bx + gx - 1) / gx;                   // this CTA's tiles: bx, bx + gx, ...

  if (tid == 0) {
    for (int g = 0; g < 2; ++g) {
      tc::mbar_init(&a_full[g], 1); tc::mbar_init(&a_empty[g], 1); tc::mbar_init(&d1_full[g], 1);
      tc::mbar_init(&h_full[g], 128); tc::mbar_init(&d2_full[g], 1);
    }
    tc::mbar_init(&w_full, 1);
    tc::fence_barrier_init();
    tma::prefetch_map(&tmX);
    tma::prefetch_map(&tmO);
  }
  if (warp == 1) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  if (tid >= 64 && tid < 64 + kC1) {
    const int c = tid - 64;
    sc1_s[c] = __ldg(scale1 + c);
    sh1_s[c] = shift1 ? __ldg(shift1 + c) : 0.f;
    if (c < kC2) { sc2_s[c] = __ldg(scale2 + c); b2_s[c] = bias2 ? __ldg(bias2 + c) : 0.f; }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_d = tmem_base_s;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      const uint32_t wbytes = (uint32_t)(nch * kW1Slab + kW2Bytes);
      tc::mbar_arrive_expect_tx(&w_full, wbytes);
      tc::bulk_g2s(w1_s, W1p, (uint32_t)(nch * kW1Slab), &w_full);
      tc::bulk_g2s(w2_s, W2p, kW2Bytes, &w_full);
      for (int k = 0; k < cnt; ++k) {
        const int g = k & 1, i = k >> 1;
        const int m0 = (bx + k * gx) * kBM;
        tc::mbar_wait(&a_empty[g], (uint32_t)(i & 1) ^ 1u, err, 1);            // MMA 1 of this group's previous tile has read the stage
        tc::mbar_arrive_expect_tx(&a_full[g], (uint32_t)(nch * kImg));
        for (int c = 0; c < nch; ++c) tf_tma_load(tc::smem_u32(a_st + g * kAStage + c * kImg), &tmX, tc::smem_u32(&a_full[g]), c * 64, m0);
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (warp-uniform code) ===========================
    constexpr uint32_t id1_2 = tf_idesc(kBM, 2 * kC1), id1_1 = tf_idesc(kBM, kC1), id2_2 = tf_idesc(kBM, 2 * kC2), id2_1 = tf_idesc(kBM, kC2);
    const uint32_t a0 = __shfl_sync(0xffffffffu, tc::smem_u32(a_st), 0);
    const uint32_t w1a = __shfl_sync(0xffffffffu, tc::smem_u32(w1_s), 0);
    const uint32_t w2a = __shfl_sync(0xffffffffu, tc::smem_u32(w2_s), 0);
    const uint32_t ha = __shfl_sync(0xffffffffu, tc::smem_u32(h_s), 0);
    const uint32_t td = __shfl_sync(0xffffffffu, tmem_d, 0);
    tc::mbar_wait(&w_full, 0u, err, 2);
    auto mma1 = [&](int k) {
      const int g = k & 1, i = k >> 1;
      // D1[g] is free: MMA 2 of this group's previous tile was issued after h_full[g], which the group's threads arrive on only after
      // they have read D1[g]; the stage itself is guarded by a_full.
      tc::mbar_wait(&a_full[g], (uint32_t)(i & 1), err, 3);
      tc::tc_fence_after_sync();
      const uint32_t d = td + (uint32_t)(g * kD1Cols);
      if (tc::elect_one()) {
        for (int c = 0; c < nch; ++c) {
          const uint64_t da = tc::smem_desc_sw128(a0 + g * kAStage + c * kImg), dw = tc::smem_desc_sw128(w1a + c * kW1Slab);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            tf_mma(d, da + (uint64_t)(ks * 2), dw + (uint64_t)(ks * 2), id1_2, (c | ks) ? 1u : 0u);          // hi . [Whi | Wlo]
            tf_mma(d, da + (uint64_t)(4 + ks * 2), dw + (uint64_t)(4 + ks * 2), id1_1, 1u);                  // lo . Whi
          }
        }
        tc::mma_commit(&a_empty[g]);
        tc::mma_commit(&d1_full[g]);
      }
      __syncwarp();
    };
    auto mma2 = [&](int k) {
      const int g = k & 1, i = k >> 1;
      // D2[g] is free: h_full[g] of this tile comes after the group finished the previous tile's second epilogue (program order)
      tc::mbar_wait(&h_full[g], (uint32_t)(i & 1), err, 4);
      tc::tc_fence_after_sync();
      const uint32_t d = td + (uint32_t)(2 * kD1Cols + g * kD2Cols);
      const uint64_t dh = tc::smem_desc_sw128(ha + g * kHBytes), dw = tc::smem_desc_sw128(w2a);
      if (tc::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          tf_mma(d, dh + (uint64_t)(ks * 2), dw + (uint64_t)(ks * 2), id2_2, ks ? 1u : 0u);                  // hi . [Whi ; Wlo]
          tf_mma(d, dh + (uint64_t)(kImg / 16 + ks * 2), dw + (uint64_t)(ks * 2), id2_1, 1u);                // lo . Whi
        }
        tc::mma_commit(&d2_full[g]);
      }
      __syncwarp();
    };
    for (int k = 0; k < cnt; k += 2) {
      mma1(k);
      if (k + 1 < cnt) mma1(k + 1);
      mma2(k);
      if (k + 1 < cnt) mma2(k + 1);
    }
  } else {
    // =========================== epilogue groups ===========================
    const int g = (warp - 2) >> 2, q = warp & 3;                 // TMEM lane quadrant = warp % 4
    const int r = q * 32 + lane;                                 // row of the tile this thread owns
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t d1 = tmem_d + lane_addr + (uint32_t)(g * kD1Cols), d2 = tmem_d + lane_addr + (uint32_t)(2 * kD1Cols + g * kD2Cols);
    unsigned char* hg = h_s + g * kHBytes;
    unsigned char* stage = hg + q * 4096;                        // this warp's 32 output rows x 128 bytes (inside the hi image)
    bool big = false;
    for (int k = g, i = 0; k < cnt; k += 2, ++i) {
      const int m0 = (bx + k * gx) * kBM;
      // ---- hidden = relu(scale1 * (x . W1) + shift1) -> hi / lo operand tile ----
      if (i > 0) {                                               // the previous tile's output store has read this warp's staging rows
        if (lane == 0) tma::store_wait_read<0>();
        __syncwarp();
      }
      tc::mbar_wait(&d1_full[g], (uint32_t)(i & 1), err, 5);
      tc::tc_fence_after_sync();
#pragma unroll 1
      for (int cb = 0; cb < kC1; cb += 16) {
        uint32_t t1[16], t2[16];
        tc::tmem_ld16_issue(d1 + (uint32_t)cb, t1);
        tc::tmem_ld16_issue(d1 + (uint32_t)(kC1 + cb), t2);
        tc::tmem_ld_wait();
        float a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          a[j] = fmaxf(fmaf(__uint_as_float(t1[j]) + __uint_as_float(t2[j]), sc1_s[cb + j], sh1_s[cb + j]), 0.f);
        THalf8 hi[2], lo[2];
        big |= !(tf_split16(a, hi, lo) <= 60000.f) && (m0 + r < n);
        const int ch = cb >> 3;
        tc::st_shared_16(hg + tc::sw128_offset(r, ch), hi[0]);
        tc::st_shared_16(hg + tc::sw128_offset(r, ch + 1), hi[1]);
        tc::st_shared_16(hg + kImg + tc::sw128_offset(r, ch), lo[0]);
        tc::st_shared_16(hg + kImg + tc::sw128_offset(r, ch + 1), lo[1]);
      }
      tc::fence_proxy_async();
      tc::tc_fence_before_sync();
      tc::mbar_arrive(&h_full[g]);
      // ---- logits = scale2 * (hidden . W2) + bias; descriptor = logits / ||logits|| ----
      tc::mbar_wait(&d2_full[g], (uint32_t)(i & 1), err, 6);
      tc::tc_fence_after_sync();
      float x[kC2];
      {
        uint32_t t1[2][16], t2[2][16];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          tc::tmem_ld16_issue(d2 + (uint32_t)(16 * b), t1[b]);
          tc::tmem_ld16_issue(d2 + (uint32_t)(kC2 + 16 * b), t2[b]);
        }
        tc::tmem_ld_wait();
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int j = 0; j < 16; ++j)
            x[16 * b + j] = fmaf(__uint_as_float(t1[b][j]) + __uint_as_float(t2[b][j]), sc2_s[16 * b + j], b2_s[16 * b + j]);
      }
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < kC2; ++j) ss = fmaf(x[j], x[j], ss);
      const float inv = normalize ? 1.0f / sqrtf(ss) : 1.0f;      // (no epsilon, like the reference: an all-zero row gives NaN)
#pragma unroll
      for (int j = 0; j < kC2; ++j) x[j] = (m0 + r < n) ? x[j] * inv : 0.f;          // rows past the end of the level: zeros
      // MMA 2 has completed (d2_full), so the hidden tile may be overwritten: rows [32 q, 32 q + 32) of the hi image stage this warp's output
#pragma unroll
      for (int j = 0; j < kC2 / 4; ++j)
        tc::st_shared_16(stage + tc::sw128_offset(lane, j), make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]));
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma::store_2d(&tmO, tc::smem_u32(stage), 0, m0 + q * 32);
        tma::store_commit();
      }
    }
    if (lane == 0) tma::store_wait<0>();
    if (big && err) atomicOr(err, 0x10000);
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_d, 512);
}

}  // namespace

// out[i, :] = normalize( scale2 * (relu(scale1 * (X[i, :] . W1) + shift1) . W2) + bias2 )   for i < min(*n_dev, n_max)
// X: h2 matrix of c0 channels, chunk width 32 (ldx halves); packed1 / packed2: imf_sparse_conv_h2_pack(W1 as [1, c0, 64], kc_in 32) /
// (W2 as [1, 64, 32], kc_in 64), their power-of-two multipliers folded into scale1 / scale2; shift1 / bias2 optional; out fp32 [n_max, 32]
// rows of ldo floats (16-byte aligned, ldo % 4 == 0).  c0 in {32, 64, 96}, c1 == 64, c2 == 32 (other shapes: the two-launch form).
// Replaces conv1_tr -> MEF.relu -> final -> L2 normalisation of /root/reference/model/resunet.py:216-233.
extern "C" int imf_tail_fused_h2_fwd(const void* X, int32_t ldx, int32_t n_max, const int32_t* n_dev, int32_t c0, int32_t c1, int32_t c2,
                                     const void* packed1, const float* scale1, const float* shift1, const void* packed2, const float* scale2,
                                     const float* bias2, int32_t normalize, float* out, int32_t ldo, int32_t* err, cudaStream_t stream) {
  IMF_CHECK_ARG(n_max >= 0 && c0 > 0 && c0 % 32 == 0 && c0 <= 32 * kMaxCh && c1 == kC1 && c2 == kC2);
  IMF_CHECK_ARG(ldx >= 2 * c0 && ldx % 8 == 0 && ldo >= c2 && ldo % 4 == 0);
  if (n_max == 0) return IMF_OK;
  IMF_CHECK_ARG(X != nullptr && packed1 != nullptr && packed2 != nullptr && scale1 != nullptr && scale2 != nullptr && out != nullptr);
  IMF_CHECK_ARG(((uintptr_t)X % 16) == 0 && ((uintptr_t)packed1 % 16) == 0 && ((uintptr_t)packed2 % 16) == 0 && ((uintptr_t)out % 16) == 0);
  CUtensorMap tmX, tmO;
  int rc = tma::encode_2d_u16(&tmX, X, (uint64_t)n_max, (uint64_t)(2 * c0), (uint64_t)ldx, 64, kBM);
  if (rc) { imf_set_error("cuTensorMapEncodeTiled (tail input) failed: %d", rc); return IMF_ERR_CUDA; }
  rc = tma::encode_2d_u16(&tmO, out, (uint64_t)n_max, (uint64_t)(2 * c2), (uint64_t)(2 * ldo), 64, 32);
  if (rc) { imf_set_error("cuTensorMapEncodeTiled (tail output) failed: %d", rc); return IMF_ERR_CUDA; }
  IMF_CHECK_CUDA(imf_set_max_smem_once(reinterpret_cast<const void*>(&k_tail_fused), kSmem));
  const int tiles = (n_max + kBM - 1) / kBM;
  const int grid = tiles < imf_sm_count() ? tiles : imf_sm_count();
  k_tail_fused<<<grid, kThreads, kSmem, stream>>>(tmX, tmO, reinterpret_cast<const unsigned char*>(packed1),
                                                  reinterpret_cast<const unsigned char*>(packed2), n_dev, n_max, c0 / 32, scale1, shift1, scale2,
                                                  bias2, normalize, err);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}
