// imfnet_b200 -- the ResNet stem (7x7 / stride 2 / padding 3 convolution of the RGB frame + BatchNorm + ReLU) as an implicit GEMM
//   x = self.conv1(x); x = self.bn1(x); x = self.relu(x)          /root/reference/model/resnet.py:195-207 (via model/Img_Encoder.py:15-18)
//
// The earlier form materialised the im2col matrix in HBM (k_image_im2col_h2: 640 bytes per output pixel, 491 MB per batch of ten 640x480
// frames) and ran the convolution kernel over it as a one-offset product: 215 + 171 us per batch.  Here:
//   * k_stem_presplit writes the frame once as two zero-padded fp16 HWC-4 images (hi and lo halves of every value, channel 3 = 0;
//     4 pad columns on the left, 3 pad rows on top), 16 bytes per pixel in total;
//   * for an output pixel (oy, ox) and kernel row ky the 8 padded pixels [2 ox, 2 ox + 8) of row 2 oy + ky are ONE contiguous, 16-byte
//     aligned 64-byte run per half: kernel columns -1 .. 6 (column -1 and channel 3 carry zero weights), K = 32 per kernel row;
//   * k_stem_conv reads those runs straight out of a copy of the input row in shared memory through an un-swizzled operand descriptor
//     whose rows overlap (see the kernel): no im2col matrix anywhere;
//   * the weights (4 slabs of two kernel rows, 64 KB) stay in shared memory for the CTA's lifetime; accumulators are double-buffered in
//     TMEM so the epilogue (BatchNorm affine, ReLU, hi/lo split, TMA store of the h2 tile) of one tile runs under the next tile's MMAs.
// Persistent grid; CTA = 10 warps: warp 0 loader (bulk copies), warp 1 MMA issuer + TMEM owner, warps 2-5 / 6-9 two epilogue groups
// (one TMEM lane quadrant per warp), alternating tiles.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"

namespace {

constexpr int kBM = 128;
constexpr int kImg = kBM * 128;                  // 16 KB
constexpr int kKY = 7;                           // kernel rows = stages per tile
constexpr int kCout = 64;
constexpr int kAccCols = 2 * kCout;
constexpr int kPadL = 4, kPadT = 3;

__host__ __device__ constexpr uint32_t st_idesc(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void st_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
struct __align__(16) SHalf8 { __half2 a, b, c, d; };

__device__ __forceinline__ float st_split16(const float* x, SHalf8* hi, SHalf8* lo) {
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
  hi[0] = SHalf8{h[0], h[1], h[2], h[3]};
  hi[1] = SHalf8{h[4], h[5], h[6], h[7]};
  lo[0] = SHalf8{l[0], l[1], l[2], l[3]};
  lo[1] = SHalf8{l[4], l[5], l[6], l[7]};
  return m;
}

// fp32 NCHW frames -> zero-padded fp16 HWC-4 images P[half][image][Hp][Wp][4] (half 0 = hi, 1 = lo); one thread per padded pixel
__global__ void __launch_bounds__(256) k_stem_presplit(const float* __restrict__ img, int H, int W, int Hp, int Wp, int num_images,
                                                       __half* __restrict__ P) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long per = (long long)Hp * Wp, total = per * num_images;
  if (idx >= total) return;
  const int b = (int)(idx / per);
  const int rem = (int)(idx - (long long)b * per);
  const int yp = rem / Wp, xp = rem - yp * Wp;
  const int y = yp - kPadT, x = xp - kPadL;
  float v[3] = {0.f, 0.f, 0.f};
  if (y >= 0 && y < H && x >= 0 && x < W) {
    const float* p = img + ((size_t)b * 3 * H + y) * W + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = __ldg(p + (size_t)c * H * W);
  }
  __half hi[4], lo[4];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    hi[c] = __float2half_rn(v[c]);
    lo[c] = __float2half_rn(v[c] - __half2float(hi[c]));
  }
  hi[3] = lo[3] = __float2half_rn(0.f);
  *reinterpret_cast<uint2*>(P + (size_t)idx * 4) = *reinterpret_cast<const uint2*>(hi);
  *reinterpret_cast<uint2*>(P + ((size_t)total + (size_t)idx) * 4) = *reinterpret_cast<const uint2*>(lo);
}

// Shared-memory operand descriptor WITHOUT swizzle, K-major: core matrix = 8 rows x 16 bytes with the rows 16 bytes apart; lbo = byte
// offset between the two 16-byte K chunks of one MMA (K = 16 halves), sbo = byte offset between 8-row groups.  Element (row r, chunk c)
// is read at start + (r / 8) * sbo + (r % 8) * 16 + c * lbo.
__device__ __forceinline__ uint64_t st_desc_noswz(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
  return d;                                        // layout type (bits 61-63) = 0: no swizzle
}

// The implicit GEMM.  Tile = 128 consecutive output pixels of ONE image row (ox0 = 128 j).  For kernel row ky the A operand of pixel
// ox0 + r is the 64-byte run of the padded input row 2 oy + ky that starts 16 r bytes after the run of pixel ox0: consecutive pixels'
// runs overlap by 48 bytes.  With lbo = 16 and sbo = 128 the un-swizzled descriptor reads element (r, c) at start + 16 (r + c), i.e.
// exactly that overlapping ("Toeplitz") view of ONE contiguous 2096-byte segment of the input row -- the im2col matrix is never formed,
// not even in shared memory.  A tile therefore needs 14 segments (7 kernel rows x hi / lo), 29 KB, fetched with 14 bulk copies by one
// thread; the first version assembled 128-byte operand rows with cp.async (112 KB per tile through the LSU and the L2) and ran at
// ~1100 cycles per kernel row, bound by the latency of its four-slot ring (profiles/r02: 185 us per batch of ten frames).
constexpr int kSeg = 2176;                       // >= 16 * 127 + 64 = 2096 bytes, kept a multiple of 128
constexpr int kSegBytes = 2096;
constexpr int kTileA = 2 * kKY * kSeg;           // 30464
constexpr int kNT = 3;                           // tiles in flight
constexpr int kWPair = 2 * kCout * 128;          // two kernel rows per slab: rows [0,64) = Whi, [64,128) = Wlo; K = [ky even 32 | ky odd 32]
constexpr int kNPair = (kKY + 1) / 2;
constexpr int kOutStage = 2 * kImg;              // hi image, lo image of one epilogue group's output tile
constexpr int kSmem2 = kNPair * kWPair + kNT * kTileA + 2 * kOutStage + 1024;
constexpr int kThreads2 = 320;

__global__ void __launch_bounds__(kThreads2, 1)
k_stem_conv(const __half* __restrict__ P, const __grid_constant__ CUtensorMap tmY, const unsigned char* __restrict__ Wp_, int Hp, int Wp,
            int H1, int W1, int num_images, const float* __restrict__ scale, const float* __restrict__ shift, __half* __restrict__ Y, int ldy,
            int* err) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* w_s = smem;                                 // kNPair x kWPair (64 KB)
  unsigned char* out_s = w_s + kNPair * kWPair;              // 2 x kOutStage (1024-aligned)
  unsigned char* a_s = out_s + 2 * kOutStage;                // kNT x kTileA
  __shared__ __align__(8) uint64_t a_full[kNT], a_empty[kNT], acc_full[2], acc_free[2], w_full;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float sc_s[kCout], sh_s[kCout];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int tpr = (W1 + kBM - 1) / kBM;                      // tiles per image row
  const int tiles = num_images * H1 * tpr;
  const int bx = blockIdx.x, gx = gridDim.x;
  if (bx >= tiles) return;
  const int cnt = (tiles - bx + gx - 1) / gx;                // this CTA's tiles: bx, bx + gx, ...

  if (tid == 0) {
    for (int s = 0; s < kNT; ++s) { tc::mbar_init(&a_full[s], 1); tc::mbar_init(&a_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&acc_full[b], 1); tc::mbar_init(&acc_free[b], 4); }
    tc::mbar_init(&w_full, 1);
    tc::fence_barrier_init();
    tma::prefetch_map(&tmY);
  }
  if (warp == 1) { tc::tmem_alloc(&tmem_base_s, 256); tc::tmem_relinquish(); }
  if (tid >= 64 && tid < 64 + kCout) {
    sc_s[tid - 64] = __ldg(scale + tid - 64);
    sh_s[tid - 64] = __ldg(shift + tid - 64);
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_d = tmem_base_s;
  const size_t half_stride = (size_t)num_images * Hp * Wp * 4;            // halves between the hi and the lo image set

  if (warp == 0) {
    // =========================== loader: the weight slabs once, then 14 input-row segments per tile ===========================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(&w_full, kNPair * kWPair);
      for (int k = 0; k < kNPair; ++k) tc::bulk_g2s(w_s + k * kWPair, Wp_ + (size_t)k * kWPair, kWPair, &w_full);
      for (int k = 0; k < cnt; ++k) {
        const int t = bx + k * gx;
        const int j = t % tpr, row = t / tpr;               // row = b * H1 + oy
        const int b = row / H1, oy = row - b * H1;
        const uint32_t s = (uint32_t)k % kNT;
        tc::mbar_wait(&a_empty[s], (((uint32_t)k / kNT) & 1u) ^ 1u, err, 4);
        tc::mbar_arrive_expect_tx(&a_full[s], 2 * kKY * kSegBytes);
        const __half* src = P + (((size_t)b * Hp + 2 * oy) * Wp + 2 * (size_t)(j * kBM)) * 4;      // padded pixel (2 oy, 2 ox0)
        unsigned char* dst = a_s + s * kTileA;
        for (int ky = 0; ky < kKY; ++ky) {
          tc::bulk_g2s(dst + (2 * ky) * kSeg, src + (size_t)ky * Wp * 4, kSegBytes, &a_full[s]);
          tc::bulk_g2s(dst + (2 * ky + 1) * kSeg, src + half_stride + (size_t)ky * Wp * 4, kSegBytes, &a_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t id2 = st_idesc(kBM, 2 * kCout), id1 = st_idesc(kBM, kCout);
    const uint32_t a0 = __shfl_sync(0xffffffffu, tc::smem_u32(a_s), 0);
    const uint32_t w0 = __shfl_sync(0xffffffffu, tc::smem_u32(w_s), 0);
    const uint32_t td = __shfl_sync(0xffffffffu, tmem_d, 0);
    tc::mbar_wait(&w_full, 0u, err, 1);
    for (int k = 0; k < cnt; ++k) {
      const uint32_t buf = (uint32_t)k & 1u, s = (uint32_t)k % kNT;
      tc::mbar_wait(&acc_free[buf], (((uint32_t)k >> 1) & 1u) ^ 1u, err, 2);          // the epilogue has drained this accumulator
      tc::mbar_wait(&a_full[s], ((uint32_t)k / kNT) & 1u, err, 3);
      tc::tc_fence_after_sync();
      const uint32_t d = td + buf * kAccCols;
      const uint32_t at = a0 + s * kTileA;
      if (tc::elect_one()) {
#pragma unroll
        for (int ky = 0; ky < kKY; ++ky) {
          const uint64_t dhi = st_desc_noswz(at + (2 * ky) * kSeg, 16, 128), dlo = st_desc_noswz(at + (2 * ky + 1) * kSeg, 16, 128);
          const uint64_t dw = tc::smem_desc_sw128(w0 + (ky >> 1) * kWPair) + (uint64_t)((ky & 1) * 4);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            st_mma(d, dhi + (uint64_t)(ks * 2), dw + (uint64_t)(ks * 2), id2, (ky | ks) ? 1u : 0u);          // hi . [Whi ; Wlo]
            st_mma(d, dlo + (uint64_t)(ks * 2), dw + (uint64_t)(ks * 2), id1, 1u);                           // lo . Whi
          }
        }
        tc::mma_commit(&a_empty[s]);
        tc::mma_commit(&acc_full[buf]);
      }
      __syncwarp();
    }
  } else {
    // =========================== epilogue: two groups of four warps, group g takes this CTA's tiles k = g, g + 2, ... ===========================
    const int g = (warp - 2) >> 2, q = warp & 3;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int r = q * 32 + lane;
    unsigned char* st_hi = out_s + g * kOutStage, *st_lo = st_hi + kImg;
    bool big = false;
    bool stored = false;
    for (int k = g; k < cnt; k += 2) {
      const uint32_t buf = (uint32_t)g;                          // tile k uses accumulator k & 1 = g
      const int t = bx + k * gx;
      const int j = t % tpr, row = t / tpr;
      const int valid = min(kBM, W1 - j * kBM);                  // pixels of this tile inside the image row
      const long long prow = (long long)row * W1 + j * kBM;      // output row of the tile's first pixel
      tc::mbar_wait(&acc_full[buf], ((uint32_t)k >> 1) & 1u, err, 5);
      tc::tc_fence_after_sync();
      if (stored) {                                              // this warp's previous stores have read its staging rows
        if (lane == 0) tma::store_wait_read<0>();
        __syncwarp();
      }
      const uint32_t d = tmem_d + lane_addr + buf * kAccCols;
      const bool full_box = q * 32 + 32 <= valid;                // the warp's 32 rows are all inside the image row: TMA store
      const bool live = r < valid;
      __half* yrow = Y + (size_t)(prow + r) * ldy;
#pragma unroll 1
      for (int cb = 0; cb < kCout; cb += 16) {
        uint32_t t1[16], t2[16];
        tc::tmem_ld16_issue(d + (uint32_t)cb, t1);
        tc::tmem_ld16_issue(d + (uint32_t)(kCout + cb), t2);
        tc::tmem_ld_wait();
        float a[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaxf(fmaf(__uint_as_float(t1[i]) + __uint_as_float(t2[i]), sc_s[cb + i], sh_s[cb + i]), 0.f);
        SHalf8 hi[2], lo[2];
        big |= !(st_split16(a, hi, lo) <= 60000.f) && live;
        if (full_box) {
          const int ch = cb >> 3;
          tc::st_shared_16(st_hi + tc::sw128_offset(r, ch), hi[0]);
          tc::st_shared_16(st_hi + tc::sw128_offset(r, ch + 1), hi[1]);
          tc::st_shared_16(st_lo + tc::sw128_offset(r, ch), lo[0]);
          tc::st_shared_16(st_lo + tc::sw128_offset(r, ch + 1), lo[1]);
        } else if (live) {                                       // a partly filled box (image width not a multiple of 32): direct stores
          tc::st_global_16(yrow + cb, hi[0]);
          tc::st_global_16(yrow + cb + 8, hi[1]);
          tc::st_global_16(yrow + kCout + cb, lo[0]);
          tc::st_global_16(yrow + kCout + cb + 8, lo[1]);
        }
      }
      tc::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_free[buf]);            // the accumulator is in registers / shared memory now
      stored = full_box;
      if (full_box) {
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma::store_2d(&tmY, tc::smem_u32(st_hi + q * 4096), 0, (int)(prow + q * 32));
          tma::store_2d(&tmY, tc::smem_u32(st_lo + q * 4096), 64, (int)(prow + q * 32));
          tma::store_commit();
        }
      }
    }
    if (lane == 0) tma::store_wait<0>();
    if (big && err) atomicOr(err, 0x10000);
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_d, 256);
}

inline void stem_geom(int H, int W, int& H1, int& W1, int& Hp, int& Wp) {
  H1 = (H - 1) / 2 + 1;
  W1 = (W - 1) / 2 + 1;
  Hp = 2 * (H1 - 1) + kKY;                       // rows 2 oy + ky, ky < 7
  if (Hp < H + kPadT) Hp = H + kPadT;
  Wp = 2 * (W1 - 1) + 8;                         // columns 2 ox .. 2 ox + 7
  if (Wp < W + kPadL) Wp = W + kPadL;
  Wp = (Wp + 1) & ~1;                            // 16-byte row pitch
}

}  // namespace

// bytes of the pre-split image set imf_image_stem_h2_fwd needs as workspace
extern "C" size_t imf_image_stem_workspace_bytes(int32_t H, int32_t W, int32_t num_images) {
  if (H <= 0 || W <= 0 || num_images <= 0) return 0;
  int H1, W1, Hp, Wp;
  stem_geom(H, W, H1, W1, Hp, Wp);
  return (size_t)2 * num_images * Hp * Wp * 4 * sizeof(__half) + 4096;          // + slack: a row's last tile reads a whole 2096-byte segment
}

// Y (h2 matrix, 64 channels, chunk width 64, ldy halves; rows = pixels of image 0, then image 1, ...) =
//   relu(scale * conv7x7/2/pad3(image) + shift)          image: fp32 [num_images, 3, H, W]
// packed = imf_sparse_conv_h2_pack of the kernel laid out as [4 (pairs of kernel rows ky = 2p, 2p + 1), 64, 64 (Cout)] with kc_in 64, where the
// 64 "input channels" of pair p are [row 2p: 8 columns kx = -1..6 x 4 channels | row 2p + 1: the same] (zeros at kx = -1, channel 3 and the
// missing row 7), its multiplier folded into scale.  Replaces conv1 -> bn1 -> relu of /root/reference/model/resnet.py:195-207.
extern "C" int imf_image_stem_h2_fwd(const float* image, int32_t H, int32_t W, int32_t num_images, const void* packed, const float* scale,
                                     const float* shift, void* workspace, size_t workspace_bytes, void* Y, int32_t ldy, int32_t* err,
                                     cudaStream_t stream) {
  IMF_CHECK_ARG(H > 0 && W > 0 && num_images >= 1 && ldy >= 2 * kCout && ldy % 8 == 0);
  IMF_CHECK_ARG(image != nullptr && packed != nullptr && scale != nullptr && shift != nullptr && workspace != nullptr && Y != nullptr);
  IMF_CHECK_ARG(((uintptr_t)packed % 16) == 0 && ((uintptr_t)workspace % 16) == 0 && ((uintptr_t)Y % 16) == 0);
  IMF_CHECK_ARG(workspace_bytes >= imf_image_stem_workspace_bytes(H, W, num_images));
  int H1, W1, Hp, Wp;
  stem_geom(H, W, H1, W1, Hp, Wp);
  const long long total_px = (long long)H1 * W1 * num_images;
  IMF_CHECK_ARG(total_px < (1ll << 31) - kBM);
  __half* P = reinterpret_cast<__half*>(workspace);
  const long long padded = (long long)Hp * Wp * num_images;
  k_stem_presplit<<<(unsigned)((padded + 255) / 256), 256, 0, stream>>>(image, H, W, Hp, Wp, num_images, P);
  IMF_CHECK_LAUNCH();
  CUtensorMap tmY;
  const int rc = tma::encode_2d_u16(&tmY, Y, (uint64_t)total_px, (uint64_t)(2 * kCout), (uint64_t)ldy, 64, 32);
  if (rc) { imf_set_error("cuTensorMapEncodeTiled (stem output) failed: %d", rc); return IMF_ERR_CUDA; }
  IMF_CHECK_CUDA(imf_set_max_smem_once(reinterpret_cast<const void*>(&k_stem_conv), kSmem2));
  const int tiles = num_images * H1 * ((W1 + kBM - 1) / kBM);
  const int grid = tiles < imf_sm_count() ? tiles : imf_sm_count();
  k_stem_conv<<<grid, kThreads2, kSmem2, stream>>>(P, tmY, reinterpret_cast<const unsigned char*>(packed), Hp, Wp, H1, W1, num_images, scale,
                                                   shift, reinterpret_cast<__half*>(Y), ldy, err);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}
