// imfnet_b200 -- the 3x3 / stride-1 convolutions of ResNet layer1 (64 -> 64 channels, + BatchNorm, + residual, + ReLU) as an implicit
// GEMM that reads every input pixel ONCE per tile:
//   out = self.conv1(x); out = self.bn1(out); out = self.relu(out); out = self.conv2(out); out = self.bn2(out); out += identity; relu
//   /root/reference/model/resnet.py:60-76 (BasicBlock.forward), called for layer1 from :208-209 via model/Img_Encoder.py:15-18
//
// Through the sparse-convolution kernel a dense image pays the sparse price: nine gathers of every 256-byte pixel row (one per kernel
// tap) from the L2 -- 442 MB per layer for ten 160x120 feature maps, which is what bounded those launches (~80 us each, tensor pipe 35 %).
// Here the activation lives in a PLANE layout ("P8"): per image 16 planes (8 chunks of 8 channels x {hi, lo} fp16 halves) of
// (H + 2) x (W + 2) zero-bordered pixels, 16 bytes per pixel and plane.  A tile is 16 rows x 8 columns of output pixels; its 18 x 10
// input patch arrives as 16 TMA boxes (46 KB, 1.4 x the tile instead of 9 x), and the A operand of tap (dy, dx) is the SAME patch read
// through an un-swizzled K-major descriptor whose start address is shifted by ((1 + dy) * 10 + (1 + dx)) * 16 bytes:
//   8 pixels of a patch row are one core matrix (8 x 16 bytes, contiguous), the next tile row is 160 bytes further (stride byte offset),
//   the next 8-channel chunk one plane further (leading byte offset).
// The nine weight slabs (144 KB) stay in shared memory for the CTA's lifetime; the hi and lo halves of a patch travel through a
// three-slot ring separately (the 36 hi . [Whi ; Wlo] MMAs of a tile run while its lo half and the next tile's hi half are in flight);
// accumulators are double-buffered in TMEM and two epilogue groups alternate tiles (BatchNorm affine, residual, ReLU, hi/lo split,
// 16-byte stores into the output planes -- or pixel-major h2 rows for the layer that feeds the strided convolution of layer2).
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"

namespace {

constexpr int kC = 64;                           // channels in = out
constexpr int kTX = 8, kTY = 16;                 // output tile (pixels)
constexpr int kPW = kTX + 2, kPH = kTY + 2;      // input patch
constexpr int kPlaneBytes = kPW * kPH * 16;      // 2880 bytes of one plane's patch
constexpr int kPlaneStride = 2944;               // ... at 128-byte aligned offsets (TMA destination alignment)
constexpr int kHalf = 8 * kPlaneStride;          // hi (or lo) half of a patch: 8 chunk planes
constexpr int kSlots = 3;
constexpr int kWSlab = 2 * kC * 128;             // per tap: rows [0,64) = Whi, [64,128) = Wlo, 64 halves of K each (SW128)
constexpr int kTaps = 9;
constexpr int kThreads = 320;
constexpr int kAccCols = 2 * kC;
constexpr int kSmem = kTaps * kWSlab + kSlots * kHalf + 1024;

__host__ __device__ constexpr uint32_t ic_idesc(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void ic_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void ic_tma_load(uint32_t dst, const CUtensorMap* map, uint32_t bar, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(col), "r"(row)
               : "memory");
}
// un-swizzled K-major operand: element (row r, 16-byte chunk c) at start + (r / 8) * sbo + (r % 8) * 16 + c * lbo
__device__ __forceinline__ uint64_t ic_desc_noswz(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

struct __align__(16) IHalf8 { __half2 a, b, c, d; };

__device__ __forceinline__ float ic_split16(const float* x, IHalf8* hi, IHalf8* lo) {
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
  hi[0] = IHalf8{h[0], h[1], h[2], h[3]};
  hi[1] = IHalf8{h[4], h[5], h[6], h[7]};
  lo[0] = IHalf8{l[0], l[1], l[2], l[3]};
  lo[1] = IHalf8{l[4], l[5], l[6], l[7]};
  return m;
}
// x[0..8) += hi + lo of one 8-channel chunk (two 16-byte loads)
__device__ __forceinline__ void ic_add_chunk(const __half* hi_p, const __half* lo_p, float* x) {
  const int4 h4 = __ldg(reinterpret_cast<const int4*>(hi_p)), l4 = __ldg(reinterpret_cast<const int4*>(lo_p));
  const __half2* h = reinterpret_cast<const __half2*>(&h4);
  const __half2* l = reinterpret_cast<const __half2*>(&l4);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 hf = __half22float2(h[i]), lf = __half22float2(l[i]);
    x[2 * i] += hf.x + lf.x;
    x[2 * i + 1] += hf.y + lf.y;
  }
}

struct P8Geom {
  int H, W, Hp, Wp, B;
  int tiles_x, tiles_y, tiles;
  size_t plane_halves;          // halves per plane: Hp * Wp * 8
};
__host__ __device__ inline P8Geom p8_geom(int H, int W, int B) {
  P8Geom g;
  g.H = H; g.W = W; g.B = B;
  g.Hp = H + 2; g.Wp = W + 2;
  g.tiles_x = (W + kTX - 1) / kTX;
  g.tiles_y = (H + kTY - 1) / kTY;
  g.tiles = B * g.tiles_x * g.tiles_y;
  g.plane_halves = (size_t)g.Hp * g.Wp * 8;
  return g;
}

__global__ void __launch_bounds__(kThreads, 1)
k_image_conv3x3_p8(const __grid_constant__ CUtensorMap tmX, const unsigned char* __restrict__ Wp_, int H, int W, int B,
                   const float* __restrict__ scale, const float* __restrict__ shift, const __half* __restrict__ R, int relu,
                   __half* __restrict__ Y, int pixel_major, int ldy, int* err) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* w_s = smem;                                  // kTaps x kWSlab
  unsigned char* ring = w_s + kTaps * kWSlab;                 // kSlots x kHalf
  __shared__ __align__(8) uint64_t h_full[kSlots], h_empty[kSlots], acc_full[2], acc_free[2], w_full;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float sc_s[kC], sh_s[kC];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const P8Geom g = p8_geom(H, W, B);
  const int bx = blockIdx.x, gx = gridDim.x;
  if (bx >= g.tiles) return;
  const int cnt = (g.tiles - bx + gx - 1) / gx;               // this CTA's tiles: bx, bx + gx, ...

  if (tid == 0) {
    for (int s = 0; s < kSlots; ++s) { tc::mbar_init(&h_full[s], 1); tc::mbar_init(&h_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&acc_full[b], 1); tc::mbar_init(&acc_free[b], 4); }
    tc::mbar_init(&w_full, 1);
    tc::fence_barrier_init();
    tma::prefetch_map(&tmX);
  }
  if (warp == 1) { tc::tmem_alloc(&tmem_base_s, 256); tc::tmem_relinquish(); }
  if (tid >= 64 && tid < 64 + kC) {
    sc_s[tid - 64] = __ldg(scale + tid - 64);
    sh_s[tid - 64] = __ldg(shift + tid - 64);
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_d = tmem_base_s;
  const int tpi = g.tiles_x * g.tiles_y;                      // tiles per image

  if (warp == 0) {
    // =========================== loader ===========================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(&w_full, kTaps * kWSlab);
      for (int k = 0; k < kTaps; ++k) tc::bulk_g2s(w_s + k * kWSlab, Wp_ + (size_t)k * kWSlab, kWSlab, &w_full);
      uint32_t it = 0;
      for (int k = 0; k < cnt; ++k) {
        const int t = bx + k * gx;
        const int b = t / tpi, rem = t - b * tpi;
        const int ty = rem / g.tiles_x, tx = rem - ty * g.tiles_x;
        for (int half = 0; half < 2; ++half, ++it) {
          const uint32_t s = it % kSlots;
          tc::mbar_wait(&h_empty[s], ((it / kSlots) & 1u) ^ 1u, err, 4);
          tc::mbar_arrive_expect_tx(&h_full[s], 8 * kPlaneBytes);
          for (int pl = 0; pl < 8; ++pl)
            ic_tma_load(tc::smem_u32(ring + s * kHalf + pl * kPlaneStride), &tmX, tc::smem_u32(&h_full[s]), tx * kTX * 8,
                        ((b * 16 + half * 8 + pl) * g.Hp) + ty * kTY);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t id2 = ic_idesc(128, 2 * kC), id1 = ic_idesc(128, kC);
    const uint32_t r0 = __shfl_sync(0xffffffffu, tc::smem_u32(ring), 0);
    const uint32_t w0 = __shfl_sync(0xffffffffu, tc::smem_u32(w_s), 0);
    const uint32_t td = __shfl_sync(0xffffffffu, tmem_d, 0);
    tc::mbar_wait(&w_full, 0u, err, 1);
    uint32_t it = 0;
    for (int k = 0; k < cnt; ++k) {
      const uint32_t buf = (uint32_t)k & 1u;
      tc::mbar_wait(&acc_free[buf], (((uint32_t)k >> 1) & 1u) ^ 1u, err, 2);          // the epilogue has drained this accumulator
      tc::tc_fence_after_sync();
      const uint32_t d = td + buf * kAccCols;
#pragma unroll 1
      for (int half = 0; half < 2; ++half, ++it) {
        const uint32_t s = it % kSlots;
        tc::mbar_wait(&h_full[s], (it / kSlots) & 1u, err, 3);
        tc::tc_fence_after_sync();
        const uint32_t a0 = r0 + s * kHalf;
        if (tc::elect_one()) {
#pragma unroll
          for (int o = 0; o < kTaps; ++o) {
            const uint32_t shift_b = (uint32_t)(((o / 3) * kPW + (o % 3)) * 16);          // tap (dy, dx) = (o / 3 - 1, o % 3 - 1)
            const uint64_t dw = tc::smem_desc_sw128(w0 + o * kWSlab);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t da = ic_desc_noswz(a0 + shift_b + ks * 2 * kPlaneStride, kPlaneStride, kPW * 16);
              if (half == 0) ic_mma(d, da, dw + (uint64_t)(ks * 2), id2, (o | ks) ? 1u : 0u);          // hi . [Whi ; Wlo]
              else ic_mma(d, da, dw + (uint64_t)(ks * 2), id1, 1u);                                      // lo . Whi
            }
          }
          tc::mma_commit(&h_empty[s]);
          if (half == 1) tc::mma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== epilogue: two groups of four warps, group e takes this CTA's tiles k = e, e + 2, ... ===========================
    const int e = (warp - 2) >> 2, q = warp & 3;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int r = q * 32 + lane;
    const int py = r >> 3, px = r & 7;                           // pixel of the tile this thread owns
    bool big = false;
    for (int k = e; k < cnt; k += 2) {
      const uint32_t buf = (uint32_t)e;
      const int t = bx + k * gx;
      const int b = t / tpi, rem = t - b * tpi;
      const int ty = rem / g.tiles_x, tx = rem - ty * g.tiles_x;
      const int y = ty * kTY + py, x = tx * kTX + px;
      const bool live = y < H && x < W;
      tc::mbar_wait(&acc_full[buf], ((uint32_t)k >> 1) & 1u, err, 5);
      tc::tc_fence_after_sync();
      const uint32_t d = tmem_d + lane_addr + buf * kAccCols;
      // element offset (halves) of this pixel inside a plane; plane p of image b starts at (b * 16 + p) * plane_halves
      const size_t pix = ((size_t)(y + 1) * g.Wp + (x + 1)) * 8;
      const size_t img0 = (size_t)b * 16 * g.plane_halves;
#pragma unroll 1
      for (int cb = 0; cb < kC; cb += 16) {
        uint32_t t1[16], t2[16];
        tc::tmem_ld16_issue(d + (uint32_t)cb, t1);
        tc::tmem_ld16_issue(d + (uint32_t)(kC + cb), t2);
        tc::tmem_ld_wait();
        float a[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(__uint_as_float(t1[i]) + __uint_as_float(t2[i]), sc_s[cb + i], sh_s[cb + i]);
        const int c8 = cb >> 3;
        if (R != nullptr && live) {
          ic_add_chunk(R + img0 + (size_t)c8 * g.plane_halves + pix, R + img0 + (size_t)(8 + c8) * g.plane_halves + pix, a);
          ic_add_chunk(R + img0 + (size_t)(c8 + 1) * g.plane_halves + pix, R + img0 + (size_t)(9 + c8) * g.plane_halves + pix, a + 8);
        }
        if (relu) {
#pragma unroll
          for (int i = 0; i < 16; ++i) a[i] = fmaxf(a[i], 0.f);
        }
        IHalf8 hi[2], lo[2];
        big |= !(ic_split16(a, hi, lo) <= 60000.f) && live;
        if (live) {
          if (pixel_major) {
            __half* yp = Y + ((size_t)b * H * W + (size_t)y * W + x) * ldy + cb;
            tc::st_global_16(yp, hi[0]);
            tc::st_global_16(yp + 8, hi[1]);
            tc::st_global_16(yp + kC, lo[0]);
            tc::st_global_16(yp + kC + 8, lo[1]);
          } else {
            tc::st_global_16(Y + img0 + (size_t)c8 * g.plane_halves + pix, hi[0]);
            tc::st_global_16(Y + img0 + (size_t)(c8 + 1) * g.plane_halves + pix, hi[1]);
            tc::st_global_16(Y + img0 + (size_t)(8 + c8) * g.plane_halves + pix, lo[0]);
            tc::st_global_16(Y + img0 + (size_t)(9 + c8) * g.plane_halves + pix, lo[1]);
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
  if (warp == 1) tc::tmem_dealloc(tmem_d, 256);
}

// K x K / stride / pad max pooling of a pixel-major h2 matrix (C = 64, chunk width kc) into the P8 plane layout; padding never wins.
// One thread = (output pixel, 8 channels).
__global__ void __launch_bounds__(256) k_image_maxpool_p8(const __half* __restrict__ X, int ldx, int kc, int Hin, int Win, int Hout, int Wout,
                                                          int K, int stride, int pad, __half* __restrict__ Y) {
  const P8Geom g = p8_geom(Hout, Wout, 1);
  X += (size_t)blockIdx.y * Hin * Win * ldx;
  Y += (size_t)blockIdx.y * 16 * g.plane_halves;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)Hout * Wout * 8) return;
  const int o = (int)(idx >> 3), c8 = (int)(idx & 7), c0 = c8 * 8;
  const int oy = o / Wout, ox = o - oy * Wout;
  const int off = (c0 / kc) * 2 * kc + (c0 % kc);
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
  for (int ky = 0; ky < K; ++ky) {
    const int iy = oy * stride - pad + ky;
    if (iy < 0 || iy >= Hin) continue;
    for (int kx = 0; kx < K; ++kx) {
      const int ix = ox * stride - pad + kx;
      if (ix < 0 || ix >= Win) continue;
      const __half* p = X + (size_t)(iy * Win + ix) * ldx + off;
      const int4 h4 = *reinterpret_cast<const int4*>(p), l4 = *reinterpret_cast<const int4*>(p + kc);
      const __half* h = reinterpret_cast<const __half*>(&h4);
      const __half* l = reinterpret_cast<const __half*>(&l4);
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], __half2float(h[i]) + __half2float(l[i]));
    }
  }
  __align__(16) __half hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    hi[i] = __float2half_rn(m[i]);
    lo[i] = __float2half_rn(m[i] - __half2float(hi[i]));
  }
  const size_t pix = ((size_t)(oy + 1) * g.Wp + (ox + 1)) * 8;
  *reinterpret_cast<int4*>(Y + (size_t)c8 * g.plane_halves + pix) = *reinterpret_cast<const int4*>(hi);
  *reinterpret_cast<int4*>(Y + (size_t)(8 + c8) * g.plane_halves + pix) = *reinterpret_cast<const int4*>(lo);
}

inline int encode_p8_map(CUtensorMap* map, const void* base, const P8Geom& g) {
  // 2-D view of the planes stacked row-wise: inner = Wp pixels x 8 halves, rows = B * 16 * Hp; box = 10 pixels x 18 rows, no swizzle
  const cuuint64_t gdim[2] = {(cuuint64_t)g.Wp * 8, (cuuint64_t)g.B * 16 * g.Hp};
  const cuuint64_t gstride[1] = {(cuuint64_t)g.Wp * 16};
  const cuuint32_t box[2] = {kPW * 8, kPH};
  const cuuint32_t estr[2] = {1, 1};
  tma::encode_tiled_fn fn = tma::encode_tiled();
  if (!fn) return -1;
  return (int)fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace

// bytes of a 64-channel activation of num_images H x W images in the P8 plane layout (zero-initialise it once: the one-pixel border of
// every plane is never written and must read as zero)
extern "C" size_t imf_image_p8_bytes(int32_t H, int32_t W, int32_t num_images) {
  if (H <= 0 || W <= 0 || num_images <= 0) return 0;
  const P8Geom g = p8_geom(H, W, num_images);
  return (size_t)num_images * 16 * g.plane_halves * sizeof(__half);
}

// 3x3 (ksize) / stride / pad max pooling of num_images pixel-major h2 matrices (64 channels, chunk width kc, ldx halves, image b = rows
// [b * Hin * Win, ...)) into a P8 activation of the pooled size.  /root/reference/model/resnet.py:203 (self.maxpool)
extern "C" int imf_image_maxpool_p8(const void* X, int32_t ldx, int32_t kc, int32_t Hin, int32_t Win, int32_t ksize, int32_t stride, int32_t pad,
                                    void* Y, int32_t num_images, cudaStream_t stream) {
  IMF_CHECK_ARG(X && Y && (kc == 32 || kc == 64) && Hin > 0 && Win > 0 && ksize >= 1 && stride >= 1 && pad >= 0 && pad < ksize);
  IMF_CHECK_ARG(num_images >= 1 && num_images <= 65535 && ldx % 8 == 0 && ldx >= 2 * kC && ((uintptr_t)X % 16) == 0 && ((uintptr_t)Y % 16) == 0);
  const int Hout = (Hin + 2 * pad - ksize) / stride + 1, Wout = (Win + 2 * pad - ksize) / stride + 1;
  IMF_CHECK_ARG(Hout > 0 && Wout > 0);
  const long long total = (long long)Hout * Wout * 8;
  k_image_maxpool_p8<<<dim3((unsigned)((total + 255) / 256), num_images), 256, 0, stream>>>(reinterpret_cast<const __half*>(X), ldx, kc, Hin, Win,
                                                                                          Hout, Wout, ksize, stride, pad,
                                                                                          reinterpret_cast<__half*>(Y));
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

// Y = act( conv3x3/1/pad1(X) * scale + shift (+ R) ) for num_images H x W images of 64 channels.  X, R (optional): P8 activations;
// packed = imf_sparse_conv_h2_pack(W as [9 (tap kx + 3 ky), 64, 64], kc_in 64), multiplier folded into scale; Y: a P8 activation
// (y_pixel_major == 0; its border must already be zero) or a pixel-major h2 matrix of chunk width 64 and ldy halves (y_pixel_major != 0).
extern "C" int imf_image_conv3x3_p8_fwd(const void* X, int32_t H, int32_t W, int32_t num_images, const void* packed, const float* scale,
                                        const float* shift, const void* residual, int32_t relu, void* Y, int32_t y_pixel_major, int32_t ldy,
                                        int32_t* err, cudaStream_t stream) {
  IMF_CHECK_ARG(H > 0 && W > 0 && num_images >= 1 && X != nullptr && packed != nullptr && scale != nullptr && shift != nullptr && Y != nullptr);
  IMF_CHECK_ARG(((uintptr_t)X % 16) == 0 && ((uintptr_t)packed % 16) == 0 && ((uintptr_t)Y % 16) == 0 && ((uintptr_t)residual % 16) == 0);
  IMF_CHECK_ARG(!y_pixel_major || (ldy >= 2 * kC && ldy % 8 == 0));
  const P8Geom g = p8_geom(H, W, num_images);
  IMF_CHECK_ARG((long long)num_images * 16 * g.Hp < (1ll << 31));
  CUtensorMap tmX;
  const int rc = encode_p8_map(&tmX, X, g);
  if (rc) { imf_set_error("cuTensorMapEncodeTiled (P8 activation) failed: %d", rc); return IMF_ERR_CUDA; }
  IMF_CHECK_CUDA(imf_set_max_smem_once(reinterpret_cast<const void*>(&k_image_conv3x3_p8), kSmem));
  const int grid = g.tiles < imf_sm_count() ? g.tiles : imf_sm_count();
  k_image_conv3x3_p8<<<grid, kThreads, kSmem, stream>>>(tmX, reinterpret_cast<const unsigned char*>(packed), H, W, num_images, scale, shift,
                                                        reinterpret_cast<const __half*>(residual), relu, reinterpret_cast<__half*>(Y),
                                                        y_pixel_major, ldy, err);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}
