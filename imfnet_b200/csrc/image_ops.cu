// imfnet_b200 -- image branch on the sparse-convolution kernels.
//
// The ResNet-34 prefix that IMFNet runs on the RGB frame (/root/reference/model/resnet.py:195-216: conv7x7/2 -> BN -> ReLU ->
// maxpool3x3/2 -> layer1 -> layer2, called from model/Img_Encoder.py:15-18) is executed as pixel-major ("token-major")
// h2 matrices [H*W, C] through imf_sparse_conv_g4_fwd: a dense image is the special case of a sparse tensor whose
// neighbour table is known in closed form.  This file provides what is specific to images:
//   * the closed-form offset-major neighbour tables of a k x k / stride s / zero-padded 2-D convolution,
//   * im2col of the 3-channel input for the 7x7 stem (columns (ky, kx, c), padded to a multiple of 32, written as h2),
//   * 3x3/2 max pooling on h2 data.
// The pixel-major output [H/8*W/8, 128] is exactly the [L, dim] token matrix AttentionFusion.forward takes as `data`
// (model/resunet.py:259-261), so the view/permute of the reference costs nothing.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

// nbr_t[k*ld_n + o], k = kx + K*ky: input pixel of output pixel o = oy*Wout + ox at tap (kx, ky), or -1 outside the image
__global__ void __launch_bounds__(128) k_image_conv_table(int Hin, int Win, int Hout, int Wout, int K, int stride, int pad,
                                                          int* __restrict__ nbr_t, int ld_n, unsigned* __restrict__ tile_mask) {
  const int n = Hout * Wout;
  const int tile = blockIdx.x, k = blockIdx.y;
  const int o = tile * 128 + threadIdx.x;
  int r = -1;
  if (o < n) {
    const int oy = o / Wout, ox = o - oy * Wout;
    const int iy = oy * stride - pad + k / K, ix = ox * stride - pad + k % K;
    if (iy >= 0 && iy < Hin && ix >= 0 && ix < Win) r = iy * Win + ix;
  }
  if (o < ld_n) nbr_t[(size_t)k * ld_n + o] = r;
  const unsigned any = __ballot_sync(0xffffffffu, r >= 0);
  if ((threadIdx.x & 31) == 0 && any) atomicOr(tile_mask + tile, 1u << k);
}

// im2col of an NCHW fp32 image for a K x K / stride / pad convolution: row = output pixel, column = c + C*(kx + K*ky),
// zero-padded to Kpad columns, written as an h2 matrix with chunk width 32.  One CTA = 32 consecutive output pixels of one
// image row: the input patch (K rows x (31*stride + K) columns x C channels) is staged in shared memory with coalesced reads,
// then every thread produces 8 columns of one pixel (two 16-byte stores).
__global__ void __launch_bounds__(256) k_image_im2col_h2(const float* __restrict__ img, int C, int H, int W, int Hout, int Wout, int K,
                                                         int stride, int pad, int Kpad, __half* __restrict__ Y, int ldy) {
  extern __shared__ float patch[];                 // [C][K][PW], then the column -> patch offset table [Kpad]
  const int PW = 31 * stride + K;
  int* off_s = reinterpret_cast<int*>(patch + C * K * PW);
  for (int col = threadIdx.x; col < Kpad; col += 256) {
    int o = -1;
    if (col < C * K * K) { const int c = col % C, t = col / C; o = (c * K + t / K) * PW + t % K; }
    off_s[col] = o;
  }
  img += (size_t)blockIdx.y * C * H * W;                    // image blockIdx.y of the batch; its rows follow the previous image's
  Y += (size_t)blockIdx.y * Hout * Wout * ldy;
  const int xblocks = (Wout + 31) / 32;
  const int oy = blockIdx.x / xblocks, ox0 = (blockIdx.x % xblocks) * 32;
  const int iy0 = oy * stride - pad, ix0 = ox0 * stride - pad;
  for (int t = threadIdx.x; t < C * K * PW; t += 256) {
    const int px = t % PW, r = (t / PW) % K, c = t / (PW * K);
    const int iy = iy0 + r, ix = ix0 + px;
    patch[t] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(img + ((size_t)c * H + iy) * W + ix) : 0.f;
  }
  __syncthreads();
  const int groups = Kpad >> 3;
  for (int w = threadIdx.x; w < 32 * groups; w += 256) {
    const int p = w / groups, g = w % groups;
    const int ox = ox0 + p;
    if (ox >= Wout) continue;
    __half hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int o = off_s[g * 8 + i];
      const float v = o >= 0 ? patch[o + p * stride] : 0.f;
      hi[i] = __float2half_rn(v);
      lo[i] = __float2half_rn(v - __half2float(hi[i]));
    }
    const int col0 = g * 8;
    __half* q = Y + (size_t)(oy * Wout + ox) * ldy + (col0 >> 5) * 64 + (col0 & 31);
    *reinterpret_cast<int4*>(q) = *reinterpret_cast<const int4*>(hi);
    *reinterpret_cast<int4*>(q + 32) = *reinterpret_cast<const int4*>(lo);
  }
}

// 3x3 (K x K) max pooling with stride / padding on a pixel-major h2 matrix; padding never wins (torch semantics).
// One thread = (output pixel, 8 channels).
__global__ void __launch_bounds__(256) k_image_maxpool_h2(const __half* __restrict__ X, int ldx, int kc, int C, int Hin, int Win, int Hout,
                                                          int Wout, int K, int stride, int pad, __half* __restrict__ Y, int ldy) {
  const int groups = C >> 3;
  X += (size_t)blockIdx.y * Hin * Win * ldx;                // image blockIdx.y of the batch
  Y += (size_t)blockIdx.y * Hout * Wout * ldy;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)Hout * Wout * groups) return;
  const int o = (int)(idx / groups), c0 = (int)(idx % groups) * 8;
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
  __half hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    hi[i] = __float2half_rn(m[i]);
    lo[i] = __float2half_rn(m[i] - __half2float(hi[i]));
  }
  __half* q = Y + (size_t)o * ldy + off;
  *reinterpret_cast<int4*>(q) = *reinterpret_cast<const int4*>(hi);
  *reinterpret_cast<int4*>(q + kc) = *reinterpret_cast<const int4*>(lo);
}

// [L, C] fp32 row-major -> [C, L] (NCHW feature map) for the ImageEncoder.forward API
__global__ void __launch_bounds__(256) k_transpose_tokens(const float* __restrict__ X, int L, int C, float* __restrict__ Y) {
  __shared__ float t[32][33];
  const int l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int l = l0 + r, c = c0 + tx;
    t[r][tx] = (l < L && c < C) ? X[(size_t)l * C + c] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, l = l0 + tx;
    if (c < C && l < L) Y[(size_t)c * L + l] = t[tx][r];
  }
}

}  // namespace

extern "C" int imf_image_conv_table(int32_t Hin, int32_t Win, int32_t ksize, int32_t stride, int32_t pad, int32_t* nbr_t, int32_t ld_n,
                                    uint32_t* tile_mask, cudaStream_t stream) {
  IMF_CHECK_ARG(Hin > 0 && Win > 0 && ksize >= 1 && ksize <= 5 && (ksize & 1) && stride >= 1 && pad >= 0 && nbr_t && tile_mask);
  const int Hout = (Hin + 2 * pad - ksize) / stride + 1, Wout = (Win + 2 * pad - ksize) / stride + 1;
  IMF_CHECK_ARG(Hout > 0 && Wout > 0 && ksize * ksize <= 27);
  const int n = Hout * Wout, tiles = (n + 127) / 128;
  IMF_CHECK_ARG(ld_n % 4 == 0 && ld_n >= tiles * 128);
  IMF_CHECK_CUDA(cudaMemsetAsync(tile_mask, 0, (size_t)(tiles + 1) * sizeof(uint32_t), stream));
  dim3 grid(tiles, ksize * ksize);
  k_image_conv_table<<<grid, 128, 0, stream>>>(Hin, Win, Hout, Wout, ksize, stride, pad, nbr_t, ld_n, tile_mask);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

extern "C" int imf_image_im2col_h2_batch(const float* image, int32_t C, int32_t H, int32_t W, int32_t ksize, int32_t stride, int32_t pad,
                                         int32_t Kpad, void* Y, int32_t ldy, int32_t num_images, cudaStream_t stream);
extern "C" int imf_image_im2col_h2(const float* image, int32_t C, int32_t H, int32_t W, int32_t ksize, int32_t stride, int32_t pad,
                                   int32_t Kpad, void* Y, int32_t ldy, cudaStream_t stream) {
  return imf_image_im2col_h2_batch(image, C, H, W, ksize, stride, pad, Kpad, Y, ldy, 1, stream);
}
// num_images contiguous NCHW images -> the rows of image b follow those of image b - 1 in Y: one launch for a whole batch
extern "C" int imf_image_im2col_h2_batch(const float* image, int32_t C, int32_t H, int32_t W, int32_t ksize, int32_t stride, int32_t pad,
                                         int32_t Kpad, void* Y, int32_t ldy, int32_t num_images, cudaStream_t stream) {
  IMF_CHECK_ARG(image && Y && C > 0 && H > 0 && W > 0 && ksize >= 1 && stride >= 1 && pad >= 0 && num_images >= 1 && num_images <= 65535);
  IMF_CHECK_ARG(Kpad % 32 == 0 && Kpad >= C * ksize * ksize && ldy % 8 == 0 && ldy >= 2 * Kpad && ((uintptr_t)Y % 16) == 0);
  const int Hout = (H + 2 * pad - ksize) / stride + 1, Wout = (W + 2 * pad - ksize) / stride + 1;
  IMF_CHECK_ARG(Hout > 0 && Wout > 0);
  const size_t smem = (size_t)C * ksize * (31 * stride + ksize) * sizeof(float) + (size_t)Kpad * sizeof(int);
  IMF_CHECK_ARG(smem <= 48 * 1024);
  k_image_im2col_h2<<<dim3((unsigned)(Hout * ((Wout + 31) / 32)), num_images), 256, smem, stream>>>(image, C, H, W, Hout, Wout, ksize, stride,
                                                                                                   pad, Kpad, reinterpret_cast<__half*>(Y), ldy);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

extern "C" int imf_image_maxpool_h2_batch(const void* X, int32_t ldx, int32_t kc, int32_t C, int32_t Hin, int32_t Win, int32_t ksize,
                                          int32_t stride, int32_t pad, void* Y, int32_t ldy, int32_t num_images, cudaStream_t stream);
extern "C" int imf_image_maxpool_h2(const void* X, int32_t ldx, int32_t kc, int32_t C, int32_t Hin, int32_t Win, int32_t ksize,
                                    int32_t stride, int32_t pad, void* Y, int32_t ldy, cudaStream_t stream) {
  return imf_image_maxpool_h2_batch(X, ldx, kc, C, Hin, Win, ksize, stride, pad, Y, ldy, 1, stream);
}
// num_images pixel-major matrices stacked row-wise (image b = rows [b * Hin*Win, ...) of X and [b * Hout*Wout, ...) of Y)
extern "C" int imf_image_maxpool_h2_batch(const void* X, int32_t ldx, int32_t kc, int32_t C, int32_t Hin, int32_t Win, int32_t ksize,
                                          int32_t stride, int32_t pad, void* Y, int32_t ldy, int32_t num_images, cudaStream_t stream) {
  IMF_CHECK_ARG(num_images >= 1 && num_images <= 65535);
  IMF_CHECK_ARG(X && Y && (kc == 32 || kc == 64) && C % kc == 0 && Hin > 0 && Win > 0 && ksize >= 1 && stride >= 1 && pad >= 0 && pad < ksize);
  IMF_CHECK_ARG(ldx % 8 == 0 && ldx >= 2 * C && ldy % 8 == 0 && ldy >= 2 * C && ((uintptr_t)X % 16) == 0 && ((uintptr_t)Y % 16) == 0);
  const int Hout = (Hin + 2 * pad - ksize) / stride + 1, Wout = (Win + 2 * pad - ksize) / stride + 1;
  IMF_CHECK_ARG(Hout > 0 && Wout > 0);
  const long long total = (long long)Hout * Wout * (C / 8);
  k_image_maxpool_h2<<<dim3((unsigned)((total + 255) / 256), num_images), 256, 0, stream>>>(reinterpret_cast<const __half*>(X), ldx, kc, C, Hin,
                                                                                          Win, Hout, Wout, ksize, stride, pad,
                                                                                          reinterpret_cast<__half*>(Y), ldy);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

extern "C" int imf_transpose_tokens(const float* X, int32_t L, int32_t C, float* Y, cudaStream_t stream) {
  IMF_CHECK_ARG(X && Y && L > 0 && C > 0);
  dim3 grid((L + 31) / 32, (C + 31) / 32);
  k_transpose_tokens<<<grid, 256, 0, stream>>>(X, L, C, Y);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}
