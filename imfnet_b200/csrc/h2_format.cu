// imfnet_b200 -- the "h2" activation / weight format of the tensor-core tier (fp16 hi/lo pairs): pack, unpack, weight packing.
//
// Activations live in HBM already split into two fp16 halves (v = hi + lo, 22 mantissa bits, same 4 bytes per channel as fp32), so
// the gather of the sparse convolution (sparse_conv_g4.cu) is a pure asynchronous copy straight into the swizzled shared tiles the
// MMAs read, and every product is accumulated as hi*Whi + hi*Wlo + lo*Whi by kind::f16 MMAs into fp32 TMEM.
//
// h2 matrix (C channels, chunk width KC in {32,64}, C % KC == 0), row stride ld (in halves, >= 2C):
//   channel c = q*KC + j  ->  hi at row*ld + q*2*KC + j,  lo at row*ld + q*2*KC + KC + j
// i.e. per row a sequence of [hi KC | lo KC] chunks: one 128-byte line (KC=32) or two adjacent lines (KC=64) per chunk.
// Serves ME.MinkowskiConvolution(+Transpose) of /root/reference/model/resunet.py:168-213, model/residual_block.py:37-53.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

inline int h2_bn(int Cout) { return Cout > 128 ? 128 : Cout; }      // output-channel tile of the convolution kernel

// Pack W[K3][Cin][Cout] (times wmul, a power of two) into per-(offset, input chunk, BN-wide output tile) slabs of two
// SW128 K-major images (row = output channel, 128 bytes = 64 halves of K):
//   KC = 64: image0 = hi(W), image1 = lo(W), K index = input channel within the chunk;
//   KC = 32: image0 = [hi | hi], image1 = [lo | 0]  (A rows are [hi32 | lo32]).
__global__ void k_pack_conv_weights_h2(const float* __restrict__ W, int K3, int Cin, int Cout, int KC, int BN, float wmul,
                                       __half* __restrict__ Wp) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)K3 * Cin * Cout;
  if (idx >= total) return;
  const int n = (int)(idx % Cout);
  const int ci = (int)((idx / Cout) % Cin);
  const int k = (int)(idx / ((long long)Cout * Cin));
  const float w = W[idx] * wmul;
  const __half hi = __float2half_rn(w);
  const __half lo = __float2half_rn(w - __half2float(hi));
  const int nchunks = Cin / KC, chunk = ci / KC, j = ci % KC;
  const int ntn = Cout / BN, zt = n / BN, nn = n % BN;
  const size_t img = (size_t)BN * 64;                                                  // halves per image
  const size_t slab = (((size_t)k * nchunks + chunk) * ntn + zt) * (2 * img);
  auto pos = [&](int kk) { return (size_t)nn * 64 + (size_t)((((kk >> 3) ^ (nn & 7)) << 3) | (kk & 7)); };
  if (KC == 64) {
    Wp[slab + pos(j)] = hi;
    Wp[slab + img + pos(j)] = lo;
  } else {
    Wp[slab + pos(j)] = hi;
    Wp[slab + pos(j + 32)] = hi;
    Wp[slab + img + pos(j)] = lo;
    Wp[slab + img + pos(j + 32)] = __float2half_rn(0.f);
  }
}

// fp32 [n, C] <-> h2
__global__ void __launch_bounds__(256) k_h2_pack(const float* __restrict__ X, int ldx, int n, int C, int KC, __half* __restrict__ H, int ldh,
                                                 int* err, const int* __restrict__ n_ptr, float mul) {
  if (n_ptr) { const int v = *n_ptr; n = v < n ? v : n; }
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)n * C) return;
  const int row = (int)(idx / C), c = (int)(idx % C);
  const float x = X[(size_t)row * ldx + c] * mul;
  const __half h = __float2half_rn(x);
  __half* p = H + (size_t)row * ldh + (c / KC) * 2 * KC + (c % KC);
  p[0] = h;
  p[KC] = __float2half_rn(x - __half2float(h));
  if (fabsf(x) > 60000.f && err) atomicOr(err, 0x10000);
}
__global__ void __launch_bounds__(256) k_h2_unpack(const __half* __restrict__ H, int ldh, int n, int C, int KC, float* __restrict__ X, int ldx,
                                                   const int* __restrict__ n_ptr, float mul) {
  if (n_ptr) { const int v = *n_ptr; n = v < n ? v : n; }
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)n * C) return;
  const int row = (int)(idx / C), c = (int)(idx % C);
  const __half* p = H + (size_t)row * ldh + (c / KC) * 2 * KC + (c % KC);
  X[(size_t)row * ldx + c] = (__half2float(p[0]) + __half2float(p[KC])) * mul;
}

// h2 [n, C] -> fp32 rows, optionally divided by their L2 norm (model/resunet.py:228-231: F / ||F||_2, no epsilon); one warp per row,
// C <= 128.  out_row (optional) scatters row i to Y[out_row[i]].
__global__ void __launch_bounds__(256) k_h2_unpack_l2norm(const __half* __restrict__ H, int ldh, int n, int C, int KC, int normalize,
                                                          const int* __restrict__ out_row, float* __restrict__ Y, int ldy,
                                                          const int* __restrict__ n_ptr) {
  if (n_ptr) { const int v = *n_ptr; n = v < n ? v : n; }
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  float v[4];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = lane + 32 * i;
    v[i] = 0.f;
    if (c < C) {
      const __half* p = H + (size_t)row * ldh + (c / KC) * 2 * KC + (c % KC);
      v[i] = __half2float(p[0]) + __half2float(p[KC]);
      ss = fmaf(v[i], v[i], ss);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = normalize ? 1.0f / sqrtf(ss) : 1.0f;
  float* y = Y + (size_t)(out_row ? __ldg(out_row + row) : row) * ldy;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = lane + 32 * i;
    if (c < C) y[c] = v[i] * inv;
  }
}

// identity "neighbour table" of a one-offset (1x1) convolution for imf_sparse_conv_g4_fwd: nbr_t[i] = i for i < n, -1 up to the next
// 128-row boundary; tile_mask = 1 for tiles with rows
__global__ void __launch_bounds__(256) k_identity_table(const int* __restrict__ n_ptr, int n_max, int* __restrict__ nbr_t, int ld_n,
                                                        unsigned* __restrict__ tile_mask) {
  int n = n_max;
  if (n_ptr) { const int v = *n_ptr; n = v < n_max ? v : n_max; }
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= ld_n) return;
  nbr_t[i] = i < n ? i : -1;
  if ((i & 127) == 0) tile_mask[i >> 7] = i < n ? 1u : 0u;
}

}  // namespace

extern "C" int imf_h2_unpack_l2norm(const void* H, int32_t ldh, int32_t n, const int32_t* n_dev, int32_t C, int32_t KC, int32_t normalize,
                                    const int32_t* out_row, float* Y, int32_t ldy, cudaStream_t stream) {
  IMF_CHECK_ARG(n >= 0 && C > 0 && C <= 128 && (KC == 32 || KC == 64) && C % KC == 0 && ldy >= C && ldh >= 2 * C);
  if (n == 0) return IMF_OK;
  IMF_CHECK_ARG(H != nullptr && Y != nullptr);
  k_h2_unpack_l2norm<<<(n + 7) / 8, 256, 0, stream>>>(reinterpret_cast<const __half*>(H), ldh, n, C, KC, normalize, out_row, Y, ldy, n_dev);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

extern "C" int imf_identity_table(const int32_t* n_dev, int32_t n_max, int32_t* nbr_t, int32_t ld_n, uint32_t* tile_mask,
                                  cudaStream_t stream) {
  IMF_CHECK_ARG(n_max >= 0 && ld_n % 128 == 0 && ld_n >= n_max);
  if (ld_n == 0) return IMF_OK;
  IMF_CHECK_ARG(nbr_t != nullptr && tile_mask != nullptr);
  k_identity_table<<<(ld_n + 255) / 256, 256, 0, stream>>>(n_dev, n_max, nbr_t, ld_n, tile_mask);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

extern "C" int imf_h2_pack_n(const float* X, int32_t ldx, int32_t n, const int32_t* n_dev, int32_t C, int32_t KC, void* H, int32_t ldh,
                             int32_t* err, cudaStream_t stream);
extern "C" int imf_h2_unpack_n(const void* H, int32_t ldh, int32_t n, const int32_t* n_dev, int32_t C, int32_t KC, float* X, int32_t ldx,
                               cudaStream_t stream);
extern "C" int imf_h2_pack(const float* X, int32_t ldx, int32_t n, int32_t C, int32_t KC, void* H, int32_t ldh, int32_t* err,
                           cudaStream_t stream) {
  return imf_h2_pack_n(X, ldx, n, nullptr, C, KC, H, ldh, err, stream);
}
extern "C" int imf_h2_pack_scaled_n(const float* X, int32_t ldx, int32_t n, const int32_t* n_dev, int32_t C, int32_t KC, float mul, void* H,
                                    int32_t ldh, int32_t* err, cudaStream_t stream);
extern "C" int imf_h2_pack_n(const float* X, int32_t ldx, int32_t n, const int32_t* n_dev, int32_t C, int32_t KC, void* H, int32_t ldh,
                             int32_t* err, cudaStream_t stream) {
  return imf_h2_pack_scaled_n(X, ldx, n, n_dev, C, KC, 1.f, H, ldh, err, stream);
}
// H = h2(X * mul): mul is the (power-of-two) activation scale of the tensor-core tier (stored activations = true values * scale)
extern "C" int imf_h2_pack_scaled_n(const float* X, int32_t ldx, int32_t n, const int32_t* n_dev, int32_t C, int32_t KC, float mul, void* H,
                                    int32_t ldh, int32_t* err, cudaStream_t stream) {
  IMF_CHECK_ARG(n >= 0 && C > 0 && (KC == 32 || KC == 64) && C % KC == 0 && ldx >= C && ldh >= 2 * C);
  if (n == 0) return IMF_OK;
  IMF_CHECK_ARG(X != nullptr && H != nullptr);
  const long long total = (long long)n * C;
  k_h2_pack<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(X, ldx, n, C, KC, reinterpret_cast<__half*>(H), ldh, err, n_dev, mul);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

extern "C" int imf_h2_unpack(const void* H, int32_t ldh, int32_t n, int32_t C, int32_t KC, float* X, int32_t ldx, cudaStream_t stream) {
  return imf_h2_unpack_n(H, ldh, n, nullptr, C, KC, X, ldx, stream);
}
extern "C" int imf_h2_unpack_scaled_n(const void* H, int32_t ldh, int32_t n, const int32_t* n_dev, int32_t C, int32_t KC, float mul, float* X,
                                      int32_t ldx, cudaStream_t stream);
extern "C" int imf_h2_unpack_n(const void* H, int32_t ldh, int32_t n, const int32_t* n_dev, int32_t C, int32_t KC, float* X, int32_t ldx,
                               cudaStream_t stream) {
  return imf_h2_unpack_scaled_n(H, ldh, n, n_dev, C, KC, 1.f, X, ldx, stream);
}
// X = (hi + lo) * mul
extern "C" int imf_h2_unpack_scaled_n(const void* H, int32_t ldh, int32_t n, const int32_t* n_dev, int32_t C, int32_t KC, float mul, float* X,
                                      int32_t ldx, cudaStream_t stream) {
  IMF_CHECK_ARG(n >= 0 && C > 0 && (KC == 32 || KC == 64) && C % KC == 0 && ldx >= C && ldh >= 2 * C);
  if (n == 0) return IMF_OK;
  IMF_CHECK_ARG(X != nullptr && H != nullptr);
  const long long total = (long long)n * C;
  k_h2_unpack<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const __half*>(H), ldh, n, C, KC, X, ldx, n_dev, mul);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

extern "C" size_t imf_sparse_conv_h2_packed_bytes(int32_t kernel_volume, int32_t Cin, int32_t Cout, int32_t kc_in) {
  // KC=64: hi + lo images of 64 K-halves; KC=32: two 64-wide images per 32-channel chunk
  const size_t per_k = (size_t)(Cin / kc_in) * Cout * 64 * 2 * sizeof(__half);
  return (size_t)kernel_volume * per_k;
}

extern "C" int imf_sparse_conv_h2_pack(const float* W, int32_t kernel_volume, int32_t Cin, int32_t Cout, int32_t kc_in, float wmul,
                                       void* packed, cudaStream_t stream) {
  IMF_CHECK_ARG(W != nullptr && packed != nullptr && kernel_volume >= 1 && (kc_in == 32 || kc_in == 64) && Cin > 0 && Cin % kc_in == 0);
  IMF_CHECK_ARG(Cout == 32 || Cout == 64 || (Cout >= 128 && Cout % 128 == 0));          // (tiles of min(Cout, 128) output channels)
  IMF_CHECK_ARG(wmul > 0.f);
  const long long total = (long long)kernel_volume * Cin * Cout;
  k_pack_conv_weights_h2<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(W, kernel_volume, Cin, Cout, kc_in, h2_bn(Cout), wmul,
                                                                            reinterpret_cast<__half*>(packed));
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}
