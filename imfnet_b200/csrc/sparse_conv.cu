// imfnet_b200 -- sparse 3-D convolution, output-stationary gather + implicit GEMM (fp32 SIMT tier).
//
// Replaces ME.MinkowskiConvolution / MinkowskiConvolutionTranspose forward for the IMFNet descriptor path
//   call sites: /root/reference/model/resunet.py:168,173,178,183,191,202,213; model/residual_block.py:40,44
// with the eval-mode MinkowskiBatchNorm (model/common.py:6), the residual add and the ReLU
// (model/residual_block.py:41-51) fused into the epilogue:
//       Y[o] = act( (sum_k X[nbr[o,k]] . W[k]) * scale + shift (+ R[o]) )
// Offsets are accumulated in ascending k and channels in ascending order by ONE thread per output element,
// so results are deterministic and independent of the internal row order (no scatter atomics).
//
// Kernel shape: CTA = 256 threads = 8 warps, tile = (8*RM) output rows x (32*TN) output channels.
// A warp owns RM rows and all 32*TN channels of the tile (lane <-> channel), therefore "is neighbour k
// present for row r" is warp-uniform and absent neighbours cost nothing.  Stages of (offset k, 32 input
// channels) are streamed with cp.async into a 3-deep shared-memory ring: gathered rows (16-byte vector
// copies, one 128 B line per row and stage) and the matching weight slab.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int kCK = 32;        // input channels per stage (one 128 B line per gathered row)
constexpr int kStages = 3;
constexpr int kMaxK3 = 27;     // this kernel covers 3x3x3 (and 1x1x1) maps; 5^3 with Cin<=8 has its own kernel

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int RM, int TN>
__global__ void __launch_bounds__(256) k_sparse_conv(const float* __restrict__ X, int ldx, const float* __restrict__ W,
                                                     const int* __restrict__ nbr, const int* __restrict__ n_ptr, int n_max,
                                                     int K3, int Cin, int Cout, const float* __restrict__ scale,
                                                     const float* __restrict__ shift, const float* __restrict__ R, int ldr,
                                                     int relu, float* __restrict__ Y, int ldy) {
  constexpr int BM = 8 * RM, BN = 32 * TN;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* A_s = reinterpret_cast<float*>(smem_raw);                         // [kStages][BM][kCK]
  float* W_s = A_s + kStages * BM * kCK;                                   // [kStages][kCK][BN]
  int* nbr_s = reinterpret_cast<int*>(W_s + kStages * kCK * BN);           // [BM][K3]
  __shared__ unsigned kmask_s;
  __shared__ int klist_s[32];
  __shared__ int nk_s;

  int n_out = n_max;
  if (n_ptr) { int v = *n_ptr; n_out = v < n_max ? v : n_max; }
  const int row0 = blockIdx.x * BM;
  if (row0 >= n_out) return;
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) kmask_s = 0u;
  __syncthreads();
  {
    unsigned local = 0u;
    const int total = BM * K3;
    const int* src = nbr + (size_t)row0 * K3;
    const int valid = (n_out - row0 < BM ? n_out - row0 : BM) * K3;
    for (int idx = tid; idx < total; idx += 256) {
      int v = -1;
      if (idx < valid) v = __ldg(src + idx);
      nbr_s[idx] = v;
      if (v >= 0) local |= 1u << (idx % K3);
    }
    local = __reduce_or_sync(0xffffffffu, local);
    if (lane == 0 && local) atomicOr(&kmask_s, local);
  }
  __syncthreads();
  if (tid == 0) {
    unsigned m = kmask_s;
    int c = 0;
    while (m) { int b = __ffs(m) - 1; m &= m - 1; klist_s[c++] = b; }
    nk_s = c;
  }
  __syncthreads();
  const int nk = nk_s;
  const int nchunks = Cin / kCK;
  const int nst = nk * nchunks;

  float acc[RM][TN];
#pragma unroll
  for (int r = 0; r < RM; ++r)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[r][j] = 0.f;

  auto load_stage = [&](int s) {
    const int buf = s % kStages;
    const int k = klist_s[s / nchunks];
    const int c0 = (s % nchunks) * kCK;
    float* a_dst = A_s + buf * BM * kCK;
    for (int c = tid; c < BM * (kCK / 4); c += 256) {
      const int row = c >> 3, ch = c & 7;
      const int n = nbr_s[row * K3 + k];
      if (n >= 0) cp_async16(a_dst + row * kCK + ch * 4, X + (size_t)n * ldx + c0 + ch * 4);
    }
    float* w_dst = W_s + buf * kCK * BN;
    const float* w_src = W + ((size_t)k * Cin + c0) * Cout + n0;
    for (int c = tid; c < kCK * (BN / 4); c += 256) {
      const int r = c / (BN / 4), ch = c % (BN / 4);
      cp_async16(w_dst + r * BN + ch * 4, w_src + (size_t)r * Cout + ch * 4);
    }
  };

#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < nst) load_stage(s);
    cp_async_commit();
  }

  for (int s = 0; s < nst; ++s) {
    cp_async_wait<kStages - 2>();
    __syncthreads();
    if (s + kStages - 1 < nst) load_stage(s + kStages - 1);
    cp_async_commit();

    const int buf = s % kStages;
    const int k = klist_s[s / nchunks];
    const bool present = (lane < RM) && (nbr_s[(warp * RM + lane) * K3 + k] >= 0);
    const unsigned m = __ballot_sync(0xffffffffu, present);
    if (m == 0u) continue;
    const float* a_base = A_s + buf * BM * kCK + warp * RM * kCK;
    const float* w_base = W_s + buf * kCK * BN + lane;
#pragma unroll 1
    for (int cc = 0; cc < kCK; cc += 8) {
      float w[8][TN];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) w[i][j] = w_base[(cc + i) * BN + j * 32];
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        if ((m >> r) & 1u) {
          const float4 a0 = *reinterpret_cast<const float4*>(a_base + r * kCK + cc);
          const float4 a1 = *reinterpret_cast<const float4*>(a_base + r * kCK + cc + 4);
#pragma unroll
          for (int j = 0; j < TN; ++j) {
            float v = acc[r][j];
            v = fmaf(a0.x, w[0][j], v);
            v = fmaf(a0.y, w[1][j], v);
            v = fmaf(a0.z, w[2][j], v);
            v = fmaf(a0.w, w[3][j], v);
            v = fmaf(a1.x, w[4][j], v);
            v = fmaf(a1.y, w[5][j], v);
            v = fmaf(a1.z, w[6][j], v);
            v = fmaf(a1.w, w[7][j], v);
            acc[r][j] = v;
          }
        }
      }
    }
  }
  cp_async_wait<0>();

  // epilogue: folded BatchNorm affine, residual, ReLU; 32 lanes write 128 contiguous bytes per row
#pragma unroll
  for (int r = 0; r < RM; ++r) {
    const int row = row0 + warp * RM + r;
    if (row < n_out) {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int c = n0 + j * 32 + lane;
        float v = acc[r][j];
        if (scale) v = fmaf(v, __ldg(scale + c), __ldg(shift + c));
        if (R) v += R[(size_t)row * ldr + c];
        if (relu) v = fmaxf(v, 0.f);
        Y[(size_t)row * ldy + c] = v;
      }
    }
  }
}

template <int RM, int TN>
int launch_conv(const float* X, int ldx, const float* W, const int* nbr, const int* n_ptr, int n_max, int K3, int Cin,
                int Cout, const float* scale, const float* shift, const float* R, int ldr, int relu, float* Y, int ldy,
                cudaStream_t stream) {
  constexpr int BM = 8 * RM, BN = 32 * TN;
  const size_t smem = (size_t)kStages * (BM * kCK + kCK * BN) * sizeof(float) + (size_t)BM * K3 * sizeof(int);
  IMF_CHECK_CUDA(cudaFuncSetAttribute(k_sparse_conv<RM, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  dim3 grid((n_max + BM - 1) / BM, Cout / BN);
  k_sparse_conv<RM, TN><<<grid, 256, smem, stream>>>(X, ldx, W, nbr, n_ptr, n_max, K3, Cin, Cout, scale, shift, R, ldr,
                                                     relu, Y, ldy);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

// ------------------------------------------------------------------------------------------------
// First layer: K^3 up to 125 offsets, Cin <= 8 (1 for IMFNet: a column of ones, util/misc.py:76-77).
// A warp owns one output voxel at a time: lanes probe the hash table for the K^3 neighbours (no neighbour table is
// materialised -- 125 int32 per voxel would be 25 MB at 50 k voxels), then lane <-> output channel accumulates
// the present offsets in ascending k.  Weights (K^3*Cin*Cout fp32, 16 KB for 125x1x32) sit in shared memory.
template <int TN, int CIN, bool H2>
__global__ void __launch_bounds__(256) k_conv_first(const float* __restrict__ X, int ldx, const float* __restrict__ W,
                                                    const int4* __restrict__ coords, const int* __restrict__ n_ptr, int n_max,
                                                    const ImfSlot* __restrict__ table, unsigned long long mask, int K,
                                                    int tstride, const float* __restrict__ scale,
                                                    const float* __restrict__ shift, int relu, float* __restrict__ Y, int ldy,
                                                    int rows_per_cta, int kc_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* W_s = reinterpret_cast<float*>(smem_raw);
  constexpr int Cout = 32 * TN;
  const int K3 = K * K * K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int n = n_max;
  if (n_ptr) { int v = *n_ptr; n = v < n_max ? v : n_max; }
  const int row_begin = blockIdx.x * rows_per_cta;
  if (row_begin >= n) return;
  for (int i = tid; i < K3 * CIN * Cout; i += 256) W_s[i] = __ldg(W + i);
  __syncthreads();
  const int row_end = min(n, row_begin + rows_per_cta);
  const int h = K / 2;
  // per-lane offsets of the (up to) 4 probe rounds, computed once
  int ox[4], oy[4], oz[4];
#pragma unroll
  for (int rd = 0; rd < 4; ++rd) {
    const int k = rd * 32 + lane;
    ox[rd] = (k % K - h) * tstride;
    oy[rd] = ((k / K) % K - h) * tstride;
    oz[rd] = (k / (K * K) - h) * tstride;
  }
  for (int row = row_begin + warp; row < row_end; row += 8) {
    const int4 c = __ldg(coords + row);
    int found[4];
    float xv[4][CIN];
    // phase 1: all probes and all feature loads of the neighbourhood are independent -> issued back to back
#pragma unroll
    for (int rd = 0; rd < 4; ++rd) {
      const int k = rd * 32 + lane;
      int r = -1;
      if (k < K3) {
        const int x = c.y + ox[rd], y = c.z + oy[rd], z = c.w + oz[rd];
        if (imf_coord_in_range(c.x, x, y, z)) r = imf_table_lookup(table, mask, imf_pack_key(c.x, x, y, z));
      }
      found[rd] = r;
    }
#pragma unroll
    for (int rd = 0; rd < 4; ++rd) {
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) xv[rd][ci] = (found[rd] >= 0) ? __ldg(X + (size_t)found[rd] * ldx + ci) : 0.f;
    }
    // phase 2: lane <-> output channel; present offsets accumulated in ascending k
    float acc[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[j] = 0.f;
#pragma unroll
    for (int rd = 0; rd < 4; ++rd) {
      unsigned m = __ballot_sync(0xffffffffu, found[rd] >= 0);
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const float* wr = W_s + (size_t)(rd * 32 + b) * CIN * Cout + lane;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float x = __shfl_sync(0xffffffffu, xv[rd][ci], b);
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[j] = fmaf(x, wr[ci * Cout + j * 32], acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int cch = j * 32 + lane;
      float v = acc[j];
      if (scale) v = fmaf(v, __ldg(scale + cch), __ldg(shift + cch));
      if (relu) v = fmaxf(v, 0.f);
      if (H2) {   // fp16 hi/lo output for the tensor-core tier (layout: sparse_conv_h2.cu); ldy counts halves
        __half* yp = reinterpret_cast<__half*>(Y) + (size_t)row * ldy + (cch / kc_out) * 2 * kc_out + (cch % kc_out);
        const __half h = __float2half_rn(v);
        yp[0] = h;
        yp[kc_out] = __float2half_rn(v - __half2float(h));
      } else {
        Y[(size_t)row * ldy + cch] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Tail: conv1_tr (1x1, C0->C1, no bias) -> ReLU -> final (1x1, C1->C2, bias) -> row-wise L2 normalisation
//   /root/reference/model/resunet.py:224-233.  One warp owns RM rows; lane <-> channel; hidden row stays in smem.
template <int TN1, bool H2>
__global__ void __launch_bounds__(256) k_pointwise_tail(const float* __restrict__ X, int ldx, int C0, const float* __restrict__ W1,
                                                        const float* __restrict__ W2, const float* __restrict__ b2, int C2,
                                                        const int* __restrict__ n_ptr, int n_max, int normalize,
                                                        float* __restrict__ Y, int ldy, int Ca, int kca, int kcb,
                                                        const int* __restrict__ out_row) {
  constexpr int RM = 8, C1 = 32 * TN1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* W1_s = reinterpret_cast<float*>(smem_raw);       // [C0][C1]
  float* W2_s = W1_s + C0 * C1;                            // [C1][32]  (C2 <= 32, zero padded)
  float* X_s = W2_s + C1 * 32;                             // [64][C0]
  float* H_s = X_s + 64 * C0;                              // [64][C1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int n = n_max;
  if (n_ptr) { int v = *n_ptr; n = v < n_max ? v : n_max; }
  const int row0 = blockIdx.x * 64;
  if (row0 >= n) return;
  for (int i = tid; i < C0 * C1; i += 256) W1_s[i] = __ldg(W1 + i);
  for (int i = tid; i < C1 * 32; i += 256) {
    const int r = i >> 5, c = i & 31;
    W2_s[i] = (c < C2) ? __ldg(W2 + r * C2 + c) : 0.f;
  }
  const int rows = min(64, n - row0);
  if (H2) {
    // X is an h2 matrix of two sections: channels [0,Ca) with chunk width kca, then [Ca,C0) with chunk width kcb
    const __half* Xh = reinterpret_cast<const __half*>(X);
    for (int i = tid; i < 64 * (C0 / 8); i += 256) {
      const int r = i / (C0 / 8), c = (i % (C0 / 8)) * 8;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      if (r < rows) {
        int off, kc;
        if (c < Ca) { kc = kca; off = (c / kca) * 2 * kca + (c % kca); }
        else { kc = kcb; const int c2 = c - Ca; off = 2 * Ca + (c2 / kcb) * 2 * kcb + (c2 % kcb); }
        const __half* p = Xh + (size_t)(row0 + r) * ldx + off;
        const uint4 hq = *reinterpret_cast<const uint4*>(p), lq = *reinterpret_cast<const uint4*>(p + kc);
        const __half2* hh = reinterpret_cast<const __half2*>(&hq);
        const __half2* ll = reinterpret_cast<const __half2*>(&lq);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = __half22float2(hh[e]), b = __half22float2(ll[e]);
          v[2 * e] = a.x + b.x;
          v[2 * e + 1] = a.y + b.y;
        }
      }
      *reinterpret_cast<float4*>(X_s + r * C0 + c) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(X_s + r * C0 + c + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
  } else {
    for (int i = tid; i < 64 * (C0 / 4); i += 256) {
      const int r = i / (C0 / 4), ch = i % (C0 / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows) v = *reinterpret_cast<const float4*>(X + (size_t)(row0 + r) * ldx + ch * 4);
      *reinterpret_cast<float4*>(X_s + r * C0 + ch * 4) = v;
    }
  }
  __syncthreads();

  float acc[RM][TN1];
#pragma unroll
  for (int r = 0; r < RM; ++r)
#pragma unroll
    for (int j = 0; j < TN1; ++j) acc[r][j] = 0.f;
  const float* xw = X_s + warp * RM * C0;
  for (int cc = 0; cc < C0; cc += 4) {
    float w[4][TN1];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < TN1; ++j) w[i][j] = W1_s[(cc + i) * C1 + j * 32 + lane];
#pragma unroll
    for (int r = 0; r < RM; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(xw + r * C0 + cc);
#pragma unroll
      for (int j = 0; j < TN1; ++j) {
        float v = acc[r][j];
        v = fmaf(a.x, w[0][j], v);
        v = fmaf(a.y, w[1][j], v);
        v = fmaf(a.z, w[2][j], v);
        v = fmaf(a.w, w[3][j], v);
        acc[r][j] = v;
      }
    }
  }
  float* hw = H_s + warp * RM * C1;
#pragma unroll
  for (int r = 0; r < RM; ++r)
#pragma unroll
    for (int j = 0; j < TN1; ++j) hw[r * C1 + j * 32 + lane] = fmaxf(acc[r][j], 0.f);
  __syncwarp();

  float o[RM];
#pragma unroll
  for (int r = 0; r < RM; ++r) o[r] = 0.f;
  for (int cc = 0; cc < C1; cc += 4) {
    const float w0 = W2_s[(cc + 0) * 32 + lane], w1 = W2_s[(cc + 1) * 32 + lane], w2 = W2_s[(cc + 2) * 32 + lane],
                w3 = W2_s[(cc + 3) * 32 + lane];
#pragma unroll
    for (int r = 0; r < RM; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(hw + r * C1 + cc);
      float v = o[r];
      v = fmaf(a.x, w0, v);
      v = fmaf(a.y, w1, v);
      v = fmaf(a.z, w2, v);
      v = fmaf(a.w, w3, v);
      o[r] = v;
    }
  }
  const float bias = (b2 != nullptr && lane < C2) ? __ldg(b2 + lane) : 0.f;
#pragma unroll
  for (int r = 0; r < RM; ++r) {
    const int row = row0 + warp * RM + r;
    float v = (lane < C2) ? o[r] + bias : 0.f;
    if (normalize) {
      const float ss = imf_warp_sum(v * v);
      v = v / sqrtf(ss);
    }
    if (row < n && lane < C2) {
      const int orow = out_row ? __ldg(out_row + row) : row;     // optional scatter back to the caller's row order
      Y[(size_t)orow * ldy + lane] = v;
    }
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" int imf_sparse_conv_fwd(const float* X, int32_t ldx, const float* W, const int32_t* nbr, const int32_t* n_out_dev,
                                   int32_t n_out_max, int32_t kernel_volume, int32_t Cin, int32_t Cout, const float* scale,
                                   const float* shift, const float* residual, int32_t ldr, int32_t relu, float* Y,
                                   int32_t ldy, cudaStream_t stream) {
  IMF_CHECK_ARG(n_out_max >= 0 && kernel_volume >= 1 && kernel_volume <= kMaxK3);
  IMF_CHECK_ARG(Cin > 0 && Cin % kCK == 0 && Cout > 0 && Cout % 32 == 0);
  IMF_CHECK_ARG((scale == nullptr) == (shift == nullptr));
  IMF_CHECK_ARG(ldx % 4 == 0 && ldx >= Cin && ldy >= Cout && (residual == nullptr || ldr >= Cout));
  if (n_out_max == 0) return IMF_OK;
  IMF_CHECK_ARG(X != nullptr && W != nullptr && nbr != nullptr && Y != nullptr);
  IMF_CHECK_ARG(((uintptr_t)X % 16) == 0 && ((uintptr_t)W % 16) == 0);
#define IMF_GO(RM, TN)                                                                                                   \
  return launch_conv<RM, TN>(X, ldx, W, nbr, n_out_dev, n_out_max, kernel_volume, Cin, Cout, scale, shift, residual, ldr, \
                             relu, Y, ldy, stream)
  // Tile choice: widest channel tile that divides Cout; shrink the row tile while the grid would leave SMs idle.
  const long long n = n_out_max;
  const int sms = imf_sm_count();
  if (Cout % 128 == 0) {
    if ((n + 63) / 64 * (Cout / 128) >= sms) IMF_GO(8, 4);
    if ((n + 31) / 32 * (Cout / 64) >= sms || n <= 32) IMF_GO(4, 2);
    IMF_GO(2, 2);
  }
  if (Cout % 64 == 0) {
    if ((n + 127) / 128 * (Cout / 64) >= sms) IMF_GO(16, 2);
    if ((n + 31) / 32 * (Cout / 64) >= sms || n <= 32) IMF_GO(4, 2);
    IMF_GO(2, 2);
  }
  if ((n + 127) / 128 * (Cout / 32) >= sms) IMF_GO(16, 1);
  IMF_GO(4, 1);
#undef IMF_GO
}

template <int TN, int CIN, bool H2>
static int launch_first(const float* X, int ldx, const float* W, const int32_t* coords, const int32_t* n_dev, int n_max, const void* table,
                        long long capacity, int K, int tstride, const float* scale, const float* shift, int relu, float* Y, int ldy,
                        int kc_out, cudaStream_t stream) {
  const int K3 = K * K * K;
  const size_t smem = (size_t)K3 * CIN * 32 * TN * sizeof(float);
  IMF_CHECK_ARG(smem <= 200 * 1024);
  const int rows_per_cta = 64;
  const int grid = (n_max + rows_per_cta - 1) / rows_per_cta;
  IMF_CHECK_CUDA(cudaFuncSetAttribute(k_conv_first<TN, CIN, H2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  k_conv_first<TN, CIN, H2><<<grid, 256, smem, stream>>>(X, ldx, W, reinterpret_cast<const int4*>(coords), n_dev, n_max,
                                                         reinterpret_cast<const ImfSlot*>(table), (unsigned long long)capacity - 1, K,
                                                         tstride, scale, shift, relu, Y, ldy, rows_per_cta, kc_out);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

template <bool H2>
static int conv_first_dispatch(const float* X, int32_t ldx, int32_t Cin, const float* W, const int32_t* coords, const int32_t* n_dev,
                               int32_t n_max, const void* table, long long capacity, int32_t kernel_size, int32_t tensor_stride,
                               int32_t Cout, const float* scale, const float* shift, int32_t relu, float* Y, int32_t ldy, int32_t kc_out,
                               cudaStream_t stream) {
  IMF_CHECK_ARG(n_max >= 0 && kernel_size >= 1 && (kernel_size & 1) && kernel_size <= 5);
  IMF_CHECK_ARG((Cin == 1 || Cin == 3 || Cin == 6) && (Cout == 32 || Cout == 64 || Cout == 128));
  IMF_CHECK_ARG((scale == nullptr) == (shift == nullptr) && ldx >= Cin && ldy >= (H2 ? 2 : 1) * Cout);
  IMF_CHECK_ARG(!H2 || ((kc_out == 32 || kc_out == 64) && Cout % kc_out == 0));
  IMF_CHECK_ARG(capacity > 0 && (capacity & (capacity - 1)) == 0);
  if (n_max == 0) return IMF_OK;
  IMF_CHECK_ARG(X != nullptr && W != nullptr && coords != nullptr && table != nullptr && Y != nullptr);
#define IMF_GO(TN, CIN)                                                                                                       \
  return launch_first<TN, CIN, H2>(X, ldx, W, coords, n_dev, n_max, table, capacity, kernel_size, tensor_stride, scale, shift, relu, \
                                   Y, ldy, kc_out, stream)
#define IMF_GO_C(TN)            \
  do {                          \
    if (Cin == 1) IMF_GO(TN, 1); \
    if (Cin == 3) IMF_GO(TN, 3); \
    IMF_GO(TN, 6);              \
  } while (0)
  if (Cout == 32) IMF_GO_C(1);
  if (Cout == 64) IMF_GO_C(2);
  IMF_GO_C(4);
#undef IMF_GO_C
#undef IMF_GO
}

extern "C" int imf_conv_first_fwd(const float* X, int32_t ldx, int32_t Cin, const float* W, const int32_t* coords,
                                  const int32_t* n_dev, int32_t n_max, const void* table, long long capacity,
                                  int32_t kernel_size, int32_t tensor_stride, int32_t Cout, const float* scale,
                                  const float* shift, int32_t relu, float* Y, int32_t ldy, cudaStream_t stream) {
  return conv_first_dispatch<false>(X, ldx, Cin, W, coords, n_dev, n_max, table, capacity, kernel_size, tensor_stride, Cout, scale,
                                    shift, relu, Y, ldy, 0, stream);
}

// Same, writing the fp16 hi/lo ("h2", sparse_conv_h2.cu) layout; ldy counts halves.
extern "C" int imf_conv_first_h2_fwd(const float* X, int32_t ldx, int32_t Cin, const float* W, const int32_t* coords,
                                     const int32_t* n_dev, int32_t n_max, const void* table, long long capacity,
                                     int32_t kernel_size, int32_t tensor_stride, int32_t Cout, const float* scale,
                                     const float* shift, int32_t relu, void* Y, int32_t ldy, int32_t kc_out, cudaStream_t stream) {
  return conv_first_dispatch<true>(X, ldx, Cin, W, coords, n_dev, n_max, table, capacity, kernel_size, tensor_stride, Cout, scale,
                                   shift, relu, reinterpret_cast<float*>(Y), ldy, kc_out, stream);
}

template <bool H2>
static int tail_dispatch(const float* X, int32_t ldx, int32_t C0, const float* W1, int32_t C1, const float* W2, const float* b2,
                         int32_t C2, const int32_t* n_dev, int32_t n_max, int32_t normalize, float* Y, int32_t ldy, int32_t Ca,
                         int32_t kca, int32_t kcb, const int32_t* out_row, cudaStream_t stream) {
  IMF_CHECK_ARG(n_max >= 0 && C0 > 0 && C0 % (H2 ? 8 : 4) == 0 && ldx % (H2 ? 8 : 4) == 0 && ldx >= (H2 ? 2 : 1) * C0);
  IMF_CHECK_ARG(C2 >= 1 && C2 <= 32 && ldy >= C2);
  IMF_CHECK_ARG(C1 == 32 || C1 == 64 || C1 == 128);
  if (H2) {
    IMF_CHECK_ARG(Ca >= 0 && Ca <= C0 && (kca == 32 || kca == 64) && (kcb == 32 || kcb == 64) && Ca % kca == 0 && (C0 - Ca) % kcb == 0);
  }
  if (n_max == 0) return IMF_OK;
  IMF_CHECK_ARG(X != nullptr && W1 != nullptr && W2 != nullptr && Y != nullptr && ((uintptr_t)X % 16) == 0);
  const size_t smem = ((size_t)C0 * C1 + (size_t)C1 * 32 + 64 * (size_t)C0 + 64 * (size_t)C1) * sizeof(float);
  IMF_CHECK_ARG(smem <= 200 * 1024);
  const int grid = (n_max + 63) / 64;
#define IMF_GO(TN1)                                                                                                         \
  do {                                                                                                                      \
    IMF_CHECK_CUDA(cudaFuncSetAttribute(k_pointwise_tail<TN1, H2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
    k_pointwise_tail<TN1, H2><<<grid, 256, smem, stream>>>(X, ldx, C0, W1, W2, b2, C2, n_dev, n_max, normalize, Y, ldy, Ca, kca, kcb, \
                                                           out_row);                                                         \
  } while (0)
  if (C1 == 32) IMF_GO(1); else if (C1 == 64) IMF_GO(2); else IMF_GO(4);
#undef IMF_GO
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

extern "C" int imf_pointwise_tail_fwd(const float* X, int32_t ldx, int32_t C0, const float* W1, int32_t C1, const float* W2,
                                      const float* b2, int32_t C2, const int32_t* n_dev, int32_t n_max, int32_t normalize,
                                      float* Y, int32_t ldy, cudaStream_t stream) {
  return tail_dispatch<false>(X, ldx, C0, W1, C1, W2, b2, C2, n_dev, n_max, normalize, Y, ldy, 0, 0, 0, nullptr, stream);
}

// Same on an h2 input made of two sections (channels [0,Ca) with chunk width kca, [Ca,C0) with kcb; ldx counts halves);
// out_row (optional, int32 [n]) scatters result row i to Y[out_row[i]] (internal -> caller row order).
extern "C" int imf_pointwise_tail_h2_fwd(const void* X, int32_t ldx, int32_t C0, int32_t Ca, int32_t kca, int32_t kcb, const float* W1,
                                         int32_t C1, const float* W2, const float* b2, int32_t C2, const int32_t* n_dev, int32_t n_max,
                                         int32_t normalize, const int32_t* out_row, float* Y, int32_t ldy, cudaStream_t stream) {
  return tail_dispatch<true>(reinterpret_cast<const float*>(X), ldx, C0, W1, C1, W2, b2, C2, n_dev, n_max, normalize, Y, ldy, Ca, kca,
                             kcb, out_row, stream);
}
