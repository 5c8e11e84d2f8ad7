// imfnet_b200 -- sparse 3-D convolution as an implicit GEMM on the tcgen05 tensor cores (3xTF32, TMEM accumulators).
//
// Same contract as imf_sparse_conv_fwd (sparse_conv.cu): output-stationary gather, ascending-k accumulation,
//   Y[o] = act( (sum_k X[nbr[o,k]] . W[k]) * scale + shift (+ R[o]) )
// for ME.MinkowskiConvolution / MinkowskiConvolutionTranspose + MinkowskiBatchNorm + ReLU / residual
//   /root/reference/model/resunet.py:168-213, model/residual_block.py:37-53.
//
// CTA = 288 threads, tile = 128 output rows x Cout (whole channel width, so every gathered row is read once):
//   warps 0-7  gather producers: neighbour row -> registers (16-byte vectors, one 128 B line per row and stage) -> hi/lo
//              TF32 split -> SW128 K-major shared tiles; absent neighbours become zero rows.  One of them also issues the
//              bulk (TMA) copy of the stage's pre-packed, pre-split, pre-swizzled weight slab.  Afterwards: epilogue
//              TMEM -> registers -> BatchNorm affine / residual / ReLU -> global.
//   warp  8    TMEM allocation + tcgen05.mma issue: per stage (offset k, 32 input channels) 4 K-slices x 3 split products.
// (offset, channel-chunk) stages whose offset has no neighbour in the whole tile are skipped.  For the small deep levels
// the stage list of a tile is split over several CTAs (grid.y); partial tiles are summed by k_conv_splitk_epilogue.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kBM = 128;

template <int BN>
struct ConvCfg {
  static constexpr int A_BYTES = kBM * 128;
  static constexpr int W_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
  static constexpr int NS = (BN <= 64) ? 4 : (BN == 128 ? 3 : 2);
};

template <int BN>
__global__ void __launch_bounds__(288, 1) k_sparse_conv_tc(const float* __restrict__ X, int ldx, const unsigned char* __restrict__ Wp,
                                                           const int* __restrict__ nbr, const int* __restrict__ n_ptr, int n_max,
                                                           int K3, int Cin, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, const float* __restrict__ R, int ldr,
                                                           int relu, float* __restrict__ Y, int ldy, float* __restrict__ P,
                                                           int* err) {
  using Cfg = ConvCfg<BN>;
  constexpr int NS = Cfg::NS;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  int* nbr_s = reinterpret_cast<int*>(smem + NS * Cfg::STAGE_BYTES);     // [128][K3]
  __shared__ __align__(8) uint64_t full_bar[NS], empty_bar[NS], acc_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ unsigned kmask_s;
  __shared__ int klist_s[32];
  __shared__ int nk_s;

  int n_out = n_max;
  if (n_ptr) { const int v = *n_ptr; n_out = v < n_max ? v : n_max; }
  const int row0 = blockIdx.x * kBM;
  if (row0 >= n_out) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    kmask_s = 0u;
    for (int s = 0; s < NS; ++s) { tc::mbar_init(&full_bar[s], 257); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&acc_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 8) { tc::tmem_alloc(&tmem_base_s, BN < 32 ? 32 : BN); tc::tmem_relinquish(); }
  __syncthreads();
  {
    unsigned local = 0u;
    const int total = kBM * K3;
    const int* src = nbr + (size_t)row0 * K3;
    const int valid = (n_out - row0 < kBM ? n_out - row0 : kBM) * K3;
    for (int idx = tid; idx < total; idx += 288) {
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
    while (m) { const int b = __ffs(m) - 1; m &= m - 1; klist_s[c++] = b; }
    nk_s = c;
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_d = tmem_base_s;
  const int nchunks = Cin >> 5;
  const int nst_tile = nk_s * nchunks;
  // split of the stage list over grid.y
  const int per = (nst_tile + (int)gridDim.y - 1) / (int)gridDim.y;
  const int st_begin = min(nst_tile, (int)blockIdx.y * per);
  const int nst = min(nst_tile, st_begin + per) - st_begin;

  if (warp < 8) {
    // ------------------------------ gather producers ------------------------------
    for (int i = 0; i < nst; ++i) {
      const int s = i % NS;
      const uint32_t ph = (uint32_t)(i / NS) & 1u;
      const int st = st_begin + i;
      const int k = klist_s[st / nchunks];
      const int chunk = st % nchunks;
      const int c0 = chunk << 5;
      // issue the gathers first (they do not touch the stage buffer), then wait for the slot
      float4 v[4];
      bool have[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int id = it * 256 + tid, r = id >> 3, c = id & 7;
        const int src = nbr_s[r * K3 + k];
        have[it] = src >= 0;
        if (have[it]) v[it] = __ldg(reinterpret_cast<const float4*>(X + (size_t)src * ldx + c0 + c * 4));
      }
      tc::mbar_wait(&empty_bar[s], ph ^ 1u, err, 1);
      unsigned char* stg = smem + s * Cfg::STAGE_BYTES;
      if (tid == 0) {
        tc::mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::W_BYTES);
        tc::bulk_g2s(stg + 2 * Cfg::A_BYTES, Wp + ((size_t)k * nchunks + chunk) * (2 * Cfg::W_BYTES), 2 * Cfg::W_BYTES, &full_bar[s]);
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int id = it * 256 + tid, r = id >> 3, c = id & 7;
        float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), lo = hi;
        if (have[it]) tc::split_tf32(v[it], hi, lo);
        const uint32_t off = tc::sw128_offset(r, c);
        *reinterpret_cast<float4*>(stg + off) = hi;
        *reinterpret_cast<float4*>(stg + Cfg::A_BYTES + off) = lo;
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(&full_bar[s]);
    }
    // ------------------------------ epilogue ------------------------------
    if (nst > 0) {
      tc::mbar_wait(&acc_bar, 0u, err, 3);
      tc::tc_fence_after_sync();
    }
    const int lane_base = (warp & 3) * 32;
    const int row = row0 + lane_base + lane;
    constexpr int CW = (BN >= 32) ? BN / 2 : BN;            // columns per warp-group half
    const int col_base = (warp >> 2) * CW;
    float* out = P ? P + ((size_t)blockIdx.y * n_out) * BN : nullptr;
#pragma unroll 1
    for (int cb = 0; cb < CW; cb += 16) {
      float a[16];
      if (nst > 0) {
        tc::tmem_ld16(tmem_d + ((uint32_t)lane_base << 16) + (uint32_t)(col_base + cb), a);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = 0.f;
      }
      if (row < n_out) {
        const int c = col_base + cb;
        if (out) {
          float4* dst = reinterpret_cast<float4*>(out + (size_t)row * BN + c);
#pragma unroll
          for (int i = 0; i < 4; ++i) dst[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float x = a[i];
            if (scale) x = fmaf(x, __ldg(scale + c + i), __ldg(shift + c + i));
            if (R) x += R[(size_t)row * ldr + c + i];
            if (relu) x = fmaxf(x, 0.f);
            a[i] = x;
          }
          float* dst = Y + (size_t)row * ldy + c;
          if ((ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              reinterpret_cast<float4*>(dst)[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) dst[i] = a[i];
          }
        }
      }
    }
  } else {
    // ------------------------------ MMA issuer ------------------------------
    constexpr uint32_t idesc = tc::idesc_tf32(kBM, BN);
    for (int i = 0; i < nst; ++i) {
      const int s = i % NS;
      const uint32_t ph = (uint32_t)(i / NS) & 1u;
      tc::mbar_wait(&full_bar[s], ph, err, 2);
      tc::tc_fence_after_sync();
      if (lane == 0) {
        const uint32_t a_hi = tc::smem_u32(smem + s * Cfg::STAGE_BYTES), a_lo = a_hi + Cfg::A_BYTES;
        const uint32_t w_hi = a_hi + 2 * Cfg::A_BYTES, w_lo = w_hi + Cfg::W_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t o = ks * 32;
          tc::mma_tf32(tmem_d, tc::smem_desc_sw128(a_lo + o), tc::smem_desc_sw128(w_hi + o), idesc, (i | ks) ? 1u : 0u);
          tc::mma_tf32(tmem_d, tc::smem_desc_sw128(a_hi + o), tc::smem_desc_sw128(w_lo + o), idesc, 1u);
          tc::mma_tf32(tmem_d, tc::smem_desc_sw128(a_hi + o), tc::smem_desc_sw128(w_hi + o), idesc, 1u);
        }
        tc::mma_commit(&empty_bar[s]);
        if (i == nst - 1) tc::mma_commit(&acc_bar);
      }
      __syncwarp();
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem_d, BN < 32 ? 32 : BN);
}

// Y = act( (sum_z P[z]) * scale + shift (+ R) ) for the split variant.
__global__ void __launch_bounds__(256) k_conv_splitk_epilogue(const float* __restrict__ P, int splits, const int* __restrict__ n_ptr,
                                                              int n_max, int Cout, const float* __restrict__ scale,
                                                              const float* __restrict__ shift, const float* __restrict__ R, int ldr,
                                                              int relu, float* __restrict__ Y, int ldy) {
  int n = n_max;
  if (n_ptr) { const int v = *n_ptr; n = v < n_max ? v : n_max; }
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;     // one float4 of one row
  const int c4 = Cout >> 2;
  if (idx >= (long long)n * c4) return;
  const int row = (int)(idx / c4), c = (int)(idx % c4) * 4;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int z = 0; z < splits; ++z) {
    const float4 p = *reinterpret_cast<const float4*>(P + ((size_t)z * n + row) * Cout + c);
    a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
  }
  float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float x = v[i];
    if (scale) x = fmaf(x, __ldg(scale + c + i), __ldg(shift + c + i));
    if (R) x += R[(size_t)row * ldr + c + i];
    if (relu) x = fmaxf(x, 0.f);
    Y[(size_t)row * ldy + c + i] = x;
  }
}

// Pack W[K3][Cin][Cout] into per-(offset, 32-channel chunk) slabs: [hi image Cout x 128 B | lo image Cout x 128 B],
// each image in the SW128 K-major layout the MMA reads (row = output channel, 16-byte chunk index XOR (row % 8)).
__global__ void k_pack_conv_weights(const float* __restrict__ W, int K3, int Cin, int Cout, float* __restrict__ Wp) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)K3 * Cin * Cout;
  if (idx >= total) return;
  const int n = (int)(idx % Cout);
  const int ci = (int)((idx / Cout) % Cin);
  const int k = (int)(idx / ((long long)Cout * Cin));
  const float w = W[idx];
  const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
  const float lo = w - hi;
  const int nchunks = Cin >> 5, chunk = ci >> 5, j = ci & 31;
  const size_t slab = ((size_t)k * nchunks + chunk) * (size_t)(2 * Cout * 32);      // floats
  const size_t pos = (size_t)n * 32 + (size_t)((((j >> 2) ^ (n & 7)) << 2) | (j & 3));
  Wp[slab + pos] = hi;
  Wp[slab + (size_t)Cout * 32 + pos] = lo;
}

template <int BN>
int launch_tc(const float* X, int ldx, const void* Wp, const int* nbr, const int* n_ptr, int n_max, int K3, int Cin,
              const float* scale, const float* shift, const float* R, int ldr, int relu, float* Y, int ldy, void* ws, size_t ws_bytes,
              int* err, cudaStream_t stream) {
  using Cfg = ConvCfg<BN>;
  const size_t smem = (size_t)Cfg::NS * Cfg::STAGE_BYTES + (size_t)kBM * K3 * sizeof(int) + 1024;
  IMF_CHECK_CUDA(cudaFuncSetAttribute(k_sparse_conv_tc<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = (n_max + kBM - 1) / kBM;
  const int nst_max = K3 * (Cin >> 5);
  int splits = 1;
  if (ws != nullptr && tiles < 100 && nst_max >= 16) {
    splits = (148 + tiles - 1) / tiles;
    if (splits > nst_max / 6) splits = nst_max / 6;
    if (splits > 32) splits = 32;
    if (splits < 1) splits = 1;
    if (ws_bytes < (size_t)splits * n_max * BN * sizeof(float)) splits = 1;
  }
  float* P = splits > 1 ? reinterpret_cast<float*>(ws) : nullptr;
  dim3 grid(tiles, splits);
  k_sparse_conv_tc<BN><<<grid, 288, smem, stream>>>(X, ldx, reinterpret_cast<const unsigned char*>(Wp), nbr, n_ptr, n_max, K3, Cin, scale,
                                                    shift, R, ldr, relu, Y, ldy, P, err);
  IMF_CHECK_LAUNCH();
  if (splits > 1) {
    const long long total = (long long)n_max * (BN / 4);
    k_conv_splitk_epilogue<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(P, splits, n_ptr, n_max, BN, scale, shift, R, ldr, relu,
                                                                               Y, ldy);
    IMF_CHECK_LAUNCH();
  }
  return IMF_OK;
}

}  // namespace

extern "C" size_t imf_sparse_conv_tc_packed_bytes(int32_t kernel_volume, int32_t Cin, int32_t Cout) {
  return (size_t)kernel_volume * Cin * Cout * 2 * sizeof(float);
}

// One-off weight packing (hi/lo TF32 split, swizzled slabs) for imf_sparse_conv_tc_fwd.
extern "C" int imf_sparse_conv_tc_pack(const float* W, int32_t kernel_volume, int32_t Cin, int32_t Cout, void* packed,
                                       cudaStream_t stream) {
  IMF_CHECK_ARG(W != nullptr && packed != nullptr && kernel_volume >= 1 && Cin > 0 && Cin % 32 == 0 && Cout > 0 && Cout % 8 == 0);
  const long long total = (long long)kernel_volume * Cin * Cout;
  k_pack_conv_weights<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(W, kernel_volume, Cin, Cout, reinterpret_cast<float*>(packed));
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

extern "C" size_t imf_sparse_conv_tc_workspace_bytes(int32_t n_out_max, int32_t Cout) {
  return (size_t)32 * (size_t)(n_out_max > 0 ? n_out_max : 1) * Cout * sizeof(float);
}

// Tensor-core version of imf_sparse_conv_fwd; `packed` comes from imf_sparse_conv_tc_pack.  Cout in {32,64,128,256}.
// workspace (optional) enables splitting a tile's offsets over several CTAs when there are few row tiles.
extern "C" int imf_sparse_conv_tc_fwd(const float* X, int32_t ldx, const void* packed, const int32_t* nbr, const int32_t* n_out_dev,
                                      int32_t n_out_max, int32_t kernel_volume, int32_t Cin, int32_t Cout, const float* scale,
                                      const float* shift, const float* residual, int32_t ldr, int32_t relu, float* Y, int32_t ldy,
                                      void* workspace, size_t workspace_bytes, int32_t* err, cudaStream_t stream) {
  IMF_CHECK_ARG(n_out_max >= 0 && kernel_volume >= 1 && kernel_volume <= 27);
  IMF_CHECK_ARG(Cin > 0 && Cin % 32 == 0 && (Cout == 32 || Cout == 64 || Cout == 128 || Cout == 256));
  IMF_CHECK_ARG((scale == nullptr) == (shift == nullptr));
  IMF_CHECK_ARG(ldx % 4 == 0 && ldx >= Cin && ldy >= Cout && (residual == nullptr || ldr >= Cout));
  if (n_out_max == 0) return IMF_OK;
  IMF_CHECK_ARG(X != nullptr && packed != nullptr && nbr != nullptr && Y != nullptr);
  IMF_CHECK_ARG(((uintptr_t)X % 16) == 0 && ((uintptr_t)packed % 16) == 0);
#define IMF_GO(BN)                                                                                                       \
  return launch_tc<BN>(X, ldx, packed, nbr, n_out_dev, n_out_max, kernel_volume, Cin, scale, shift, residual, ldr, relu, Y, ldy, \
                       workspace, workspace_bytes, err, stream)
  if (Cout == 32) IMF_GO(32);
  if (Cout == 64) IMF_GO(64);
  if (Cout == 128) IMF_GO(128);
  IMF_GO(256);
#undef IMF_GO
}
