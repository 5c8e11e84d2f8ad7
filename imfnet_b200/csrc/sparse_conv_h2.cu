// imfnet_b200 -- sparse 3-D convolution as an implicit GEMM on the tcgen05 tensor cores, "h2" tier:
// activations live in HBM already split into two fp16 halves (v = hi + lo, 22 mantissa bits), so the gather is a pure
// asynchronous copy (cp.async, 16-byte vectors, zero-fill for absent neighbours) straight into the swizzled shared
// tiles the MMA reads, and every product is accumulated as lo*Whi + hi*Wlo + hi*Whi by kind::f16 MMAs into fp32 TMEM.
//
// Same contract as imf_sparse_conv_fwd (sparse_conv.cu): output-stationary gather over a neighbour table,
//   Y[o] = act( (sum_k X[nbr[o,k]] . W[k]) * scale + shift (+ R[o]) )
// for ME.MinkowskiConvolution / MinkowskiConvolutionTranspose + MinkowskiBatchNorm + ReLU / residual
//   /root/reference/model/resunet.py:168-213, model/residual_block.py:37-53.
//
// h2 matrix (C channels, chunk width KC in {32,64}, C % KC == 0), row stride ld (in halves, >= 2C):
//   channel c = q*KC + j  ->  hi at row*ld + q*2*KC + j,  lo at row*ld + q*2*KC + KC + j
// i.e. per row a sequence of [hi KC | lo KC] chunks: one 128-byte line (KC=32) or two adjacent lines (KC=64) per chunk.
//
// CTA = 288 threads, tile = 128 output rows x BN output channels (BN = min(Cout,128); grid.z walks Cout/BN):
//   warps 0-7  producers: per stage (offset k, input chunk) 128 rows x 2*KC halves by cp.async into SW128 K-major images;
//              D stages of copies stay in flight per thread (wait_group -> fence.proxy.async -> mbarrier arrive);
//              one thread adds the bulk (TMA) copy of the stage's pre-packed weight slab.  Then the epilogue:
//              TMEM -> registers -> BN affine / residual / ReLU -> fp16 hi/lo split -> global.
//   warp  8    TMEM allocation + tcgen05.mma issue (one lane): KC=64: 4 K-slices x 3 products; KC=32: 6 MMAs.
// Offsets with no neighbour in the whole tile are skipped; small levels split a tile's stage list over grid.y.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kBM = 128;
constexpr int kImg = kBM * 128;          // one 128-row x 128-byte operand image
constexpr int kLook = 2;                 // cp.async groups in flight per producer thread beyond the one being issued

template <int BN, int KC>
struct H2Cfg {
  static constexpr int A_BYTES = (KC == 64 ? 2 : 1) * kImg;
  static constexpr int W_IMG = BN * 128;
  static constexpr int W_BYTES = 2 * W_IMG;
  static constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
  static constexpr int NS_FIT = (205 * 1024) / STAGE_BYTES;
  static constexpr int NS = NS_FIT > 6 ? 6 : NS_FIT;
};

// 16-byte asynchronous copy global -> shared; a negative `row` writes zeros instead (ignore-src form: no address fix-ups).
__device__ __forceinline__ void cp_async16_row(uint32_t smem_dst, const void* gmem_src, int row) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.lt.s32 p, %2, 0;\n"
      "cp.async.cg.shared.global [%0], [%1], 16, p;\n"
      "}\n" ::"r"(smem_dst),
      "l"(gmem_src), "r"(row)
      : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ int lds_s32(uint32_t smem_addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];\n" : "=r"(v) : "r"(smem_addr));
  return v;
}

__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4) /*D = f32*/ | (0u << 7) /*A = f16*/ | (0u << 10) /*B = f16*/ | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct __align__(16) Half8 { __half2 a, b, c, d; };

// 16 floats -> hi/lo fp16 (two 32-byte runs); returns true if any |x| exceeds the fp16 range guard
__device__ __forceinline__ bool split16_store(const float* x, __half* hi_dst, __half* lo_dst) {
  __half2 h[8], l[8];
  bool big = false;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float x0 = x[2 * i], x1 = x[2 * i + 1];
    big |= (fabsf(x0) > 60000.f) | (fabsf(x1) > 60000.f);
    const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
    h[i] = __halves2half2(h0, h1);
    l[i] = __halves2half2(__float2half_rn(x0 - __half2float(h0)), __float2half_rn(x1 - __half2float(h1)));
  }
  Half8* hd = reinterpret_cast<Half8*>(hi_dst);
  Half8* ld = reinterpret_cast<Half8*>(lo_dst);
  hd[0] = Half8{h[0], h[1], h[2], h[3]};
  hd[1] = Half8{h[4], h[5], h[6], h[7]};
  ld[0] = Half8{l[0], l[1], l[2], l[3]};
  ld[1] = Half8{l[4], l[5], l[6], l[7]};
  return big;
}
__device__ __forceinline__ void load16_h2(const __half* hi_src, const __half* lo_src, float* x) {
  const Half8* hs = reinterpret_cast<const Half8*>(hi_src);
  const Half8* ls = reinterpret_cast<const Half8*>(lo_src);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const Half8 h = hs[q], l = ls[q];
    const __half2 hv[4] = {h.a, h.b, h.c, h.d}, lv[4] = {l.a, l.b, l.c, l.d};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 hf = __half22float2(hv[i]), lf = __half22float2(lv[i]);
      x[q * 8 + 2 * i] = hf.x + lf.x;
      x[q * 8 + 2 * i + 1] = hf.y + lf.y;
    }
  }
}

template <int BN, int KC>
__global__ void __launch_bounds__(288, 1) k_sparse_conv_h2(const __half* __restrict__ X, int ldx, const unsigned char* __restrict__ Wp,
                                                           const int* __restrict__ nbr, const int* __restrict__ n_ptr, int n_max,
                                                           int K3, int nchunks, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, const __half* __restrict__ R, int ldr,
                                                           int kc_r, int relu, __half* __restrict__ Y, int ldy, int kc_out,
                                                           float* __restrict__ P, int cout_total, int* err, int dbg, long long* __restrict__ trace) {
  using Cfg = H2Cfg<BN, KC>;
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
  const int ntn = gridDim.z, zt = blockIdx.z;
  if (blockIdx.x | blockIdx.y | blockIdx.z) trace = nullptr;            // profiling hook: CTA 0 only
#define H2_TRACE(slot) do { if (trace) trace[slot] = clock64(); } while (0)
  if (tid == 0) H2_TRACE(0);

  if (tid == 0) {
    kmask_s = 0u;
    for (int s = 0; s < NS; ++s) { tc::mbar_init(&full_bar[s], 9); tc::mbar_init(&empty_bar[s], 1); }   // full: 8 producer warps + expect_tx
    tc::mbar_init(&acc_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 8) { tc::tmem_alloc(&tmem_base_s, BN < 32 ? 32 : BN); tc::tmem_relinquish(); }
  __syncthreads();
  if (tid == 0) H2_TRACE(1);
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
  if (tid == 0) H2_TRACE(2);
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
  if (tid == 0) H2_TRACE(3);
  const int nst_tile = nk_s * nchunks;
  const int per = (nst_tile + (int)gridDim.y - 1) / (int)gridDim.y;
  const int st_begin = min(nst_tile, (int)blockIdx.y * per);
  const int nst = min(nst_tile, st_begin + per) - st_begin;

  if (warp < 8) {
    // ------------------------------ gather producers ------------------------------
    // Thread constants: KC=64: 16 lanes per row (hi 8 x 16 B | lo 8 x 16 B), rows r0 + 16*it; KC=32: 8 lanes per row, rows r0 + 32*it.
    constexpr int NIT = (KC == 64) ? 8 : 4;
    constexpr int RSTEP = (KC == 64) ? 16 : 32;
    const int r0 = (KC == 64) ? (tid >> 4) : (tid >> 3);
    const int c = tid & 7, part = (KC == 64) ? ((tid >> 3) & 1) : 0;
    const uint32_t smem_base = tc::smem_u32(smem);
    const uint32_t dst0 = part * kImg + tc::sw128_offset(r0, c);            // + it * RSTEP * 128 (r0 & 7 is unchanged by +16 / +32)
    const uint32_t nbr_addr0 = tc::smem_u32(nbr_s) + (uint32_t)(r0 * K3) * 4u;
    const char* xthr = reinterpret_cast<const char*>(X + part * 64 + c * 8);
    const unsigned ldx_bytes = (unsigned)ldx * 2u;
    int kidx = st_begin / nchunks, chunk = st_begin % nchunks;
    for (int i = 0; i < nst + kLook; ++i) {
      if (i < nst) {
        const int s = i % NS;
        const uint32_t ph = (uint32_t)(i / NS) & 1u;
        const int k = klist_s[kidx];
        int src[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) src[it] = lds_s32(nbr_addr0 + (uint32_t)((it * RSTEP * K3 + k) * 4));
        tc::mbar_wait(&empty_bar[s], ph ^ 1u, err, 1);
        if (tid == 0) H2_TRACE(16 + 4 * i);
        const uint32_t stg = smem_base + s * Cfg::STAGE_BYTES;
        if (tid == 0) {
          if (dbg & 1) {
            tc::mbar_arrive(&full_bar[s]);
          } else {
            tc::mbar_arrive_expect_tx(&full_bar[s], Cfg::W_BYTES);
            tc::bulk_g2s(smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES,
                         Wp + (((size_t)k * nchunks + chunk) * ntn + zt) * Cfg::W_BYTES, Cfg::W_BYTES, &full_bar[s]);
          }
        }
        const char* xc = xthr + chunk * (4 * KC);
        if (!(dbg & 2))
#pragma unroll
        for (int it = 0; it < NIT; ++it)      // absent neighbour: the (valid) address of row 0 is passed but ignored
          cp_async16_row(stg + dst0 + it * (RSTEP * 128), xc + (unsigned long long)((unsigned)max(src[it], 0)) * ldx_bytes, src[it]);
        if (++chunk == nchunks) { chunk = 0; ++kidx; }
        if (tid == 0) H2_TRACE(16 + 4 * i + 1);
      }
      cp_async_commit();
      if (i >= kLook) {
        cp_async_wait<kLook>();            // this thread's copies of stage i-kLook have landed
        if (tid == 0) H2_TRACE(16 + 4 * (i - kLook) + 2);
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&full_bar[(i - kLook) % NS]);
      }
    }
    // ------------------------------ epilogue ------------------------------
    if (tid == 0) H2_TRACE(4);
    if (nst > 0) {
      tc::mbar_wait(&acc_bar, 0u, err, 3);
      tc::tc_fence_after_sync();
    }
    if (tid == 0) H2_TRACE(5);
    const int lane_base = (warp & 3) * 32;
    const int row = row0 + lane_base + lane;
    constexpr int CW = BN / 2;                               // columns per warp-group half
    const int col_base = (warp >> 2) * CW;
    bool big = false;
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
        const int c = zt * BN + col_base + cb;               // absolute output channel of a[0]
        if (P) {
          float4* dst = reinterpret_cast<float4*>(P + ((size_t)blockIdx.y * n_out + row) * cout_total + c);
#pragma unroll
          for (int i = 0; i < 4; ++i) dst[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
        } else {
          float r16[16];
          if (R) {
            const __half* rp = R + (size_t)row * ldr + (c / kc_r) * 2 * kc_r + (c % kc_r);
            load16_h2(rp, rp + kc_r, r16);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float x = fmaf(a[i], __ldg(scale + c + i), __ldg(shift + c + i));
            if (R) x += r16[i];
            if (relu) x = fmaxf(x, 0.f);
            a[i] = x;
          }
          __half* yp = Y + (size_t)row * ldy + (c / kc_out) * 2 * kc_out + (c % kc_out);
          big |= split16_store(a, yp, yp + kc_out);
        }
      }
    }
    if (big && err) atomicOr(err, 0x10000);
    if (tid == 0) H2_TRACE(6);
  } else {
    // ------------------------------ MMA issuer ------------------------------
    constexpr uint32_t idesc = idesc_f16(kBM, BN);
    for (int i = 0; i < nst; ++i) {
      const int s = i % NS;
      const uint32_t ph = (uint32_t)(i / NS) & 1u;
      tc::mbar_wait(&full_bar[s], ph, err, 2);
      tc::tc_fence_after_sync();
      if (lane == 0) H2_TRACE(16 + 4 * i + 3);
      if (lane == 0 && (dbg & 4)) {
        tc::mma_commit(&empty_bar[s]);
        if (i == nst - 1) tc::mma_commit(&acc_bar);
      } else if (lane == 0) {
        const uint32_t a0 = tc::smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t w0 = a0 + Cfg::A_BYTES, w1 = w0 + Cfg::W_IMG;
        if (KC == 64) {
          const uint32_t a_hi = a0, a_lo = a0 + kImg;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t o = ks * 32;
            mma_f16(tmem_d, tc::smem_desc_sw128(a_lo + o), tc::smem_desc_sw128(w0 + o), idesc, (i | ks) ? 1u : 0u);
            mma_f16(tmem_d, tc::smem_desc_sw128(a_hi + o), tc::smem_desc_sw128(w1 + o), idesc, 1u);
            mma_f16(tmem_d, tc::smem_desc_sw128(a_hi + o), tc::smem_desc_sw128(w0 + o), idesc, 1u);
          }
        } else {
          // A row = [hi32 | lo32]; image w0 = [Whi | Whi], image w1 = [Wlo | 0]
          mma_f16(tmem_d, tc::smem_desc_sw128(a0 + 64), tc::smem_desc_sw128(w0 + 64), idesc, i ? 1u : 0u);   // lo . Whi
          mma_f16(tmem_d, tc::smem_desc_sw128(a0 + 96), tc::smem_desc_sw128(w0 + 96), idesc, 1u);
          mma_f16(tmem_d, tc::smem_desc_sw128(a0), tc::smem_desc_sw128(w1), idesc, 1u);                         // hi . Wlo
          mma_f16(tmem_d, tc::smem_desc_sw128(a0 + 32), tc::smem_desc_sw128(w1 + 32), idesc, 1u);
          mma_f16(tmem_d, tc::smem_desc_sw128(a0), tc::smem_desc_sw128(w0), idesc, 1u);                         // hi . Whi
          mma_f16(tmem_d, tc::smem_desc_sw128(a0 + 32), tc::smem_desc_sw128(w0 + 32), idesc, 1u);
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
  if (tid == 0) H2_TRACE(7);
#undef H2_TRACE
}

// Y(h2) = act( (sum_z P[z]) * scale + shift (+ R) ) for the split variant; one thread per (row, 16 channels).
__global__ void __launch_bounds__(256) k_conv_h2_splitk_epilogue(const float* __restrict__ P, int splits, const int* __restrict__ n_ptr,
                                                                 int n_max, int Cout, const float* __restrict__ scale,
                                                                 const float* __restrict__ shift, const __half* __restrict__ R, int ldr,
                                                                 int kc_r, int relu, __half* __restrict__ Y, int ldy, int kc_out, int* err) {
  int n = n_max;
  if (n_ptr) { const int v = *n_ptr; n = v < n_max ? v : n_max; }
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const int c16 = Cout >> 4;
  if (idx >= (long long)n * c16) return;
  const int row = (int)(idx / c16), c = (int)(idx % c16) * 16;
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 0.f;
  for (int z = 0; z < splits; ++z) {
    const float4* p = reinterpret_cast<const float4*>(P + ((size_t)z * n + row) * Cout + c);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = p[i];
      a[4 * i] += v.x; a[4 * i + 1] += v.y; a[4 * i + 2] += v.z; a[4 * i + 3] += v.w;
    }
  }
  float r16[16];
  if (R) {
    const __half* rp = R + (size_t)row * ldr + (c / kc_r) * 2 * kc_r + (c % kc_r);
    load16_h2(rp, rp + kc_r, r16);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float x = fmaf(a[i], __ldg(scale + c + i), __ldg(shift + c + i));
    if (R) x += r16[i];
    if (relu) x = fmaxf(x, 0.f);
    a[i] = x;
  }
  __half* yp = Y + (size_t)row * ldy + (c / kc_out) * 2 * kc_out + (c % kc_out);
  if (split16_store(a, yp, yp + kc_out) && err) atomicOr(err, 0x10000);
}

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
                                                 int* err, const int* __restrict__ n_ptr) {
  if (n_ptr) { const int v = *n_ptr; n = v < n ? v : n; }
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)n * C) return;
  const int row = (int)(idx / C), c = (int)(idx % C);
  const float x = X[(size_t)row * ldx + c];
  const __half h = __float2half_rn(x);
  __half* p = H + (size_t)row * ldh + (c / KC) * 2 * KC + (c % KC);
  p[0] = h;
  p[KC] = __float2half_rn(x - __half2float(h));
  if (fabsf(x) > 60000.f && err) atomicOr(err, 0x10000);
}
__global__ void __launch_bounds__(256) k_h2_unpack(const __half* __restrict__ H, int ldh, int n, int C, int KC, float* __restrict__ X, int ldx,
                                                   const int* __restrict__ n_ptr) {
  if (n_ptr) { const int v = *n_ptr; n = v < n ? v : n; }
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)n * C) return;
  const int row = (int)(idx / C), c = (int)(idx % C);
  const __half* p = H + (size_t)row * ldh + (c / KC) * 2 * KC + (c % KC);
  X[(size_t)row * ldx + c] = __half2float(p[0]) + __half2float(p[KC]);
}

extern int g_h2_debug_fwd;
extern long long* g_h2_trace;
template <int BN, int KC>
int launch_h2(const __half* X, int ldx, const void* Wp, const int* nbr, const int* n_ptr, int n_max, int K3, int Cin, int Cout,
              const float* scale, const float* shift, const __half* R, int ldr, int kc_r, int relu, __half* Y, int ldy, int kc_out,
              void* ws, size_t ws_bytes, int* err, cudaStream_t stream) {
  using Cfg = H2Cfg<BN, KC>;
  const size_t smem = (size_t)Cfg::NS * Cfg::STAGE_BYTES + (size_t)kBM * K3 * sizeof(int) + 1024;
  IMF_CHECK_CUDA(cudaFuncSetAttribute(k_sparse_conv_h2<BN, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = (n_max + kBM - 1) / kBM;
  const int ntn = Cout / BN;
  const int nchunks = Cin / KC;
  const int nst_max = K3 * nchunks;
  int splits = 1;
  if (ws != nullptr && tiles * ntn < 100 && nst_max >= 8) {
    splits = (148 + tiles * ntn - 1) / (tiles * ntn);
    if (splits > nst_max / 4) splits = nst_max / 4;
    if (splits > 32) splits = 32;
    if (splits < 1) splits = 1;
    if (ws_bytes < (size_t)splits * n_max * Cout * sizeof(float)) splits = 1;
  }
  float* P = splits > 1 ? reinterpret_cast<float*>(ws) : nullptr;
  dim3 grid(tiles, splits, ntn);
  k_sparse_conv_h2<BN, KC><<<grid, 288, smem, stream>>>(X, ldx, reinterpret_cast<const unsigned char*>(Wp), nbr, n_ptr, n_max, K3, nchunks,
                                                        scale, shift, R, ldr, kc_r, relu, Y, ldy, kc_out, P, Cout, err, g_h2_debug_fwd, g_h2_trace);
  IMF_CHECK_LAUNCH();
  if (splits > 1) {
    const long long total = (long long)n_max * (Cout / 16);
    k_conv_h2_splitk_epilogue<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(P, splits, n_ptr, n_max, Cout, scale, shift, R, ldr, kc_r,
                                                                                  relu, Y, ldy, kc_out, err);
    IMF_CHECK_LAUNCH();
  }
  return IMF_OK;
}

inline int h2_bn(int Cout) { return Cout > 128 ? 128 : Cout; }

int g_h2_debug_fwd = 0;
long long* g_h2_trace = nullptr;

}  // namespace

// Profiling hook: bit 0 skips the weight copies, bit 1 the gathers, bit 2 the MMAs of imf_sparse_conv_h2_fwd (results are then
// meaningless); used by tools/conv_microbench.py to attribute the kernel time.  0 = normal operation.
extern "C" int imf_debug_conv_flags(int32_t flags) {
  const int old = g_h2_debug_fwd;
  g_h2_debug_fwd = flags;
  return old;
}
// Profiling hook: device buffer of >= 16 + 4*stages int64 that CTA (0,0,0) of imf_sparse_conv_h2_fwd fills with clock64() stamps
// (slots: 0 start, 1 barriers+TMEM ready, 2 neighbour rows staged, 3 stage list ready, 4 producer loop done, 5 accumulator ready,
// 6 epilogue done, 7 exit; 16+4i+{0 slot acquired, 1 copies issued, 2 copies landed, 3 MMA warp saw full}).  NULL = off.
extern "C" int imf_debug_conv_trace(long long* trace) {
  g_h2_trace = trace;
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
extern "C" int imf_h2_pack_n(const float* X, int32_t ldx, int32_t n, const int32_t* n_dev, int32_t C, int32_t KC, void* H, int32_t ldh,
                             int32_t* err, cudaStream_t stream) {
  IMF_CHECK_ARG(n >= 0 && C > 0 && (KC == 32 || KC == 64) && C % KC == 0 && ldx >= C && ldh >= 2 * C);
  if (n == 0) return IMF_OK;
  IMF_CHECK_ARG(X != nullptr && H != nullptr);
  const long long total = (long long)n * C;
  k_h2_pack<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(X, ldx, n, C, KC, reinterpret_cast<__half*>(H), ldh, err, n_dev);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

extern "C" int imf_h2_unpack(const void* H, int32_t ldh, int32_t n, int32_t C, int32_t KC, float* X, int32_t ldx, cudaStream_t stream) {
  return imf_h2_unpack_n(H, ldh, n, nullptr, C, KC, X, ldx, stream);
}
extern "C" int imf_h2_unpack_n(const void* H, int32_t ldh, int32_t n, const int32_t* n_dev, int32_t C, int32_t KC, float* X, int32_t ldx,
                               cudaStream_t stream) {
  IMF_CHECK_ARG(n >= 0 && C > 0 && (KC == 32 || KC == 64) && C % KC == 0 && ldx >= C && ldh >= 2 * C);
  if (n == 0) return IMF_OK;
  IMF_CHECK_ARG(X != nullptr && H != nullptr);
  const long long total = (long long)n * C;
  k_h2_unpack<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const __half*>(H), ldh, n, C, KC, X, ldx, n_dev);
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
  IMF_CHECK_ARG(Cout == 32 || Cout == 64 || Cout == 128 || Cout == 256);
  IMF_CHECK_ARG(wmul > 0.f);
  const long long total = (long long)kernel_volume * Cin * Cout;
  k_pack_conv_weights_h2<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(W, kernel_volume, Cin, Cout, kc_in, h2_bn(Cout), wmul,
                                                                            reinterpret_cast<__half*>(packed));
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

extern "C" size_t imf_sparse_conv_h2_workspace_bytes(int32_t n_out_max, int32_t Cout) {
  return (size_t)32 * (size_t)(n_out_max > 0 ? n_out_max : 1) * Cout * sizeof(float);
}

extern "C" int imf_sparse_conv_h2_fwd(const void* X, int32_t ldx, int32_t kc_in, const void* packed, const int32_t* nbr,
                                      const int32_t* n_out_dev, int32_t n_out_max, int32_t kernel_volume, int32_t Cin, int32_t Cout,
                                      const float* scale, const float* shift, const void* residual, int32_t ldr, int32_t kc_r,
                                      int32_t relu, void* Y, int32_t ldy, int32_t kc_out, void* workspace, size_t workspace_bytes,
                                      int32_t* err, cudaStream_t stream) {
  IMF_CHECK_ARG(n_out_max >= 0 && kernel_volume >= 1 && kernel_volume <= 27);
  IMF_CHECK_ARG((kc_in == 32 || kc_in == 64) && Cin > 0 && Cin % kc_in == 0 && (Cout == 32 || Cout == 64 || Cout == 128 || Cout == 256));
  IMF_CHECK_ARG((kc_out == 32 || kc_out == 64) && Cout % kc_out == 0);
  IMF_CHECK_ARG(scale != nullptr && shift != nullptr);
  IMF_CHECK_ARG(ldx % 8 == 0 && ldx >= 2 * Cin && ldy % 8 == 0 && ldy >= 2 * Cout);
  IMF_CHECK_ARG(residual == nullptr || ((kc_r == 32 || kc_r == 64) && Cout % kc_r == 0 && ldr % 8 == 0 && ldr >= 2 * Cout));
  if (n_out_max == 0) return IMF_OK;
  IMF_CHECK_ARG(X != nullptr && packed != nullptr && nbr != nullptr && Y != nullptr);
  IMF_CHECK_ARG(((uintptr_t)X % 16) == 0 && ((uintptr_t)packed % 16) == 0 && ((uintptr_t)Y % 16) == 0 && ((uintptr_t)residual % 16) == 0);
  const __half* Xh = reinterpret_cast<const __half*>(X);
  const __half* Rh = reinterpret_cast<const __half*>(residual);
  __half* Yh = reinterpret_cast<__half*>(Y);
#define IMF_GO(BN, KC)                                                                                                              \
  return launch_h2<BN, KC>(Xh, ldx, packed, nbr, n_out_dev, n_out_max, kernel_volume, Cin, Cout, scale, shift, Rh, ldr, kc_r, relu, Yh, \
                           ldy, kc_out, workspace, workspace_bytes, err, stream)
  const int bn = h2_bn(Cout);
  if (kc_in == 64) {
    if (bn == 32) IMF_GO(32, 64);
    if (bn == 64) IMF_GO(64, 64);
    IMF_GO(128, 64);
  }
  if (bn == 32) IMF_GO(32, 32);
  if (bn == 64) IMF_GO(64, 32);
  IMF_GO(128, 32);
#undef IMF_GO
}
