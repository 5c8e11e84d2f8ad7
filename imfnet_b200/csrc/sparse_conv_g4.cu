// imfnet_b200 -- sparse 3-D convolution, "g4" kernel: TMA tile::gather4 implicit GEMM on the tcgen05 tensor cores.
//
// Same operation and h2 data format as sparse_conv_h2.cu,
//   Y[o] = act( (sum_k X[nbr(o,k)] . W[k]) * scale + shift (+ R[o]) )
//   ME.MinkowskiConvolution / MinkowskiConvolutionTranspose + MinkowskiBatchNorm + ReLU / residual
//   /root/reference/model/resunet.py:168-213, model/residual_block.py:37-53,
// re-designed around what bounded that kernel on B200 (profiles/r01: 110 us for 64->64 at 50 k voxels, 70 % of it pipeline
// skeleton, 2 % DRAM, every 128-row tile re-reading all 27 weight slabs from L2):
//   * persistent grid (one CTA per SM); a CTA owns a contiguous range of output rows (several 128-row sub-tiles, one TMEM
//     accumulator each) and walks offsets OUTER / sub-tiles INNER, so a weight slab is fetched once per CTA, not once per tile;
//   * neighbour rows are fetched by the TMA engine (cp.async.bulk.tensor tile::gather4: four rows per instruction, absent
//     neighbours = out-of-range index = zero fill, 128-byte swizzle applied by the hardware) straight into the operand ring;
//     one warp issues a whole 128-row stage with one or two instructions per lane; completion is counted in bytes on an
//     mbarrier, so there is no per-thread wait / fence / arrive chain and every ring slot can be in flight;
//   * the neighbour table is offset-major (nbr_t[k][row]), so the indices of a stage are one coalesced 512-byte read, and a
//     per-tile offset mask (built by imf_kernel_map_t) lets (offset, sub-tile) pairs without any neighbour be skipped;
//   * the epilogue converts TMEM -> BatchNorm affine / residual / ReLU -> fp16 hi/lo in registers, stages the tile in shared
//     memory (swizzled, conflict-free) and writes it with tiled TMA stores;
//   * all sizes are read on the device (n_out_dev): the launch shape depends only on the SM count, which is what lets the
//     whole forward be captured in a CUDA graph;
//   * small levels (fewer tiles than SMs) split a tile's stage list over several CTAs into fp32 partials + a reduce kernel.
//
// CTA = 320 threads: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocation), warps 2-9 = epilogue.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"

namespace {

constexpr int kBM = 128;
constexpr int kImg = kBM * 128;          // one 128-row x 128-byte operand image (16 KB)
constexpr int kThreads = 320;
constexpr int kNW = 2;                   // weight-slab ring depth
constexpr int kMaxSubAll = 16;           // 512 TMEM columns / 32
constexpr int kSMs = 148;

template <int BN, int KC>
struct G4Cfg {
  static constexpr int A_BYTES = (KC == 64 ? 2 : 1) * kImg;
  static constexpr int W_IMG = BN * 128;
  static constexpr int W_BYTES = 2 * W_IMG;
  static constexpr int BUDGET = 200 * 1024;
  static constexpr int NA_FIT = (BUDGET - kNW * W_BYTES) / A_BYTES;
  static constexpr int NA = NA_FIT > 8 ? 8 : NA_FIT;
  static constexpr int MAXSUB = 512 / BN;
  static constexpr int OUT_BYTES = kBM * BN * 4;       // one staged output sub-tile (BN/32 images)
  static constexpr int RING_BYTES = NA * A_BYTES;
  static_assert(2 * OUT_BYTES <= RING_BYTES, "the epilogue double buffer lives in the operand ring");
};

__host__ __device__ constexpr uint32_t g4_idesc_f16(int M, int N) {
  return (1u << 4) /*D = f32*/ | (0u << 7) /*A = f16*/ | (0u << 10) /*B = f16*/ | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void g4_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void named_barrier(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---- work partition (identical in the convolution and in the reduce kernel) ---------------------------------------------
// row mode  : CTA x owns output rows [x*R, (x+1)*R) (R a multiple of 32), all offsets;
// split mode: CTA x owns tile x / S and the x % S -th part of that tile's stage list; partials go to the workspace when S > 1.
struct G4Part {
  int row_mode;
  int R;          // rows per CTA (row mode)
  int S;          // splits per tile (split mode)
  int T;          // 128-row tiles
};
__host__ __device__ inline G4Part g4_partition(int n, int gx, int nst_max, bool have_ws) {
  G4Part p;
  p.T = (n + kBM - 1) / kBM;
  p.row_mode = (p.T >= gx) ? 1 : 0;
  p.R = ((((n + gx - 1) / gx) + 31) / 32) * 32;
  int S = 1;
  if (!p.row_mode && have_ws && p.T > 0) {
    S = gx / p.T;
    if (S > nst_max / 4) S = nst_max / 4;
    if (S > 32) S = 32;
    if (S < 1) S = 1;
  }
  p.S = S;
  return p;
}

struct __align__(16) Half8 { __half2 a, b, c, d; };

// 16 floats -> fp16 hi / lo halves (8 + 8 per 16-byte vector)
__device__ __forceinline__ bool g4_split16(const float* x, Half8* hi, Half8* lo) {
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
  hi[0] = Half8{h[0], h[1], h[2], h[3]};
  hi[1] = Half8{h[4], h[5], h[6], h[7]};
  lo[0] = Half8{l[0], l[1], l[2], l[3]};
  lo[1] = Half8{l[4], l[5], l[6], l[7]};
  return big;
}
__device__ __forceinline__ void g4_load16_h2(const __half* hi_src, const __half* lo_src, float* x) {
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

// Position in the (offset-chunk, sub-tile) walk shared by the producer and the MMA warp.
struct G4It {
  int w, j;
};

template <int BN, int KC>
__global__ void __launch_bounds__(kThreads, 1)
k_sparse_conv_g4(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const unsigned char* __restrict__ Wp,
                 const int* __restrict__ nbr_t, int ld_n, const unsigned* __restrict__ tile_mask, const int* __restrict__ n_ptr, int n_max,
                 int K3, int nchunks, const float* __restrict__ scale, const float* __restrict__ shift, const __half* __restrict__ R, int ldr,
                 int kc_r, int relu, int kc_out, float* __restrict__ P, int cout_total, int* err, long long* __restrict__ trace) {
  using Cfg = G4Cfg<BN, KC>;
  constexpr int NA = Cfg::NA;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* a_ring = smem;                          // NA x A_BYTES   (re-used as the epilogue staging double buffer)
  unsigned char* w_ring = smem + Cfg::RING_BYTES;        // kNW x W_BYTES
  __shared__ __align__(8) uint64_t full_a[NA], empty_a[NA], full_w[kNW], empty_w[kNW], acc_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ unsigned submask_s[kMaxSubAll];
  __shared__ int klist_s[32];
  __shared__ int nk_s;
  __shared__ float sc_s[BN], sh_s[BN];

  int n = n_max;
  if (n_ptr) { const int v = *n_ptr; n = v < n_max ? v : n_max; }
  if (n <= 0) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gx = gridDim.x, bx = blockIdx.x, zt = blockIdx.z, ntn = gridDim.z;
  const int nst_max = K3 * nchunks;
  const G4Part part = g4_partition(n, gx, nst_max, P != nullptr);
  const int n32 = (n + 31) & ~31;

  // rows of this CTA and its sub-tiles
  int row_begin, row_end, split = 0;
  if (part.row_mode) {
    row_begin = bx * part.R;
    row_end = min(row_begin + part.R, n32);
  } else {
    const int tile = bx / part.S;
    split = bx % part.S;
    if (tile >= part.T) return;
    row_begin = tile * kBM;
    row_end = min(row_begin + kBM, n32);
  }
  if (row_begin >= row_end) return;
  const int nsub_total = (row_end - row_begin + kBM - 1) / kBM;
  if (bx | zt) trace = nullptr;
#define G4_TRACE(slot) do { if (trace) trace[slot] = clock64(); } while (0)
  if (tid == 0) G4_TRACE(0);

  if (tid == 0) {
    for (int s = 0; s < NA; ++s) { tc::mbar_init(&full_a[s], 1); tc::mbar_init(&empty_a[s], 1); }
    for (int s = 0; s < kNW; ++s) { tc::mbar_init(&full_w[s], 1); tc::mbar_init(&empty_w[s], 1); }
    tc::mbar_init(&acc_bar, 1);
    tc::fence_barrier_init();
    tma::prefetch_map(&tmX);
    tma::prefetch_map(&tmY);
  }
  if (warp == 1) {
    const int nsub_pass = nsub_total < Cfg::MAXSUB ? nsub_total : Cfg::MAXSUB;
    uint32_t cols = 32;
    while ((int)cols < nsub_pass * BN) cols <<= 1;
    tc::tmem_alloc(&tmem_base_s, cols);
    tc::tmem_relinquish();
  }
  if (tid >= 64 && tid < 64 + BN) {
    sc_s[tid - 64] = __ldg(scale + zt * BN + tid - 64);
    sh_s[tid - 64] = __ldg(shift + zt * BN + tid - 64);
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_d = tmem_base_s;
  uint32_t tmem_cols = 32;
  {
    const int nsub_pass = nsub_total < Cfg::MAXSUB ? nsub_total : Cfg::MAXSUB;
    while ((int)tmem_cols < nsub_pass * BN) tmem_cols <<= 1;
  }
  if (tid == 0) G4_TRACE(1);

  int ac = 0, wc = 0;                     // running stage counters (producer and MMA warp advance them identically)
  const int npass = (nsub_total + Cfg::MAXSUB - 1) / Cfg::MAXSUB;
  for (int pass = 0; pass < npass; ++pass) {
    const int sub0 = pass * Cfg::MAXSUB;
    const int nsub = min(Cfg::MAXSUB, nsub_total - sub0);
    const int prow = row_begin + sub0 * kBM;            // first row of this pass
    // ---- offset masks of the sub-tiles, offset list of the pass ----
    if (tid < nsub) {
      const int r0 = prow + tid * kBM;
      const int r1 = min(r0 + kBM, row_end) - 1;
      submask_s[tid] = __ldg(tile_mask + r0 / kBM) | __ldg(tile_mask + r1 / kBM);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned m = 0;
      for (int j = 0; j < nsub; ++j) m |= submask_s[j];
      int c = 0;
      while (m) { const int b = __ffs(m) - 1; m &= m - 1; klist_s[c++] = b; }
      nk_s = c;
    }
    __syncthreads();
    if (tid == 0) G4_TRACE(2);
    const int nw_all = nk_s * nchunks;                  // weight steps of the pass
    int w_begin = 0, w_end = nw_all;
    if (!part.row_mode && part.S > 1) {
      const int per = (nw_all + part.S - 1) / part.S;
      w_begin = min(nw_all, split * per);
      w_end = min(nw_all, w_begin + per);
    }
    auto active = [&](const G4It& it) { return (submask_s[it.j] >> klist_s[it.w / nchunks]) & 1u; };
    auto advance = [&](G4It& it) {
      do {
        if (++it.j == nsub) { it.j = 0; ++it.w; }
      } while (it.w < w_end && !active(it));
    };

    if (warp == 0) {
      // =========================== TMA producer ===========================
      auto load_idx = [&](const G4It& it) {
        int4 v = make_int4(-1, -1, -1, -1);
        const int row = prow + it.j * kBM + 4 * lane;
        if (row < row_end) v = __ldg(reinterpret_cast<const int4*>(nbr_t + (size_t)klist_s[it.w / nchunks] * ld_n + row));
        return v;
      };
      G4It cur{w_begin, -1};
      advance(cur);
      int4 idx = make_int4(-1, -1, -1, -1);
      if (cur.w < w_end) idx = load_idx(cur);
      int last_w = -1;
      while (cur.w < w_end) {
        G4It nxt = cur;
        advance(nxt);
        int4 idx_n = make_int4(-1, -1, -1, -1);
        if (nxt.w < w_end) idx_n = load_idx(nxt);       // prefetch the next stage's indices
        const int k = klist_s[cur.w / nchunks], chunk = cur.w % nchunks;
        if (cur.w != last_w) {
          const int ws = wc % kNW;
          tc::mbar_wait(&empty_w[ws], ((uint32_t)(wc / kNW) & 1u) ^ 1u, err, 1);
          if (lane == 0) {
            tc::mbar_arrive_expect_tx(&full_w[ws], Cfg::W_BYTES);
            tc::bulk_g2s(w_ring + ws * Cfg::W_BYTES, Wp + (((size_t)k * nchunks + chunk) * ntn + zt) * Cfg::W_BYTES, Cfg::W_BYTES,
                         &full_w[ws]);
          }
          ++wc;
          last_w = cur.w;
        }
        const int as = ac % NA;
        tc::mbar_wait(&empty_a[as], ((uint32_t)(ac / NA) & 1u) ^ 1u, err, 2);
        if (lane == 0) {
          if (trace && ac < 64) trace[16 + 2 * ac] = clock64();
          tc::mbar_arrive_expect_tx(&full_a[as], Cfg::A_BYTES);
        }
        __syncwarp();
        const uint32_t dst = tc::smem_u32(a_ring + as * Cfg::A_BYTES) + lane * 512;
        const uint32_t bar = tc::smem_u32(&full_a[as]);
        if (KC == 64) {
          tma::gather4(dst, &tmX, bar, chunk * 128, idx.x, idx.y, idx.z, idx.w);               // hi halves of the 64-channel chunk
          tma::gather4(dst + kImg, &tmX, bar, chunk * 128 + 64, idx.x, idx.y, idx.z, idx.w);   // lo halves
        } else {
          tma::gather4(dst, &tmX, bar, chunk * 64, idx.x, idx.y, idx.z, idx.w);                // [hi32 | lo32]
        }
        ++ac;
        cur = nxt;
        idx = idx_n;
      }
      // keep the MMA warp's counters in step (it walks the same sequence)
    } else if (warp == 1) {
      // =========================== MMA issuer ===========================
      constexpr uint32_t idesc = g4_idesc_f16(kBM, BN);
      G4It cur{w_begin, -1};
      advance(cur);
      int last_w = -1, ws = 0;
      unsigned started = 0u;
      while (cur.w < w_end) {
        if (cur.w != last_w) {
          if (last_w >= 0 && lane == 0) tc::mma_commit(&empty_w[ws]);
          ws = wc % kNW;
          tc::mbar_wait(&full_w[ws], (uint32_t)(wc / kNW) & 1u, err, 3);
          ++wc;
          last_w = cur.w;
        }
        const int as = ac % NA;
        tc::mbar_wait(&full_a[as], (uint32_t)(ac / NA) & 1u, err, 4);
        tc::tc_fence_after_sync();
        if (lane == 0) {
          if (trace && ac < 64) trace[17 + 2 * ac] = clock64();
          const uint32_t a0 = tc::smem_u32(a_ring + as * Cfg::A_BYTES);
          const uint32_t w0 = tc::smem_u32(w_ring + ws * Cfg::W_BYTES), w1 = w0 + Cfg::W_IMG;
          const uint32_t d = tmem_d + (uint32_t)(cur.j * BN);
          const uint32_t acc0 = (started >> cur.j) & 1u;
          if (KC == 64) {
            const uint32_t a_hi = a0, a_lo = a0 + kImg;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t o = ks * 32;
              g4_mma_f16(d, tc::smem_desc_sw128(a_lo + o), tc::smem_desc_sw128(w0 + o), idesc, (acc0 | (uint32_t)ks) ? 1u : 0u);
              g4_mma_f16(d, tc::smem_desc_sw128(a_hi + o), tc::smem_desc_sw128(w1 + o), idesc, 1u);
              g4_mma_f16(d, tc::smem_desc_sw128(a_hi + o), tc::smem_desc_sw128(w0 + o), idesc, 1u);
            }
          } else {
            // A row = [hi32 | lo32]; image w0 = [Whi | Whi], image w1 = [Wlo | 0]
            g4_mma_f16(d, tc::smem_desc_sw128(a0 + 64), tc::smem_desc_sw128(w0 + 64), idesc, acc0);     // lo . Whi
            g4_mma_f16(d, tc::smem_desc_sw128(a0 + 96), tc::smem_desc_sw128(w0 + 96), idesc, 1u);
            g4_mma_f16(d, tc::smem_desc_sw128(a0), tc::smem_desc_sw128(w1), idesc, 1u);                 // hi . Wlo
            g4_mma_f16(d, tc::smem_desc_sw128(a0 + 32), tc::smem_desc_sw128(w1 + 32), idesc, 1u);
            g4_mma_f16(d, tc::smem_desc_sw128(a0), tc::smem_desc_sw128(w0), idesc, 1u);                 // hi . Whi
            g4_mma_f16(d, tc::smem_desc_sw128(a0 + 32), tc::smem_desc_sw128(w0 + 32), idesc, 1u);
          }
          tc::mma_commit(&empty_a[as]);
        }
        started |= 1u << cur.j;
        ++ac;
        __syncwarp();
        advance(cur);
      }
      if (lane == 0) {
        if (last_w >= 0) tc::mma_commit(&empty_w[ws]);
        tc::mma_commit(&acc_bar);
      }
      __syncwarp();
    } else {
      // =========================== epilogue (8 warps) ===========================
      const int e = warp - 2, q = warp & 3, h = e >> 2;
      const int etid = tid - 64;
      if (tid == 64) G4_TRACE(3);
      tc::mbar_wait(&acc_bar, (uint32_t)pass & 1u, err, 5);
      tc::tc_fence_after_sync();
      if (tid == 64) G4_TRACE(4);
      // sub-tile j has an accumulator iff at least one of its (offset, chunk) stages was walked by this CTA
      unsigned started = 0u;
      if (w_end > w_begin)
        for (int j = 0; j < nsub; ++j) started |= (submask_s[j] != 0u ? 1u : 0u) << j;
      constexpr int CW = BN / 2;
      const bool partial = (P != nullptr) && !part.row_mode && part.S > 1;
      bool big = false;
      for (int j = 0; j < nsub; ++j) {
        const int r_in = q * 32 + lane;
        const int grow = prow + j * kBM + r_in;
        unsigned char* stage = a_ring + (j & 1) * Cfg::OUT_BYTES;
        if (!partial) {
          if (j >= 2 && etid == 0) tma::store_wait_read<1>();       // the stores that read this buffer two sub-tiles ago
          named_barrier(1, 256);
        }
#pragma unroll 1
        for (int cb = 0; cb < CW; cb += 16) {
          const int cl = h * CW + cb;                                // column inside this CTA's BN-wide tile
          float a[16];
          if ((started >> j) & 1u) {
            tc::tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * BN + cl), a);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = 0.f;
          }
          const int c = zt * BN + cl;                                // absolute output channel of a[0]
          if (partial) {
            if (grow < n) {
              float4* dst = reinterpret_cast<float4*>(P + ((size_t)split * n + grow) * cout_total + c);
#pragma unroll
              for (int i = 0; i < 4; ++i) dst[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
            }
            continue;
          }
          float r16[16];
          const bool has_r = (R != nullptr) && grow < n;
          if (has_r) {
            const __half* rp = R + (size_t)grow * ldr + (c / kc_r) * 2 * kc_r + (c % kc_r);
            g4_load16_h2(rp, rp + kc_r, r16);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float x = fmaf(a[i], sc_s[cl + i], sh_s[cl + i]);
            if (has_r) x += r16[i];
            if (relu) x = fmaxf(x, 0.f);
            a[i] = x;
          }
          Half8 hi[2], lo[2];
          const bool b = g4_split16(a, hi, lo);
          big |= b && (grow < n);
          // staged layout: image = 64 halves of the h2 row; kc_out = 64: images (hi, lo) per 64 channels; 32: one image [hi32|lo32]
          int img_hi, img_lo, ch_hi, ch_lo;
          if (kc_out == 64) {
            img_hi = (cl >> 6) * 2; img_lo = img_hi + 1; ch_hi = (cl & 63) >> 3; ch_lo = ch_hi;
          } else {
            img_hi = cl >> 5; img_lo = img_hi; ch_hi = (cl & 31) >> 3; ch_lo = 4 + ch_hi;
          }
          unsigned char* ph = stage + img_hi * kImg;
          unsigned char* pl = stage + img_lo * kImg;
          *reinterpret_cast<Half8*>(ph + tc::sw128_offset(r_in, ch_hi)) = hi[0];
          *reinterpret_cast<Half8*>(ph + tc::sw128_offset(r_in, ch_hi + 1)) = hi[1];
          *reinterpret_cast<Half8*>(pl + tc::sw128_offset(r_in, ch_lo)) = lo[0];
          *reinterpret_cast<Half8*>(pl + tc::sw128_offset(r_in, ch_lo + 1)) = lo[1];
        }
        if (!partial) {
          tc::fence_proxy_async();
          named_barrier(1, 256);
          if (etid == 0) {
            const int r0 = prow + j * kBM;
#pragma unroll 1
            for (int rb = 0; rb < 4; ++rb) {
              if (r0 + rb * 32 >= row_end) break;
#pragma unroll 1
              for (int img = 0; img < BN / 32; ++img)
                tma::store_2d(&tmY, tc::smem_u32(stage + img * kImg + rb * 4096), zt * BN * 2 + img * 64, r0 + rb * 32);
            }
            tma::store_commit();
          }
        }
      }
      if (!partial && etid == 0) tma::store_wait_read<0>();
      if (big && err) atomicOr(err, 0x10000);
      if (tid == 64) G4_TRACE(5);
    }
    tc::tc_fence_before_sync();
    __syncthreads();                       // pass boundary: TMEM drained, staging reads done, ring reusable
    tc::tc_fence_after_sync();
  }
  if (warp == 1) tc::tmem_dealloc(tmem_d, tmem_cols);
  if (tid == 0) G4_TRACE(6);
#undef G4_TRACE
}

// Y(h2) = act( (sum_s P[s]) * scale + shift (+ R) ) for the levels that ran in split mode with S > 1; one thread per (row, 16 ch).
__global__ void __launch_bounds__(256) k_conv_g4_reduce(const float* __restrict__ P, const int* __restrict__ n_ptr, int n_max, int gx,
                                                        int nst_max, int Cout, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, const __half* __restrict__ R, int ldr, int kc_r,
                                                        int relu, __half* __restrict__ Y, int ldy, int kc_out, int* err) {
  int n = n_max;
  if (n_ptr) { const int v = *n_ptr; n = v < n_max ? v : n_max; }
  const G4Part part = g4_partition(n, gx, nst_max, true);
  if (part.row_mode || part.S <= 1) return;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const int c16 = Cout >> 4;
  if (idx >= (long long)n * c16) return;
  const int row = (int)(idx / c16), c = (int)(idx % c16) * 16;
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 0.f;
  for (int z = 0; z < part.S; ++z) {
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
    g4_load16_h2(rp, rp + kc_r, r16);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float x = fmaf(a[i], __ldg(scale + c + i), __ldg(shift + c + i));
    if (R) x += r16[i];
    if (relu) x = fmaxf(x, 0.f);
    a[i] = x;
  }
  Half8 hi[2], lo[2];
  const bool big = g4_split16(a, hi, lo);
  __half* yp = Y + (size_t)row * ldy + (c / kc_out) * 2 * kc_out + (c % kc_out);
  reinterpret_cast<Half8*>(yp)[0] = hi[0];
  reinterpret_cast<Half8*>(yp)[1] = hi[1];
  reinterpret_cast<Half8*>(yp + kc_out)[0] = lo[0];
  reinterpret_cast<Half8*>(yp + kc_out)[1] = lo[1];
  if (big && err) atomicOr(err, 0x10000);
}

long long* g_g4_trace = nullptr;
int g_g4_grid = 0;        // profiling hook: overrides the number of CTAs per output-channel tile (0 = one per SM)

template <int BN, int KC>
int launch_g4(const CUtensorMap& tmX, const CUtensorMap& tmY, const void* Wp, const int* nbr_t, int ld_n, const unsigned* tile_mask,
              const int* n_ptr, int n_max, int K3, int Cin, int Cout, const float* scale, const float* shift, const __half* R, int ldr,
              int kc_r, int relu, __half* Y, int ldy, int kc_out, void* ws, size_t ws_bytes, int* err, cudaStream_t stream) {
  using Cfg = G4Cfg<BN, KC>;
  const size_t smem = (size_t)Cfg::RING_BYTES + (size_t)kNW * Cfg::W_BYTES + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    IMF_CHECK_CUDA(cudaFuncSetAttribute(k_sparse_conv_g4<BN, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int ntn = Cout / BN;
  const int nchunks = Cin / KC;
  int gx = (g_g4_grid > 0 ? g_g4_grid : kSMs) / ntn;
  const int tiles_max = (n_max + kBM - 1) / kBM;
  const int nst_max = K3 * nchunks;
  // never launch more CTAs than any partition of n <= n_max rows can use
  int useful = tiles_max >= gx ? gx : tiles_max * (nst_max / 4 > 1 ? (nst_max / 4 > 32 ? 32 : nst_max / 4) : 1);
  if (useful < gx) gx = useful < 1 ? 1 : useful;
  float* P = nullptr;
  if (ws != nullptr && ws_bytes >= (size_t)kSMs * kBM * Cout * sizeof(float)) P = reinterpret_cast<float*>(ws);
  dim3 grid(gx, 1, ntn);
  k_sparse_conv_g4<BN, KC><<<grid, kThreads, smem, stream>>>(tmX, tmY, reinterpret_cast<const unsigned char*>(Wp), nbr_t, ld_n, tile_mask,
                                                            n_ptr, n_max, K3, nchunks, scale, shift, R, ldr, kc_r, relu, kc_out, P, Cout,
                                                            err, g_g4_trace);
  IMF_CHECK_LAUNCH();
  if (P != nullptr) {      // split mode is possible for small n: the reduce kernel decides on the device (no-op otherwise)
    const int rows = n_max < gx * kBM ? n_max : gx * kBM;
    const long long total = (long long)rows * (Cout / 16);
    k_conv_g4_reduce<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(P, n_ptr, n_max, gx, nst_max, Cout, scale, shift, R, ldr, kc_r,
                                                                         relu, Y, ldy, kc_out, err);
    IMF_CHECK_LAUNCH();
  }
  return IMF_OK;
}

}  // namespace

extern "C" int imf_debug_conv_g4_trace(long long* trace, int32_t grid) {
  g_g4_trace = trace;
  g_g4_grid = grid;
  return IMF_OK;
}

extern "C" size_t imf_sparse_conv_g4_workspace_bytes(int32_t Cout) { return (size_t)kSMs * kBM * (size_t)Cout * sizeof(float); }

extern "C" int imf_sparse_conv_g4_fwd(const void* X, int32_t ldx, int32_t n_in_rows, int32_t kc_in, const void* packed, const int32_t* nbr_t,
                                      int32_t ld_n, const uint32_t* tile_mask, const int32_t* n_out_dev, int32_t n_out_max,
                                      int32_t kernel_volume, int32_t Cin, int32_t Cout, const float* scale, const float* shift,
                                      const void* residual, int32_t ldr, int32_t kc_r, int32_t relu, void* Y, int32_t ldy,
                                      int32_t n_y_rows, int32_t kc_out, void* workspace, size_t workspace_bytes, int32_t* err,
                                      cudaStream_t stream) {
  IMF_CHECK_ARG(n_out_max >= 0 && kernel_volume >= 1 && kernel_volume <= 27 && n_in_rows >= 0);
  IMF_CHECK_ARG((kc_in == 32 || kc_in == 64) && Cin > 0 && Cin % kc_in == 0 && (Cout == 32 || Cout == 64 || Cout == 128 || Cout == 256));
  IMF_CHECK_ARG((kc_out == 32 || kc_out == 64) && Cout % kc_out == 0);
  IMF_CHECK_ARG(scale != nullptr && shift != nullptr);
  IMF_CHECK_ARG(ldx % 8 == 0 && ldx >= 2 * Cin && ldy % 8 == 0 && ldy >= 2 * Cout && n_y_rows >= n_out_max);
  IMF_CHECK_ARG(residual == nullptr || ((kc_r == 32 || kc_r == 64) && Cout % kc_r == 0 && ldr % 8 == 0 && ldr >= 2 * Cout));
  IMF_CHECK_ARG(ld_n % 4 == 0 && ld_n >= ((n_out_max + 31) & ~31));
  if (n_out_max == 0) return IMF_OK;
  IMF_CHECK_ARG(X != nullptr && packed != nullptr && nbr_t != nullptr && tile_mask != nullptr && Y != nullptr && n_in_rows > 0);
  IMF_CHECK_ARG(((uintptr_t)X % 16) == 0 && ((uintptr_t)packed % 16) == 0 && ((uintptr_t)Y % 16) == 0 && ((uintptr_t)residual % 16) == 0 &&
                ((uintptr_t)nbr_t % 16) == 0);
  CUtensorMap tmX, tmY;
  int rc = tma::encode_2d_u16(&tmX, X, (uint64_t)n_in_rows, (uint64_t)(2 * Cin), (uint64_t)ldx, 64, 1);
  if (rc) { imf_set_error("cuTensorMapEncodeTiled(X) failed: %d", rc); return IMF_ERR_CUDA; }
  rc = tma::encode_2d_u16(&tmY, Y, (uint64_t)n_y_rows, (uint64_t)(2 * Cout), (uint64_t)ldy, 64, 32);
  if (rc) { imf_set_error("cuTensorMapEncodeTiled(Y) failed: %d", rc); return IMF_ERR_CUDA; }
  const __half* Rh = reinterpret_cast<const __half*>(residual);
  __half* Yh = reinterpret_cast<__half*>(Y);
#define IMF_GO(BN, KC)                                                                                                               \
  return launch_g4<BN, KC>(tmX, tmY, packed, nbr_t, ld_n, tile_mask, n_out_dev, n_out_max, kernel_volume, Cin, Cout, scale, shift, Rh, \
                           ldr, kc_r, relu, Yh, ldy, kc_out, workspace, workspace_bytes, err, stream)
  const int bn = Cout > 128 ? 128 : Cout;
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
