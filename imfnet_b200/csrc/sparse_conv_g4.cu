// imfnet_b200 -- sparse 3-D convolution, "g4" kernel: persistent output-stationary implicit GEMM on the tcgen05 tensor cores.
//
// Same operation and h2 data format as sparse_conv_h2.cu,
//   Y[o] = act( (sum_k X[nbr(o,k)] . W[k]) * scale + shift (+ R[o]) )
//   ME.MinkowskiConvolution / MinkowskiConvolutionTranspose + MinkowskiBatchNorm + ReLU / residual
//   /root/reference/model/resunet.py:168-213, model/residual_block.py:37-53,
// re-designed around what bounded that kernel on B200 (profiles/r01: 110 us for 64->64 at 50 k voxels, 70 % of it pipeline
// skeleton, 2 % DRAM, every 128-row tile re-reading all 27 weight slabs from L2):
//   * persistent grid (one CTA per SM); a CTA owns consecutive whole 128-row tiles (one TMEM accumulator pair each) and walks
//     offsets OUTER / sub-tiles INNER, so a weight slab is fetched once per CTA, not once per tile;
//   * neighbour rows are gathered with 16-byte cp.async (LDGSTS) straight into 128-byte-swizzled operand images, two warps per
//     ring slot, completion counted on the slot's mbarrier (cp.async.mbarrier.arrive.noinc): no thread waits for its own copies
//     and every ring slot can be in flight.  (TMA tile::gather4 was measured first -- hence the kernel's name -- and rejected:
//     84 cycles per instruction and issuing warp, 44 B/cycle/SM at best, zero-filled rows 5x slower; tools/tma_rate.py);
//   * the neighbour table is offset-major (nbr_t[k][row]), so the indices of a stage are one coalesced 512-byte read, and a
//     per-tile offset mask (built by imf_kernel_map_t) lets (offset, sub-tile) pairs without any neighbour be skipped;
//   * MMAs are issued from warp-uniform code (descriptors in uniform registers, bare UTCHMMA back to back), each accumulator by
//     one thread only, the issuing warps taking turns in walk order: bit-reproducible sums (see the MMA section);
//   * the epilogue converts TMEM -> BatchNorm affine / residual / ReLU -> fp16 hi/lo in registers, stages the tile in shared
//     memory (swizzled, conflict-free) and writes it with tiled TMA stores;
//   * all sizes are read on the device (n_out_dev): the launch shape depends only on the SM count, which is what lets the
//     whole forward be captured in a CUDA graph;
//   * small levels (fewer tiles than SMs) run one CTA per tile, or -- when the caller passes a workspace -- split a tile's stage
//     list over several CTAs into fp32 partials + a reduce kernel (shorter latency, more SM time).
//
// CTA = 13 to 17 warps: warps 0-7 = gather producers, then the epilogue; warps 8-11 = MMA issuers (warp 8 also owns the TMEM
// allocation); warp 12 = weight-slab loader (bulk TMA copies, residual prefetch); warps 13-16 = further gather producers when two
// warps share a ring slot (5 or 6 slots of 32 KB).
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"

namespace {

constexpr int kBM = 128;
constexpr int kImg = kBM * 128;          // one 128-row x 128-byte operand image (16 KB)
constexpr int kNPW = 8;                  // gather producer warps that also run the epilogue (warps 0-7)
constexpr int kNMW = 4;                  // MMA issuing warps (8-11)
constexpr int kNXW = 4;                  // extra gather producer warps (13-16) for the configurations with two warps per ring slot
constexpr int kNW = 2;                   // weight-slab ring depth
constexpr int kMaxSubAll = 8;            // 512 TMEM columns / (2 * 32)

template <int BN, int KC>
struct G4Cfg {
  static constexpr int A_BYTES = (KC == 64 ? 2 : 1) * kImg;
  static constexpr int W_IMG = BN * 128;
  static constexpr int W_BYTES = 2 * W_IMG;
  // Rings: everything the 227 KB of a CTA allow next to ~1.5 KB of static shared memory and the 1 KB alignment slack (the pipeline is
  // latency-bound -- no unit above 60 % in ncu, profiles/r02/call28 -- so every further stage in flight counts: 64 -> 64 runs with 6
  // slots of 32 KB, 128-wide output tiles with 5; the neighbour indices of a stage travel through registers, not shared memory).
  static constexpr int BUDGET = 224 * 1024;
  static constexpr int NA_FIT = (BUDGET - kNW * W_BYTES) / A_BYTES;
  static constexpr int NA = NA_FIT > 8 ? 8 : NA_FIT;
  static constexpr int HALVES = NA <= 6 ? 2 : 1;         // producer warps per ring slot (each copies 128 / HALVES rows of a stage)
  static constexpr int NPROD = NA * HALVES;
  static_assert(NPROD <= kNPW + kNXW, "not enough producer warps");
  static constexpr int THREADS = (kNPW + kNMW + 1 + (NPROD > kNPW ? NPROD - kNPW : 0)) * 32;   // warp 12 = weight-slab loader
  static constexpr int ACC_COLS = 2 * BN;               // D1 = hi.Whi + lo.Whi, D2 = hi.Wlo (summed in the epilogue)
  static constexpr int MAXSUB = 512 / ACC_COLS;
  static constexpr int OUT_BYTES = kBM * BN * 4;       // one staged output sub-tile (BN/32 images)
  static constexpr int RING_BYTES = NA * A_BYTES;
  // Neighbour indices of a stage: through shared memory where it is not scarce (KC = 32: 128 KB ring), by shuffle otherwise (KC = 64:
  // the 5 KB of staging are what the last ring slot needs; the four shuffles per row group cost the 32 -> 32 layers 5 %, call 33).
  static constexpr bool IDX_SMEM = (KC == 32);
  // Residual sub-tiles are copied (coalesced cp.async) into the ring slots the epilogue's staging double buffer leaves free.
  static constexpr int NRB_FIT = (RING_BYTES - 2 * (kBM * BN * 4)) / (kBM * BN * 4);
  static constexpr int NRB = NRB_FIT > 8 ? 8 : (NRB_FIT < 0 ? 0 : NRB_FIT);
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
// 16-byte asynchronous copy global -> shared; a negative `row` writes zeros instead (ignore-src form)
__device__ __forceinline__ void g4_cp_async16_row(uint32_t smem_dst, const void* gmem_src, int row) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.lt.s32 p, %2, 0;\n"
      "cp.async.cg.shared.global [%0], [%1], 16, p;\n"
      "}\n" ::"r"(smem_dst),
      "l"(gmem_src), "r"(row)
      : "memory");
}
__device__ __forceinline__ void g4_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void g4_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
// make `bar` receive one (pre-counted) arrival once all cp.async issued so far by this thread have completed
__device__ __forceinline__ void g4_cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
// zero 16 consecutive TMEM columns of this warp's 32 lanes
__device__ __forceinline__ void g4_tmem_zero16(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(taddr), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void named_barrier(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---- work partition (identical in the convolution and in the reduce kernel) ---------------------------------------------
// row mode  : CTA x owns T / gx (+1 for the first T % gx CTAs) consecutive 128-row tiles, all offsets;
// split mode: CTA x owns tile x / S and the x % S -th part of that tile's stage list; partials go to the workspace when S > 1.
struct G4Part {
  int row_mode;
  int S;          // splits per tile (split mode)
  int T;          // 128-row tiles
};
__host__ __device__ inline G4Part g4_partition(int n, int gx, int nst_max, bool have_ws, int stages_per_split) {
  G4Part p;
  p.T = (n + kBM - 1) / kBM;
  p.row_mode = (p.T >= gx) ? 1 : 0;
  int S = 1;
  if (!p.row_mode && have_ws && p.T > 0) {
    // as many splits as free SMs allow, but not below `stages_per_split` stages per CTA: a split costs a partial-tile round trip
    // through L2 and a share of the reduce kernel, which only pays when it shortens a long stage list
    // (stages_per_split: low 16 bits = minimum stages per split CTA, high 16 bits = upper limit of S, 0 = 32)
    const int sps = stages_per_split & 0xffff, cap = (stages_per_split >> 16) ? (stages_per_split >> 16) : 32;
    S = gx / p.T;
    const int smax = (nst_max + sps - 1) / sps;
    if (S > smax) S = smax;
    if (S > cap) S = cap;
    if (S < 1) S = 1;
  }
  p.S = S;
  return p;
}

struct __align__(16) Half8 { __half2 a, b, c, d; };

// 16 floats -> fp16 hi / lo halves (8 + 8 per 16-byte vector), packed conversions (one instruction per pair); returns max |x|
__device__ __forceinline__ float g4_split16(const float* x, Half8* hi, Half8* lo) {
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
  hi[0] = Half8{h[0], h[1], h[2], h[3]};
  hi[1] = Half8{h[4], h[5], h[6], h[7]};
  lo[0] = Half8{l[0], l[1], l[2], l[3]};
  lo[1] = Half8{l[4], l[5], l[6], l[7]};
  return m;
}
// 8 + 8 halves (hi, lo) -> 8 floats
__device__ __forceinline__ void g4_unpack8_h2(const int4& h4, const int4& l4, float* x) {
  Half8 h, l;
  *reinterpret_cast<int4*>(&h) = h4;
  *reinterpret_cast<int4*>(&l) = l4;
  const __half2 hv[4] = {h.a, h.b, h.c, h.d}, lv[4] = {l.a, l.b, l.c, l.d};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 hf = __half22float2(hv[i]), lf = __half22float2(lv[i]);
    x[2 * i] = hf.x + lf.x;
    x[2 * i + 1] = hf.y + lf.y;
  }
}
__device__ __forceinline__ int4 g4_lds16(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void g4_load16_h2(const __half* hi_src, const __half* lo_src, float* x) {
  const Half8* hs = reinterpret_cast<const Half8*>(hi_src);
  const Half8* ls = reinterpret_cast<const Half8*>(lo_src);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    // explicit 16-byte loads (plain struct copies compile to eight 4-byte LDG.E.CONSTANT per 32 bytes)
    const int4 h4 = __ldg(reinterpret_cast<const int4*>(hs) + q), l4 = __ldg(reinterpret_cast<const int4*>(ls) + q);
    Half8 h, l;
    *reinterpret_cast<int4*>(&h) = h4;
    *reinterpret_cast<int4*>(&l) = l4;
    const __half2 hv[4] = {h.a, h.b, h.c, h.d}, lv[4] = {l.a, l.b, l.c, l.d};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 hf = __half22float2(hv[i]), lf = __half22float2(lv[i]);
      x[q * 8 + 2 * i] = hf.x + lf.x;
      x[q * 8 + 2 * i + 1] = hf.y + lf.y;
    }
  }
}

// Position in the (offset, chunk, sub-tile) walk shared by the producers and the MMA warps.
struct G4It {
  int w;          // weight step = kk * nchunks + chunk
  int kk;         // index into the offset list
  int chunk;      // input-channel chunk
  int k;          // klist[kk]
  unsigned jm;    // sub-tiles that have a neighbour at offset k
  unsigned rem;   // sub-tiles of this weight step still to visit (lowest set bit = current)
  int j;          // current sub-tile
};

template <int BN, int KC>
__global__ void __launch_bounds__((G4Cfg<BN, KC>::THREADS), 1)
k_sparse_conv_g4(const __half* __restrict__ X, int ldx, const __grid_constant__ CUtensorMap tmY, const unsigned char* __restrict__ Wp,
                 const int* __restrict__ nbr_t, int ld_n, const unsigned* __restrict__ tile_mask, const int* __restrict__ n_ptr, int n_max,
                 int K3, int nchunks, const float* __restrict__ scale, const float* __restrict__ shift, const __half* __restrict__ R, int ldr,
                 int kc_r, int relu, int kc_out, float* __restrict__ P, int cout_total, const int* __restrict__ out_row, __half* __restrict__ Yp,
                 int ldy, int* err, long long* __restrict__ trace, int dbg, int sps) {
  using Cfg = G4Cfg<BN, KC>;
  constexpr int NA = Cfg::NA, HALVES = Cfg::HALVES, NPROD = Cfg::NPROD, HROWS = kBM / HALVES;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* a_ring = smem;                          // NA x A_BYTES   (re-used as the epilogue staging double buffer)
  unsigned char* w_ring = smem + Cfg::RING_BYTES;        // kNW x W_BYTES
  __shared__ __align__(8) uint64_t full_a[NA], empty_a[NA], full_w[kNW], empty_w[kNW], acc_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ unsigned submask_s[kMaxSubAll];
  __shared__ int klist_s[32];
  __shared__ unsigned jmask_s[32];
  __shared__ int nk_s;
  __shared__ int ring_s[6];
  __shared__ int slot_turn_s[NA];                        // per ring slot: uses of the slot whose stage has been seen full by its MMA warp
  __shared__ __align__(16) float sc_s[BN], sh_s[BN];
  __shared__ __align__(8) uint64_t res_bar[Cfg::NRB > 0 ? Cfg::NRB : 1];     // residual sub-tile landed in its buffer (256 cp.async arrivals)
  __shared__ __align__(16) int idx_s[Cfg::IDX_SMEM ? NPROD : 1][2][Cfg::IDX_SMEM ? HROWS : 4];   // KC = 32: indices of a producer warp's current / next stage

#if !defined(IMF_G4_TRACE)
  trace = nullptr;      // the clock64 trace hooks only exist in a -DIMF_G4_TRACE build (tools/conv_g4_bench.py --trace)
#endif
  int n = n_max;
  if (n_ptr) { const int v = *n_ptr; n = v < n_max ? v : n_max; }
  if (n <= 0) return;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler too (role branches, uniform registers)
  const int gx = gridDim.x, bx = blockIdx.x, zt = blockIdx.z, ntn = gridDim.z;
  const int nst_max = K3 * nchunks;
  const G4Part part = g4_partition(n, gx, nst_max, P != nullptr, sps);
  const int n32 = (n + 31) & ~31;

  // Rows of this CTA, pass by pass.  Row mode: the T tiles are cut into `npass` equal WINDOWS of at most gx * MAXSUB tiles; in pass p
  // every CTA takes its share of window p -- whole 128-row tiles, the first (window size % gx) CTAs one more than the others (no CTA
  // carries a mostly empty trailing sub-tile: an MMA costs the same for 32 rows as for 128).  All CTAs therefore work on the same
  // ~gx * MAXSUB * 128 rows (75 k at BN = 64: about 1.5 fragments of a batched launch) at the same time, and their gathers -- a row's
  // neighbours are rows of the same fragment -- stay inside the L2 (126 MB) instead of sweeping all rows of a 500 k-row batch
  // (128 MB of 64-channel activations + 54 MB of indices) concurrently; per CTA the tile count and the number of passes are what a
  // contiguous partition gives.  Split mode: one tile per CTA, one pass.
  int split = 0, npass = 1;
  if (part.row_mode) {
    npass = (part.T + gx * Cfg::MAXSUB - 1) / (gx * Cfg::MAXSUB);
  } else {
    split = bx % part.S;
    if (bx / part.S >= part.T) return;
  }
  auto pass_rows = [&](int pass, int& prow, int& pend) {       // rows [prow, pend) of this CTA in pass `pass`; returns the sub-tiles
    int t0, cnt;
    if (part.row_mode) {
      const int w0 = (int)(((long long)part.T * pass) / npass), w1 = (int)(((long long)part.T * (pass + 1)) / npass);
      const int wt = w1 - w0, base = wt / gx, extra = wt - base * gx;
      t0 = w0 + bx * base + min(bx, extra);
      cnt = base + (bx < extra ? 1 : 0);
    } else {
      t0 = bx / part.S;
      cnt = 1;
    }
    prow = t0 * kBM;
    pend = min((t0 + cnt) * kBM, n32);
    return pend > prow ? (pend - prow + kBM - 1) / kBM : 0;
  };
  int nsub_max = 0;
  {
    int a, b;
    for (int p = 0; p < npass; ++p) nsub_max = max(nsub_max, pass_rows(p, a, b));
  }
  if (nsub_max == 0) return;
  long long* cta_trace = (trace && zt == 0) ? trace + 256 : nullptr;      // per-CTA lifetime (profiling hook)
  const long long t_cta0 = clock64();
  if (bx | zt) trace = nullptr;
#define G4_TRACE(slot) do { if (trace) trace[slot] = clock64(); } while (0)
  if (tid == 0) G4_TRACE(0);

  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < nsub_max * Cfg::ACC_COLS) tmem_cols <<= 1;
  if (tid == 0) {
    for (int s = 0; s < NA; ++s) { tc::mbar_init(&full_a[s], 32 * HALVES); tc::mbar_init(&empty_a[s], 1); }
    for (int s = 0; s < kNW; ++s) { tc::mbar_init(&full_w[s], 1); tc::mbar_init(&empty_w[s], kNMW); }
    tc::mbar_init(&acc_bar, kNMW);
    for (int s = 0; s < Cfg::NRB; ++s) tc::mbar_init(&res_bar[s], 256);
    for (int s = 0; s < NA; ++s) slot_turn_s[s] = 0;
    tc::fence_barrier_init();
    tma::prefetch_map(&tmY);
  }
  if (warp == kNPW) { tc::tmem_alloc(&tmem_base_s, tmem_cols); tc::tmem_relinquish(); }
  if (tid < BN) {
    sc_s[tid] = __ldg(scale + zt * BN + tid);
    sh_s[tid] = __ldg(shift + zt * BN + tid);
  }
  // offset masks of the first pass' sub-tiles: loaded here so that their latency overlaps the TMEM allocation
  if (tid >= 64 && tid < 64 + kMaxSubAll) {
    const int j = tid - 64;
    int prow0, pend0;
    const int nsub0 = pass_rows(0, prow0, pend0);
    if (j < nsub0) {
      const int r0 = prow0 + j * kBM;
      const int r1 = min(r0 + kBM, pend0) - 1;
      submask_s[j] = __ldg(tile_mask + r0 / kBM) | __ldg(tile_mask + r1 / kBM);
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_d = tmem_base_s;
  if (tid == 0) G4_TRACE(1);

  // running ring positions (the producers and the MMA warps advance them identically)
  int ac = 0, a_slot = 0, w_slot = 0, a_use = 0;       // a_use: how often the ring has wrapped (= use index of a slot, a_phase = a_use & 1)
  uint32_t a_phase = 0u, w_phase = 0u;
  int passes_done = 0;                                  // executed passes (parity of the accumulator barrier)
  uint32_t res_phase = 0u;                              // bit b: parity of residual buffer b's next completion
  for (int pass = 0; pass < npass; ++pass) {
    int prow, row_end;                                  // first row of this pass, end of this CTA's rows in it
    const int nsub = pass_rows(pass, prow, row_end);
    if (nsub == 0) continue;                            // (CTA-uniform: a window with fewer tiles than CTAs)
    // ---- offset masks of the sub-tiles; offset list of the pass with, per offset, the sub-tiles that need it ----
    if (pass > 0) {                                      // (pass 0: loaded before the set-up barrier)
      if (tid < nsub) {
        const int r0 = prow + tid * kBM;
        const int r1 = min(r0 + kBM, row_end) - 1;
        submask_s[tid] = __ldg(tile_mask + r0 / kBM) | __ldg(tile_mask + r1 / kBM);
      }
      __syncthreads();
    }
    if (warp == 0) {            // lane b = offset b: is it used by any sub-tile, by which ones; compacted in offset order
      unsigned jm = 0;
      for (int j = 0; j < nsub; ++j) jm |= ((submask_s[j] >> lane) & 1u) << j;
      const unsigned used = __ballot_sync(0xffffffffu, jm != 0u);
      const int pos = __popc(used & ((1u << lane) - 1u));
      const int cnt = __popc(used);
      if (jm != 0u) { klist_s[pos] = lane; jmask_s[pos] = jm; }
      if (lane >= cnt) { klist_s[lane] = 0; jmask_s[lane] = 1u; }      // padding entries (never walked, only prefetched)
      if (lane == 0) nk_s = cnt;
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
    // The walk: offsets outer, chunks, sub-tiles inner; (offset, sub-tile) pairs without any neighbour are skipped.  No integer
    // division and no shared-memory read per step (only per offset): a lone issuing warp cannot hide their latency
    // (measured: ~900 cycles per step with divisions in it).
    auto first = [&]() {
      G4It it;
      it.w = w_begin;
      it.kk = w_begin / nchunks;
      it.chunk = w_begin - it.kk * nchunks;
      it.k = klist_s[it.kk & 31];
      it.jm = jmask_s[it.kk & 31];
      it.rem = it.jm;
      it.j = __ffs(it.rem) - 1;
      return it;
    };
    auto advance = [&](G4It& it) {
      it.rem &= it.rem - 1;
      if (it.rem == 0u) {
        ++it.w;
        if (++it.chunk == nchunks) {
          it.chunk = 0;
          ++it.kk;
          it.k = klist_s[it.kk & 31];
          it.jm = jmask_s[it.kk & 31];
        }
        it.rem = it.jm;
      }
      it.j = __ffs(it.rem) - 1;
    };

    // producer index: warps 0-7 -> 0-7, warps 13-14 -> 8-9; producer pw copies rows [half * HROWS, +HROWS) of every stage of slot pslot
    const int pw = warp < kNPW ? warp : warp - (kNMW + 1);
    const bool is_prod = (warp < kNPW || warp > kNPW + kNMW) && pw < NPROD;
    if (is_prod) {
      // =========================== gather producers (cp.async) ===========================
      // Measured on B200 (tools/tma_rate.py, profiles/r01): TMA tile::gather4 of 128-byte rows tops out at ~44 B/cycle/SM with
      // 8 issuing warps (an instruction holds its warp ~2700 cycles) and drops to ~31 B/cycle inside this kernel; absent rows
      // cost it as much as present ones (more when zero-filled out of range).  The LSU path (LDGSTS, 16 B per lane) issues
      // 64 B/cycle/SM and its zero fill of an absent row moves no data, so the row gather uses cp.async; completion is still
      // tracked by the stage's mbarrier (cp.async.mbarrier.arrive.noinc), i.e. no thread ever waits for its own row copies.
      // A ring slot (= every NA-th stage) belongs to HALVES warps, each copying 128 / HALVES rows of the stage: their per-stage
      // barrier / walk / issue time (~1000 cycles for 64 instructions per lane) overlaps the other slots' copies, and with two
      // warps per slot the slot's turnaround fits the MMA rate.
      // KC = 64: 16 lanes per row (hi 8 x 16 B | lo 8 x 16 B): instruction i covers rows 4*(2*(i/4) + lane/16) + i%4;
      // KC = 32:  8 lanes per row: instruction i covers rows 4*(4*(i/4) + lane/8) + i%4  -> whole rows per instruction.
      constexpr int LPR = (KC == 64) ? 16 : 8;           // lanes per row
      constexpr int RPI = 32 / LPR;                      // rows per instruction
      constexpr int NINS = HROWS / RPI;                  // copy instructions per stage and lane
      const int pslot = pw % NA, half = pw / NA;
      const int cl = lane & (LPR - 1), hw = lane / LPR;
      const int part = (KC == 64) ? (cl >> 3) : 0, c16 = cl & 7;
      const uint32_t ring_base = tc::smem_u32(a_ring) + part * kImg;
      const char* xthr = reinterpret_cast<const char*>(X) + part * 128 + c16 * 16;
      const unsigned ldx_bytes = (unsigned)ldx * 2u;
      // The neighbour indices of a warp's NEXT stage are loaded into registers (plain 16-byte load) while it handles the current one;
      // the copy instructions fetch theirs from the owning lane by shuffle (no shared-memory staging: those 5 KB pay for a ring slot).
      // (They used to be prefetched with cp.async: the stage's cp.async.mbarrier.arrive.noinc then also waited for the prefetch
      // issued just before it -- a cold read of the 54 MB table -- so a slot took ~1300 cycles from "free" to "full" even with the
      // row copies disabled: clock64 trace, profiles/r02/experiments/call31_*.)
      auto load_idx = [&](const G4It& it) {              // lane l: neighbour rows of output rows 4l..4l+3 of this warp's half
        int4 v = make_int4(-1, -1, -1, -1);
        if (lane < HROWS / 4) {
          const int row = prow + it.j * kBM + half * HROWS + 4 * lane;
          if (row < row_end) v = __ldg(reinterpret_cast<const int4*>(nbr_t + (size_t)it.k * ld_n + row));
        }
        return v;
      };
      // Stage s lives in ring slot s % NA: the producers of a slot take part in every use of it, so none can run two phases
      // ahead of the slot's consumer (which the parity wait could not tell apart from "free").
      G4It it = first();
      const int first_i = pslot >= a_slot ? pslot - a_slot : pslot - a_slot + NA;  // (the ring may start a pass at any slot)
      for (int i = 0; i < first_i && it.w < w_end; ++i) advance(it);               // first stage of this warp
      int my_ac = ac + first_i;
      uint32_t my_phase = pslot >= a_slot ? a_phase : a_phase ^ 1u;                // phase of slot `pslot` at its next use
      // The swizzled shared-memory address of a copy is not recomputed per LDGSTS (that cost ~12 ALU instructions per copy and made
      // the producers issue-bound): row half*HROWS + 4*RPI*i4 + q of instruction group i4 (q = 4*hw + i) has the swizzled offset
      // i4 * (RPI/2) * 1024 + a per-lane constant, so four destination registers per lane are enough and the i4 term is an
      // immediate of the fully unrolled loop (4.25 instructions per LDGSTS in the SASS).
      uint32_t dstq[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) dstq[i] = ring_base + pslot * Cfg::A_BYTES + tc::sw128_offset(half * HROWS + 4 * hw + i, c16);
#define G4_DST(i4, i) (dstq[i] + (uint32_t)((i4) * (RPI / 2) * 1024))
#define G4_SRC(r) reinterpret_cast<const char*>(xb + (unsigned long long)((unsigned)max((r), 0)) * ldx_bytes)
      int4 cur_idx = make_int4(-1, -1, -1, -1);
      if (it.w < w_end) cur_idx = load_idx(it);
      int* my_idx = &idx_s[Cfg::IDX_SMEM ? pw : 0][0][0];
      int slot = 0;
      while (it.w < w_end) {
        G4It nxt = it;
        for (int i = 0; i < NA && nxt.w < w_end; ++i) advance(nxt);
        if (Cfg::IDX_SMEM && lane < HROWS / 4) *reinterpret_cast<int4*>(my_idx + slot * HROWS + 4 * lane) = cur_idx;
        int4 nxt_idx = make_int4(-1, -1, -1, -1);
        if (nxt.w < w_end) nxt_idx = load_idx(nxt);                       // in flight during this stage's wait and copies
        if (trace && lane == 0 && half == 0 && my_ac < 36) trace[16 + 4 * my_ac] = clock64();
        tc::mbar_wait(&empty_a[pslot], my_phase ^ 1u, err, 2);
        if (trace && lane == 0 && half == 0 && my_ac < 36) trace[17 + 4 * my_ac] = clock64();
        if (Cfg::IDX_SMEM) __syncwarp();     // the index slot written above is visible to the whole warp (double-buffered: no WAR)
        const int4* idx4p = reinterpret_cast<const int4*>(my_idx + slot * HROWS);
        unsigned long long xb = reinterpret_cast<unsigned long long>(xthr + it.chunk * (4 * KC));
        asm volatile("" : "+l"(xb));        // keep base + chunk offset in one register pair (one IMAD.WIDE per copy)
        if (!(dbg & 2)) {
#pragma unroll
          for (int i4 = 0; i4 < NINS / 4; ++i4) {
            const int m = RPI * i4 + hw;                                  // row group (of this warp's half) for these 4 instructions
            int4 r;
            if (Cfg::IDX_SMEM) {
              r = idx4p[m];
            } else {
              r = make_int4(__shfl_sync(0xffffffffu, cur_idx.x, m), __shfl_sync(0xffffffffu, cur_idx.y, m),
                            __shfl_sync(0xffffffffu, cur_idx.z, m), __shfl_sync(0xffffffffu, cur_idx.w, m));
            }
            const int r4[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)      // absent neighbour: the (valid) address of row 0 is passed but ignored (zero fill)
              g4_cp_async16_row(G4_DST(i4, i), G4_SRC(r4[i]), r4[i]);
          }
        }
#undef G4_DST
#undef G4_SRC
        g4_cp_async_arrive_noinc(&full_a[pslot]);                         // fires when this thread's copies have landed
        g4_cp_async_commit();
        my_ac += NA;
        my_phase ^= 1u;
        it = nxt;
        cur_idx = nxt_idx;
        slot ^= 1;
      }
      g4_cp_async_wait<0>();
    } else if (warp == kNPW + kNMW) {
      // =========================== weight-slab loader ===========================
      int kk = w_begin / nchunks, chunk = w_begin - kk * nchunks;
      for (int w = w_begin; w < w_end; ++w) {
        tc::mbar_wait(&empty_w[w_slot], w_phase ^ 1u, err, 1);
        if (lane == 0 && (dbg & 1)) tc::mbar_arrive(&full_w[w_slot]);
        if (lane == 0 && !(dbg & 1)) {
          tc::mbar_arrive_expect_tx(&full_w[w_slot], Cfg::W_BYTES);
          tc::bulk_g2s(w_ring + w_slot * Cfg::W_BYTES, Wp + (((size_t)klist_s[kk] * nchunks + chunk) * ntn + zt) * Cfg::W_BYTES,
                       Cfg::W_BYTES, &full_w[w_slot]);
        }
        if (++w_slot == kNW) { w_slot = 0; w_phase ^= 1u; }
        if (++chunk == nchunks) { chunk = 0; ++kk; }
      }
    } else if (warp >= kNPW && warp < kNPW + kNMW) {
      // =========================== MMA issuers ===========================
      // Three split products with two instructions per K step:
      //   D[:, 0:2BN] += a_hi . [Whi | Wlo]^T   (the weight slab is one 2BN-row image: rows [0,BN) = Whi, [BN,2BN) = Wlo)
      //   D[:, 0:BN]  += a_lo . Whi^T            the epilogue adds the two halves
      constexpr uint32_t idesc2 = g4_idesc_f16(kBM, 2 * BN), idesc1 = g4_idesc_f16(kBM, BN);
      const int mw = warp - kNPW;
      // A stage is issued by the MMA warp that owns its sub-tile (j % kNMW): every accumulator is fed by ONE thread, whose
      // tcgen05.mma execute in issue order.  Feeding an accumulator from several threads is not ordered by the tensor pipe even
      // when the issue order is (measured: last-bit differences from launch to launch, tools/conv_g4_check.py), so this
      // ownership is what makes the sums bit-reproducible.  With >= 2 sub-tiles consecutive stages fall to different warps, whose
      // per-stage bookkeeping (~600 cycles) then stays off the issue chain.  Hand-off PER RING SLOT (slot_turn_s): a warp waits on a
      // stage's full_a only after the previous use of the SAME slot has been seen full by its owner -- the barrier is then at most one
      // phase behind, which keeps the parity wait unambiguous (tests/test_g4_protocol_model.py shows what happens without it).  The
      // first form passed ONE turn through all stages in walk order: every stage then paid the hand-off chain (shared-memory poll,
      // barrier try_wait, turn store: ~380 cycles) serially, and with gathers, weight copies and MMAs disabled the 64 -> 64 launch
      // still took 171 of 327 us, the 32 -> 32 launch 138 of 187 us (profiles/r02/call28_conv_g4_bench_*).  Stages on different slots
      // now proceed independently; every accumulator still has ONE issuing thread, whose MMAs execute in issue order.  The
      // accumulators are zeroed here (each MMA warp clears one TMEM lane quadrant) so that every MMA accumulates.
      {
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        for (int col = 0; col < nsub * Cfg::ACC_COLS; col += 16) g4_tmem_zero16(tmem_d + lane_base + (uint32_t)col);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc::tc_fence_before_sync();
        named_barrier(2, kNMW * 32);
        tc::tc_fence_after_sync();
      }
      G4It cur = first();
      int last_w = -1, ws = 0;
      while (cur.w < w_end) {
        if (cur.w != last_w) {
          if (last_w >= 0 && tc::elect_one()) tc::mma_commit(&empty_w[ws]);      // this warp's MMAs on the previous slab (same elected thread)
          ws = w_slot;
          tc::mbar_wait(&full_w[ws], w_phase, err, 3);                     // every MMA warp waits: keeps them inside the slab ring
          if (++w_slot == kNW) { w_slot = 0; w_phase ^= 1u; }
          last_w = cur.w;
        }
        const int owner = (dbg & 8) ? 0 : (dbg & 16) ? (ac & (kNMW - 1)) : (cur.j & (kNMW - 1));
        if (owner == mw) {
          // The descriptors do not depend on the hand-off (warp-uniform values shuffled from lane 0, so the compiler keeps them in
          // uniform registers and emits bare UTCHMMA instructions): they are built first; the slot's turn is passed on as soon as
          // this stage's slot has been seen full, before issuing.
          const uint32_t a0 = __shfl_sync(0xffffffffu, tc::smem_u32(a_ring + a_slot * Cfg::A_BYTES), 0);
          const uint32_t w0 = __shfl_sync(0xffffffffu, tc::smem_u32(w_ring + ws * Cfg::W_BYTES), 0);
          const uint32_t d = __shfl_sync(0xffffffffu, tmem_d + (uint32_t)(cur.j * Cfg::ACC_COLS), 0);
          const uint64_t da = tc::smem_desc_sw128(a0), dw = tc::smem_desc_sw128(w0);
          while (*reinterpret_cast<volatile int*>(&slot_turn_s[a_slot]) != a_use) {}
          tc::mbar_wait(&full_a[a_slot], a_phase, err, 4);
          if (trace && lane == 0 && ac < 36) trace[18 + 4 * ac] = clock64();
          if (tc::elect_one()) {
            *reinterpret_cast<volatile int*>(&slot_turn_s[a_slot]) = a_use + 1;
            if (dbg & 4) {
            } else if (KC == 64) {
              // (descriptor start-address field = byte address >> 4: the hi image is at +0, the lo image at +kImg)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                g4_mma_f16(d, da + (uint64_t)(ks * 2), dw + (uint64_t)(ks * 2), idesc2, 1u);
                g4_mma_f16(d, da + (uint64_t)(kImg / 16 + ks * 2), dw + (uint64_t)(ks * 2), idesc1, 1u);
              }
            } else {
              // A row = [hi32 | lo32]; slab rows [0,BN) = [Whi | Whi], rows [BN,2BN) = [Wlo | 0]
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                g4_mma_f16(d, da + (uint64_t)(ks * 2), dw + (uint64_t)(ks * 2), idesc2, 1u);                  // hi . [Whi | Wlo]
                g4_mma_f16(d, da + (uint64_t)(4 + ks * 2), dw + (uint64_t)(4 + ks * 2), idesc1, 1u);          // lo . Whi
              }
            }
            tc::mma_commit(&empty_a[a_slot]);
          }
          if (trace && lane == 0 && ac < 36) trace[19 + 4 * ac] = clock64();
          __syncwarp();
        }
        ++ac;
        if (++a_slot == NA) { a_slot = 0; a_phase ^= 1u; ++a_use; }
        advance(cur);
      }
      if (tc::elect_one()) {                 // the thread that issued this warp's MMAs (elect.sync is deterministic per mask)
        if (last_w >= 0) tc::mma_commit(&empty_w[ws]);
        tc::mma_commit(&acc_bar);
        if (mw == 0) {                     // ring positions after this pass, for everybody (the producers skip through theirs)
          ring_s[0] = ac; ring_s[1] = a_slot; ring_s[2] = (int)a_phase; ring_s[3] = w_slot; ring_s[4] = (int)w_phase; ring_s[5] = a_use;
        }
      }
      __syncwarp();
    }
    if (warp < 8) {
      // =========================== epilogue (warps 0-7) ===========================
      const int q = warp & 3, h = warp >> 2;
      if (tid == 0) G4_TRACE(3);
      tc::mbar_wait(&acc_bar, (uint32_t)passes_done & 1u, err, 5);
      tc::tc_fence_after_sync();
      if (tid == 0) G4_TRACE(4);
      const unsigned started = 0xFFFFFFFFu;             // the MMA warps zero every accumulator before the first MMA
      constexpr int CW = BN / 2;
      const bool partial = (P != nullptr) && !part.row_mode && part.S > 1;
      bool big = false;
      // Residual rows: each thread reading its own row from global memory (32 rows = 32 cache lines per warp instruction, issued when
      // the sub-tile is due) cost block2_tr.conv2 63 us on top of conv1's 321 -- and prefetching the rows into the L2 changed nothing
      // (call 33: the cost is LSU wavefronts and exposed latency, not DRAM).  The ring is idle during the epilogue: the slots behind
      // the staging double buffer receive whole residual sub-tiles by coalesced cp.async (same swizzled image layout as the operands),
      // all of them issued when the accumulators are ready, completion per buffer on an mbarrier; a thread then reads its row's
      // channels with conflict-free 16-byte shared-memory loads.
      const bool res_smem = Cfg::NRB > 0 && R != nullptr && !partial && !out_row && (BN % kc_r) == 0 && !(dbg & 32);
      const int res_b0 = ((zt * BN) / kc_r) * 4 * kc_r;                  // byte offset of this CTA's channels in a residual row
      auto issue_res = [&](int j) {                                       // sub-tile j -> buffer j % NRB (256 threads, 16 B each per round)
        constexpr int PPR = BN / 4;                                       // 16-byte pieces per row ([hi | lo] chunks of BN channels)
        const uint32_t base = tc::smem_u32(a_ring) + (uint32_t)((2 + j % Cfg::NRB) * Cfg::OUT_BYTES);
        const char* rb = reinterpret_cast<const char*>(R) + res_b0;
#pragma unroll
        for (int i = 0; i < PPR / 2; ++i) {
          const int idx = tid + 256 * i, r = idx / PPR, pc = idx % PPR;
          const int gr = prow + j * kBM + r;
          const bool ok = gr < n;
          g4_cp_async16_row(base + (uint32_t)((pc >> 3) * kImg) + tc::sw128_offset(r, pc & 7),
                            rb + (size_t)(ok ? gr : 0) * ldr * 2 + pc * 16, ok ? 0 : -1);
        }
        g4_cp_async_arrive_noinc(&res_bar[j % Cfg::NRB]);
      };
      if (res_smem)
        for (int j = 0; j < nsub && j < Cfg::NRB; ++j) issue_res(j);
      for (int j = 0; j < nsub; ++j) {
        const int r_in = q * 32 + lane;
        const int grow = prow + j * kBM + r_in;
        unsigned char* stage = a_ring + (j & 1) * Cfg::OUT_BYTES;
        if (!partial && !out_row) {
          if (j >= 2 && lane == 0) tma::store_wait_read<1>();      // this thread's stores that read this buffer two sub-tiles ago
          named_barrier(1, 256);
        }
        if (tid == 0 && j < 8) G4_TRACE(160 + 8 * j);
        // NB 16-column blocks per round: the residual rows (global loads) and all TMEM loads of a round are in flight together
        constexpr int NB = CW >= 32 ? 2 : 1;
        const bool has_r = !partial && (R != nullptr) && grow < n;
        const float act_floor = relu ? 0.f : -INFINITY;
        uint32_t res_base = 0u;
        if (res_smem) {
          const int b = j % Cfg::NRB;
          tc::mbar_wait(&res_bar[b], (res_phase >> b) & 1u, err, 6);
          res_phase ^= 1u << b;
          res_base = tc::smem_u32(a_ring) + (uint32_t)((2 + b) * Cfg::OUT_BYTES);
        }
#pragma unroll 1
        for (int cb = 0; cb < CW; cb += 16 * NB) {
          float r16[NB][16];
          if (has_r && res_smem) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              const int cl = h * CW + cb + 16 * b;                       // 16 channels inside this CTA's BN-wide tile
              const int ph = ((cl / kc_r) * 4 * kc_r + 2 * (cl % kc_r)) >> 4, pl = ph + (kc_r >> 3);      // 16-byte pieces: hi, lo
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int4 h4 = g4_lds16(res_base + (uint32_t)(((ph + u) >> 3) * kImg) + tc::sw128_offset(r_in, (ph + u) & 7));
                const int4 l4 = g4_lds16(res_base + (uint32_t)(((pl + u) >> 3) * kImg) + tc::sw128_offset(r_in, (pl + u) & 7));
                g4_unpack8_h2(h4, l4, r16[b] + 8 * u);
              }
            }
          } else if (has_r) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              const int c = zt * BN + h * CW + cb + 16 * b;
              const __half* rp = R + (size_t)grow * ldr + (c / kc_r) * 2 * kc_r + (c % kc_r);
              g4_load16_h2(rp, rp + kc_r, r16[b]);
            }
          }
          uint32_t t1[NB][16], t2[NB][16];
          if ((started >> j) & 1u) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              const uint32_t col = (uint32_t)(j * Cfg::ACC_COLS + h * CW + cb + 16 * b);
              tc::tmem_ld16_issue(tmem_d + ((uint32_t)(q * 32) << 16) + col, t1[b]);
              tc::tmem_ld16_issue(tmem_d + ((uint32_t)(q * 32) << 16) + col + BN, t2[b]);
            }
            tc::tmem_ld_wait();
            if (tid == 0 && j < 8) G4_TRACE(161 + 8 * j);
          } else {
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
              for (int i = 0; i < 16; ++i) t1[b][i] = t2[b][i] = 0u;
          }
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            const int cl = h * CW + cb + 16 * b;                       // column inside this CTA's BN-wide tile
            const int c = zt * BN + cl;                                // absolute output channel of a[0]
            float a[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = __uint_as_float(t1[b][i]) + __uint_as_float(t2[b][i]);
            if (partial) {
              if (grow < n) {
                float4* dst = reinterpret_cast<float4*>(P + ((size_t)split * n + grow) * cout_total + c);
#pragma unroll
                for (int i = 0; i < 4; ++i) dst[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
              }
              continue;
            }
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 sc4 = *reinterpret_cast<const float4*>(&sc_s[cl + 4 * i4]);
              const float4 sh4 = *reinterpret_cast<const float4*>(&sh_s[cl + 4 * i4]);
              const float scv[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, shv[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float x = fmaf(a[4 * i4 + i], scv[i], shv[i]);
                if (has_r) x += r16[b][4 * i4 + i];
                a[4 * i4 + i] = fmaxf(x, act_floor);
              }
            }
            Half8 hi[2], lo[2];
            const float amax = g4_split16(a, hi, lo);
            big |= !(amax <= 60000.f) && (grow < n);
            if (out_row) {        // parity-grouped transposed convolution: table row grow describes output row out_row[grow]
              if (grow < n && grow < row_end) {      // rows past row_end belong to the next CTA's range
                __half* yp = Yp + (size_t)__ldg(out_row + grow) * ldy + (c / kc_out) * 2 * kc_out + (c % kc_out);
                tc::st_global_16(yp, hi[0]);
                tc::st_global_16(yp + 8, hi[1]);
                tc::st_global_16(yp + kc_out, lo[0]);
                tc::st_global_16(yp + kc_out + 8, lo[1]);
              }
              continue;
            }
            // staged layout: image = 64 halves of the h2 row; kc_out = 64: images (hi, lo) per 64 channels; 32: one image [hi32|lo32]
            int img_hi, img_lo, ch_hi, ch_lo;
            if (kc_out == 64) {
              img_hi = (cl >> 6) * 2; img_lo = img_hi + 1; ch_hi = (cl & 63) >> 3; ch_lo = ch_hi;
            } else {
              img_hi = cl >> 5; img_lo = img_hi; ch_hi = (cl & 31) >> 3; ch_lo = 4 + ch_hi;
            }
            unsigned char* ph = stage + img_hi * kImg;
            unsigned char* pl = stage + img_lo * kImg;
            tc::st_shared_16(ph + tc::sw128_offset(r_in, ch_hi), hi[0]);
            tc::st_shared_16(ph + tc::sw128_offset(r_in, ch_hi + 1), hi[1]);
            tc::st_shared_16(pl + tc::sw128_offset(r_in, ch_lo), lo[0]);
            tc::st_shared_16(pl + tc::sw128_offset(r_in, ch_lo + 1), lo[1]);
          }
        }
        if (tid == 0 && j < 8) G4_TRACE(162 + 8 * j);
        if (!partial && !out_row) {
          tc::fence_proxy_async();
          if (tid == 0 && j < 8) G4_TRACE(163 + 8 * j);
          named_barrier(1, 256);
          if (tid == 0 && j < 8) G4_TRACE(164 + 8 * j);
          if (res_smem && j + Cfg::NRB < nsub) issue_res(j + Cfg::NRB);      // (every thread is done reading buffer j % NRB)
          // warp (q, h) stores the 32-row box q of images h, h + 2, ...: eight issuing threads instead of one
          const int r0 = prow + j * kBM + q * 32;
          if (lane == 0) {
            if (r0 < row_end) {
#pragma unroll 1
              for (int img = h; img < BN / 32; img += 2)
                tma::store_2d(&tmY, tc::smem_u32(stage + img * kImg + q * 4096), zt * BN * 2 + img * 64, r0);
            }
            tma::store_commit();      // one (possibly empty) group per sub-tile: keeps the wait_read<1> arithmetic exact
          }
          if (tid == 0 && j < 8) G4_TRACE(165 + 8 * j);
        }
      }
      if (!partial && !out_row && lane == 0) tma::store_wait_read<0>();
      if (big && err) atomicOr(err, 0x10000);
      if (tid == 0) G4_TRACE(5);
    }
    tc::tc_fence_before_sync();
    __syncthreads();                       // pass boundary: TMEM drained, staging reads done, ring reusable
    tc::tc_fence_after_sync();
    ac = ring_s[0]; a_slot = ring_s[1]; a_phase = (uint32_t)ring_s[2]; w_slot = ring_s[3]; w_phase = (uint32_t)ring_s[4]; a_use = ring_s[5];
    ++passes_done;
  }
  if (warp == kNPW) tc::tmem_dealloc(tmem_d, tmem_cols);
  if (tid == 0) G4_TRACE(6);
  if (tid == 0 && cta_trace) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    cta_trace[bx] = clock64() - t_cta0;
    cta_trace[160 + bx] = (long long)smid;
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    cta_trace[320 + bx] = (long long)gt;
  }
#undef G4_TRACE
}

// Y(h2) = act( (sum_s P[s]) * scale + shift (+ R) ) for the levels that ran in split mode with S > 1; one thread per (row, 16 ch).
__global__ void __launch_bounds__(256) k_conv_g4_reduce(const float* __restrict__ P, const int* __restrict__ n_ptr, int n_max, int gx,
                                                        int nst_max, int Cout, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, const __half* __restrict__ R, int ldr, int kc_r,
                                                        int relu, __half* __restrict__ Y, int ldy, int kc_out, int* err,
                                                        const int* __restrict__ out_row, int sps) {
  int n = n_max;
  if (n_ptr) { const int v = *n_ptr; n = v < n_max ? v : n_max; }
  const G4Part part = g4_partition(n, gx, nst_max, true, sps);
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
  const bool big = !(g4_split16(a, hi, lo) <= 60000.f);
  __half* yp = Y + (size_t)(out_row ? __ldg(out_row + row) : row) * ldy + (c / kc_out) * 2 * kc_out + (c % kc_out);
  tc::st_global_16(yp, hi[0]);
  tc::st_global_16(yp + 8, hi[1]);
  tc::st_global_16(yp + kc_out, lo[0]);
  tc::st_global_16(yp + kc_out + 8, lo[1]);
  if (big && err) atomicOr(err, 0x10000);
}

long long* g_g4_trace = nullptr;
int g_g4_dbg = 0;       // profiling hook: bit 0 skips the weight copies, bit 1 the gathers, bit 2 the MMAs (results meaningless);
                        // bits 3/4 change the MMA warp that owns a stage; bit 5 (32): the epilogue reads residual rows straight from global memory
int g_g4_sps = 4;         // stages per split CTA at least (profiling hook can change it)
int g_g4_grid = 0;        // profiling hook: overrides the number of CTAs per output-channel tile (0 = one per SM)
int g4_split_param() {    // experiment knob: IMF_G4_MAX_SPLITS caps the split factor of the small levels (default 32)
  static const int cap = [] { const char* e = getenv("IMF_G4_MAX_SPLITS"); const int v = e ? atoi(e) : 0; return v > 0 && v < 32 ? v : 0; }();
  return g_g4_sps | (cap << 16);
}

template <int BN, int KC>
int launch_g4(const __half* X, int ldx, const CUtensorMap& tmY, const void* Wp, const int* nbr_t, int ld_n, const unsigned* tile_mask,
              const int* n_ptr, int n_max, int K3, int Cin, int Cout, const float* scale, const float* shift, const __half* R, int ldr,
              int kc_r, int relu, __half* Y, int ldy, int kc_out, const int* out_row, void* ws, size_t ws_bytes, int* err,
              cudaStream_t stream) {
  using Cfg = G4Cfg<BN, KC>;
  const size_t smem = (size_t)Cfg::RING_BYTES + (size_t)kNW * Cfg::W_BYTES + 1024;
  IMF_CHECK_CUDA(imf_set_max_smem_once(reinterpret_cast<const void*>(&k_sparse_conv_g4<BN, KC>), (int)smem));
  const int ntn = Cout / BN;
  const int nchunks = Cin / KC;
  // one CTA per SM and output-channel tile, whatever n_max: the row partition (hence the fp32 summation order of the split mode)
  // then depends only on the actual row count, so a fragment gives the same bits in an exact-size and in a bucketed launch
  const int sms = imf_sm_count();
  const int gx = (g_g4_grid > 0 ? g_g4_grid : sms) / ntn;
  const int nst_max = K3 * nchunks;
  float* P = nullptr;
  if (ws != nullptr && ws_bytes >= (size_t)sms * kBM * Cout * sizeof(float)) P = reinterpret_cast<float*>(ws);
  dim3 grid(gx, 1, ntn);
  k_sparse_conv_g4<BN, KC><<<grid, Cfg::THREADS, smem, stream>>>(X, ldx, tmY, reinterpret_cast<const unsigned char*>(Wp), nbr_t, ld_n, tile_mask,
                                                            n_ptr, n_max, K3, nchunks, scale, shift, R, ldr, kc_r, relu, kc_out, P, Cout,
                                                            out_row, Y, ldy, err, g_g4_trace, g_g4_dbg, g4_split_param());
  IMF_CHECK_LAUNCH();
  if (P != nullptr) {      // split mode is possible for small n: the reduce kernel decides on the device (no-op otherwise)
    const int rows = n_max < gx * kBM ? n_max : gx * kBM;      // split mode only exists below gx tiles
    const long long total = (long long)rows * (Cout / 16);
    k_conv_g4_reduce<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(P, n_ptr, n_max, gx, nst_max, Cout, scale, shift, R, ldr, kc_r,
                                                                         relu, Y, ldy, kc_out, err, out_row, g4_split_param());
    IMF_CHECK_LAUNCH();
  }
  return IMF_OK;
}

}  // namespace

extern "C" int imf_debug_conv_g4_trace(long long* trace, int32_t grid, int32_t producer_warps, int32_t flags) {
  g_g4_trace = trace;
  g_g4_dbg = flags;
  g_g4_grid = grid;
  if (producer_warps > 0) g_g4_sps = producer_warps;      // (argument re-used: minimum stages per split CTA)
  return IMF_OK;
}

extern "C" size_t imf_sparse_conv_g4_workspace_bytes(int32_t Cout) { return (size_t)imf_sm_count() * kBM * (size_t)Cout * sizeof(float); }

extern "C" int imf_sparse_conv_g4_fwd_perm(const void* X, int32_t ldx, int32_t kc_in, const void* packed, const int32_t* nbr_t,
                                           int32_t ld_n, const uint32_t* tile_mask, const int32_t* n_out_dev, int32_t n_out_max,
                                           int32_t kernel_volume, int32_t Cin, int32_t Cout, const float* scale, const float* shift,
                                           const void* residual, int32_t ldr, int32_t kc_r, int32_t relu, void* Y, int32_t ldy,
                                           int32_t n_y_rows, int32_t kc_out, const int32_t* out_row, void* workspace,
                                           size_t workspace_bytes, int32_t* err, cudaStream_t stream);

extern "C" int imf_sparse_conv_g4_fwd(const void* X, int32_t ldx, int32_t kc_in, const void* packed, const int32_t* nbr_t,
                                      int32_t ld_n, const uint32_t* tile_mask, const int32_t* n_out_dev, int32_t n_out_max,
                                      int32_t kernel_volume, int32_t Cin, int32_t Cout, const float* scale, const float* shift,
                                      const void* residual, int32_t ldr, int32_t kc_r, int32_t relu, void* Y, int32_t ldy,
                                      int32_t n_y_rows, int32_t kc_out, void* workspace, size_t workspace_bytes, int32_t* err,
                                      cudaStream_t stream) {
  return imf_sparse_conv_g4_fwd_perm(X, ldx, kc_in, packed, nbr_t, ld_n, tile_mask, n_out_dev, n_out_max, kernel_volume, Cin, Cout, scale,
                                     shift, residual, ldr, kc_r, relu, Y, ldy, n_y_rows, kc_out, nullptr, workspace, workspace_bytes, err,
                                     stream);
}

// Same with a row permutation: table row v (and tile masks) describe output row out_row[v]; the result of table row v is written
// to Y[out_row[v]].  Used with imf_parity_perm for transposed convolutions, whose 128-row tiles then walk 1-8 offsets instead of 27.
// residual (if any) is read at table-row order, so it must be NULL unless it is permuted the same way.
extern "C" int imf_sparse_conv_g4_fwd_perm(const void* X, int32_t ldx, int32_t kc_in, const void* packed, const int32_t* nbr_t,
                                           int32_t ld_n, const uint32_t* tile_mask, const int32_t* n_out_dev, int32_t n_out_max,
                                           int32_t kernel_volume, int32_t Cin, int32_t Cout, const float* scale, const float* shift,
                                           const void* residual, int32_t ldr, int32_t kc_r, int32_t relu, void* Y, int32_t ldy,
                                           int32_t n_y_rows, int32_t kc_out, const int32_t* out_row, void* workspace,
                                           size_t workspace_bytes, int32_t* err, cudaStream_t stream) {
  IMF_CHECK_ARG(out_row == nullptr || residual == nullptr);
  IMF_CHECK_ARG(n_out_max >= 0 && kernel_volume >= 1 && kernel_volume <= 27);
  IMF_CHECK_ARG((kc_in == 32 || kc_in == 64) && Cin > 0 && Cin % kc_in == 0 && (Cout == 32 || Cout == 64 || Cout == 128 || Cout == 256));
  IMF_CHECK_ARG((kc_out == 32 || kc_out == 64) && Cout % kc_out == 0);
  IMF_CHECK_ARG(scale != nullptr && shift != nullptr);
  IMF_CHECK_ARG(ldx % 8 == 0 && ldx >= 2 * Cin && ldy % 8 == 0 && ldy >= 2 * Cout && n_y_rows >= n_out_max);
  IMF_CHECK_ARG(residual == nullptr || ((kc_r == 32 || kc_r == 64) && Cout % kc_r == 0 && ldr % 8 == 0 && ldr >= 2 * Cout));
  IMF_CHECK_ARG(ld_n % 4 == 0 && ld_n >= ((n_out_max + 31) & ~31));
  if (n_out_max == 0) return IMF_OK;
  IMF_CHECK_ARG(X != nullptr && packed != nullptr && nbr_t != nullptr && tile_mask != nullptr && Y != nullptr);
  IMF_CHECK_ARG(((uintptr_t)X % 16) == 0 && ((uintptr_t)packed % 16) == 0 && ((uintptr_t)Y % 16) == 0 && ((uintptr_t)residual % 16) == 0 &&
                ((uintptr_t)nbr_t % 16) == 0);
  CUtensorMap tmY;
  int rc = tma::encode_2d_u16(&tmY, Y, (uint64_t)n_y_rows, (uint64_t)(2 * Cout), (uint64_t)ldy, 64, 32);
  if (rc) { imf_set_error("cuTensorMapEncodeTiled(Y) failed: %d", rc); return IMF_ERR_CUDA; }
  const __half* Rh = reinterpret_cast<const __half*>(residual);
  __half* Yh = reinterpret_cast<__half*>(Y);
#define IMF_GO(BN, KC)                                                                                                               \
  return launch_g4<BN, KC>(reinterpret_cast<const __half*>(X), ldx, tmY, packed, nbr_t, ld_n, tile_mask, n_out_dev, n_out_max, kernel_volume, Cin, Cout, scale, shift, Rh, \
                           ldr, kc_r, relu, Yh, ldy, kc_out, out_row, workspace, workspace_bytes, err, stream)
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
