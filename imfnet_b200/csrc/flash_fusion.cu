// imfnet_b200 -- attention of the fusion module as ONE persistent tcgen05 kernel for ALL batch items:
//   S = q k^T, softmax, o = softmax(S) v
//   /root/reference/model/attention_fusion.py:84-93 (einsum 'b i d, b j d -> b i j', softmax(dim=-1), einsum 'b i j, b j d -> b i d'),
// one cross head of 128 channels; per batch item b its M_b point tokens (queries, rows [seg[b], seg[b] + cnt[b]) of the stride-8
// level, model/resunet.py:240-271) against the L image tokens of image b (keys / values).
//
// Operands are fp16 hi/lo pairs ("h2", see h2_format.cu): q, K and V^T are pre-split once, P = exp2(S - m) is split on the fly; every
// product is accumulated as hi.hi + hi.lo + lo.hi with two tcgen05.mma (kind::f16) per K step by concatenating [hi ; lo] of the B
// operand along N, exactly as the sparse convolution does.  All operand tiles are dense, so they are fetched with tiled TMA loads
// (cp.async.bulk.tensor.2d, 128-byte swizzle) -- the [M, L] score matrix never exists in memory.
//
// Work decomposition ("stream-K"): the (query tile, 64-token block) pairs of all items form one flat axis, tile-major; every CTA of
// the persistent grid takes one contiguous range of `per` blocks, whatever tiles it crosses.  A (CTA, tile) intersection is a PIECE:
// its un-normalised output, row maxima and row sums go to slot `tile + cta` of the workspace and k_flash_combine merges the pieces of
// a tile (log-sum-exp weights).  Every CTA therefore does the same number of blocks (no wave quantisation, no per-item launches),
// and all sizes are read on the device (seg / cnt), which keeps the launch capturable.
//
// CTA = 19 warps:
//   warp 0    TMA producer of Q and K: Q of the piece, then the K blocks (pass 1: their hi halves only) into a 2-stage ring
//   warp 18   TMA producer of V (pass 2) into its own 2-stage ring.  K and V were ONE stage at first: a stage was then released only
//             by the P.V MMAs of its block, so the loads of block j + 2 started after P.V(j) and S(j + 2) waited for their whole
//             latency (ncu, 8192 x 4800: the softmax warps waited 43 % of the time for S, ~8 k cycles per block for ~2 k of MMAs);
//             with separate rings K(j + 2) is fetched as soon as S(j) has completed, V(j + 2) as soon as P.V(j) has
//   warp 1    MMA issuer + TMEM owner: S double-buffered (2 x 128 columns), O (256 columns)
//   warps 2-17 softmax, four warps per TMEM lane quadrant (16 of a block's 64 tokens each): pass 1 finds the row maximum of the piece
//             from the CHEAP product q_hi . k_hi^T (any m close to the maximum serves: it only has to keep exp2(S - m) in range, and
//             the pieces are merged with exact weights exp2(m_piece - m)); pass 2 recomputes S with all three products, writes
//             P = exp2(S - m) (hi/lo) to shared memory for the P.V MMAs and sums the row -- O is never rescaled.
//             (P as a tensor-memory A operand, written in place over S, was tried in calls 36-51: parity tests green, but the bench's
//             streaming steps then failed intermittently -- S(j + 2) is issued right behind P.V(j) into the columns P.V(j) still reads
//             its A operand from, and issue order alone does not protect that read; profiles/r02/experiments/call36-54_attention.txt.)
//             While the softmax warps work on block j the tensor pipe already computes S of block j + 1 (second S buffer).
// The queries arrive scaled by log2(e) / sqrt(d) (dense.cu), so the exponentials are bare ex2.approx.
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"

namespace {

constexpr int kD = 128;                 // head dimension
constexpr int kTQ = 128;                // queries per tile
constexpr int kTB = 64;                 // tokens per block
constexpr int kQImg = kTQ * 128;        // 16 KB: 128 rows x 64 halves
constexpr int kKImg = kTB * 128;        //  8 KB:  64 rows x 64 halves
constexpr int kVImg = kD * 128;         // 16 KB: 128 dims x 64 tokens
constexpr int kQBytes = 4 * kQImg;      // hi0 lo0 hi1 lo1
constexpr int kKBytes = 4 * kKImg;
constexpr int kVBytes = 2 * kVImg;      // Vhi Vlo
constexpr int kPBytes = 2 * kQImg;      // Phi Plo (128 rows x 64 tokens)
constexpr int kSmem = kQBytes + 2 * kKBytes + 2 * kVBytes + kPBytes;
constexpr int kThreads = 608;
constexpr int kSoftWarps = 16;
constexpr int kMaxPieces = 16;          // pieces a CTA may hold (the host sizes the grid so that this suffices)
constexpr int kMaxItems = 256;

__host__ __device__ constexpr uint32_t ff_idesc(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void ff_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void ff_tma_load(uint32_t dst, const CUtensorMap* map, uint32_t bar, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(col), "r"(row)
               : "memory");
}
__device__ __forceinline__ float ff_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ff_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

struct __align__(16) FHalf8 { __half2 a, b, c, d; };

#ifdef IMF_FF_TRACE      // clock64 timeline of CTA 0 (profiling build only: IMFNET_B200_NVCC_FLAGS=-DIMF_FF_TRACE, tools/flash_trace.py)
__device__ long long g_ff_trace[8 * 256];
#define FF_TRACE(cond, idx, slot) do { if (blockIdx.x == 0 && (cond) && (idx) < 256) g_ff_trace[(idx) * 8 + (slot)] = clock64(); } while (0)
#else
#define FF_TRACE(cond, idx, slot) do { } while (0)
#endif

struct FFPiece {
  int item;        // batch item
  int row0;        // first query row of the tile (global row of the level matrix)
  int row_end;     // end of the item's rows
  int blk0;        // first 64-token block of the piece
  int nblk;        // blocks of the piece
  int slot;        // workspace slot of its partial result
  int pad[2];
};

// K [B*L, ldk] fp32 (columns 0..127 of row b*L + t) -> h2 [B*Lpad, 256 halves] (chunk width 64); V [B*L, ldv] -> V^T as h2 over the token
// axis [B*128, 2*Lpad halves]; padding tokens = 0.  v_transposed != 0: V is given as V^T [128, ldv] (single item, the old kv layout).
// grid = (Lpad / 64, B, 2): z = 0 packs a 64-token block of K (two channels per thread: 8-byte loads, half2 stores), z = 1 transposes
// the block of V through shared memory (coalesced token rows in, one h2 chunk [hi 64 | lo 64] per dimension out).  The first form used
// one thread per element: 2-byte stores everywhere and, for V, a read stride of a whole token row per thread (56 us per batch of ten
// images, profiles/r02/call43_ncu_other_kernels.txt).
__global__ void __launch_bounds__(256) k_flash_pack_kv(const float* __restrict__ K, int ldk, const float* __restrict__ V, int ldv,
                                                       int v_transposed, int L, int Lpad, int B, __half* __restrict__ Kh,
                                                       __half* __restrict__ Vh) {
  __shared__ float tile[64][kD + 1];
  const int t0 = blockIdx.x * kTB, b = blockIdx.y, tid = threadIdx.x;
  if (blockIdx.z == 0) {
    const bool vec = ((reinterpret_cast<uintptr_t>(K) & 7) == 0) && (ldk & 1) == 0;
#pragma unroll 4
    for (int e = tid; e < kTB * (kD / 2); e += 256) {          // (token, channel pair)
      const int r = e / (kD / 2), c = 2 * (e % (kD / 2));
      const int t = t0 + r;
      float2 v = make_float2(0.f, 0.f);
      if (t < L) {
        const float* src = K + ((size_t)b * L + t) * ldk + c;
        v = vec ? *reinterpret_cast<const float2*>(src) : make_float2(src[0], src[1]);
      }
      const __half2 h = __floats2half2_rn(v.x, v.y);
      const float2 hf = __half22float2(h);
      __half* p = Kh + ((size_t)b * Lpad + t) * (2 * kD) + (c >> 6) * 128 + (c & 63);
      *reinterpret_cast<__half2*>(p) = h;
      *reinterpret_cast<__half2*>(p + 64) = __floats2half2_rn(v.x - hf.x, v.y - hf.y);
    }
    return;
  }
  // V: tile[token][dim]
  for (int e = tid; e < kTB * kD; e += 256) {
    const int r = v_transposed ? e % kTB : e / kD, d = v_transposed ? e / kTB : e % kD;
    const int t = t0 + r;
    float v = 0.f;
    if (t < L) v = v_transposed ? V[(size_t)d * ldv + t] : V[((size_t)b * L + t) * ldv + d];
    tile[r][d] = v;
  }
  __syncthreads();
#pragma unroll 4
  for (int e = tid; e < kD * (kTB / 2); e += 256) {            // (dim, token pair)
    const int d = e / (kTB / 2), r = 2 * (e % (kTB / 2));
    const float v0 = tile[r][d], v1 = tile[r + 1][d];
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(h);
    __half* p = Vh + ((size_t)b * kD + d) * (2 * Lpad) + (size_t)blockIdx.x * 128 + r;
    *reinterpret_cast<__half2*>(p) = h;
    *reinterpret_cast<__half2*>(p + 64) = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
  }
}

// rows of item b that lie inside the matrix (an item whose range runs past M_max -- a plan whose token capacity was exceeded; the
// host then discards the result -- is clipped, so nothing is ever read or written outside the M_max rows);
// cnt == nullptr: one item with *m_ptr (or M_max) rows starting at row 0
__device__ __forceinline__ int ff_item_rows(const int* seg, const int* cnt, const int* m_ptr, int M_max, int b) {
  int v, s0 = 0;
  if (cnt) { v = cnt[b]; s0 = seg[b]; }
  else v = m_ptr ? *m_ptr : M_max;
  if (s0 < 0 || s0 >= M_max) return 0;
  if (v > M_max - s0) v = M_max - s0;
  return v < 0 ? 0 : v;
}

__global__ void __launch_bounds__(kThreads, 1)
k_flash_fusion(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
               const int* __restrict__ seg, const int* __restrict__ cnt, const int* __restrict__ m_ptr, int B, int M_max, int L, int Lpad,
               float* __restrict__ Opart, float* __restrict__ ml, int* err) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* q_s = smem;
  unsigned char* kring = smem + kQBytes;
  unsigned char* vring = kring + 2 * kKBytes;
  unsigned char* p_s = vring + 2 * kVBytes;
  __shared__ __align__(8) uint64_t q_full, kfull[2], kempty[2], vfull[2], vempty[2], s_ready[2], s_free[2], p_ready, p_free, o_done, o_free;
  __shared__ uint32_t tmem_base_s;
  __shared__ FFPiece piece_s[kMaxPieces];
  __shared__ int npiece_s;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int nb = (L + kTB - 1) / kTB;                 // blocks per tile (every image of the batch has L tokens)

  if (tid == 0) {
    // ---- this CTA's pieces: flat blocks [c * per, (c + 1) * per) of the tile-major (tile, block) axis ----
    long long T = 0;
    for (int b = 0; b < B; ++b) T += (ff_item_rows(seg, cnt, m_ptr, M_max, b) + kTQ - 1) / kTQ;
    const long long total = T * nb;
    const long long per = (total + gridDim.x - 1) / gridDim.x;
    long long f0 = per * blockIdx.x;
    const long long f1 = f0 + per < total ? f0 + per : total;
    int np = 0;
    if (per > 0 && f0 < f1) {
      long long tile = f0 / nb;
      int b = 0;
      long long tb = 0;                                // first tile of item b
      for (;;) {                                       // the item of `tile`
        const int tiles_b = (ff_item_rows(seg, cnt, m_ptr, M_max, b) + kTQ - 1) / kTQ;
        if (tile < tb + tiles_b) break;
        tb += tiles_b;
        ++b;
      }
      while (f0 < f1 && np < kMaxPieces) {
        const int rows_b = ff_item_rows(seg, cnt, m_ptr, M_max, b);
        const int s0 = seg ? seg[b] : 0;
        const long long tend = (tile + 1) * nb;
        const long long pe = tend < f1 ? tend : f1;
        FFPiece& p = piece_s[np++];
        p.item = b;
        p.row0 = s0 + (int)(tile - tb) * kTQ;
        p.row_end = s0 + rows_b;
        p.blk0 = (int)(f0 - tile * nb);
        p.nblk = (int)(pe - f0);
        p.slot = (int)tile + (int)blockIdx.x;
        f0 = pe;
        if (f0 == tend) {
          ++tile;
          while (b < B && tile >= tb + (ff_item_rows(seg, cnt, m_ptr, M_max, b) + kTQ - 1) / kTQ) {
            tb += (ff_item_rows(seg, cnt, m_ptr, M_max, b) + kTQ - 1) / kTQ;
            ++b;
          }
        }
      }
      if (f0 < f1 && err) atomicExch(err, 9);          // more pieces than kMaxPieces: the host sized the grid wrongly
    }
    npiece_s = np;
    tc::mbar_init(&q_full, 1);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&kfull[s], 1);
      tc::mbar_init(&kempty[s], 1);
      tc::mbar_init(&vfull[s], 1);
      tc::mbar_init(&vempty[s], 1);
      tc::mbar_init(&s_ready[s], 1);
      tc::mbar_init(&s_free[s], kSoftWarps);
    }
    tc::mbar_init(&p_ready, kSoftWarps);
    tc::mbar_init(&p_free, 1);
    tc::mbar_init(&o_done, 1);
    tc::mbar_init(&o_free, kSoftWarps);
    tc::fence_barrier_init();
    tma::prefetch_map(&tmQ);
    tma::prefetch_map(&tmK);
    tma::prefetch_map(&tmV);
  }
  __syncthreads();
  const int npiece = npiece_s;
  if (npiece == 0) return;
  if (warp == 1) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_s = tmem_base_s;            // S buffers: columns [0,128) and [128,256)   (each S1 | S2)
  const uint32_t tmem_o = tmem_base_s + 256u;     // O: columns [256,512) (O1 | O2)

  if (warp == 0) {
    // =========================== TMA producer: Q and K ===========================
    if (lane == 0) {
      uint32_t it = 0;
      for (int pc = 0; pc < npiece; ++pc) {
        const FFPiece p = piece_s[pc];
        if (pc > 0) tc::mbar_wait(&o_done, (uint32_t)(pc - 1) & 1u, err, 1);      // the previous piece's MMAs are done with Q
        tc::mbar_arrive_expect_tx(&q_full, kQBytes);
        for (int i = 0; i < 4; ++i) ff_tma_load(tc::smem_u32(q_s + i * kQImg), &tmQ, tc::smem_u32(&q_full), i * 64, p.row0);
        const int krow = p.item * Lpad;
        for (int pass = 0; pass < 2; ++pass) {
          // pass 1 (q_hi . k_hi^T only) takes the blocks in PAIRS: the hi images of blocks a and a + 1 fill one ring stage in the place
          // of [K_hi ; K_lo] of one block, so a pair is ONE N = 128 product, one barrier round trip and one TMEM read instead of two
          for (int j = 0; j < p.nblk; j += (pass == 0 ? 2 : 1), ++it) {
            const int st = it & 1;
            tc::mbar_wait(&kempty[st], ((it >> 1) & 1u) ^ 1u, err, 2);
            unsigned char* kd = kring + st * kKBytes;
            const int t0 = krow + (p.blk0 + j) * kTB;
            if (pass == 0) {        // hi images of both channel chunks (columns 0 and 128 of the h2 rows), of one or two blocks
              const bool pair = j + 1 < p.nblk;
              tc::mbar_arrive_expect_tx(&kfull[st], (pair ? 4 : 2) * kKImg);
              ff_tma_load(tc::smem_u32(kd), &tmK, tc::smem_u32(&kfull[st]), 0, t0);
              ff_tma_load(tc::smem_u32(kd + 2 * kKImg), &tmK, tc::smem_u32(&kfull[st]), 128, t0);
              if (pair) {
                ff_tma_load(tc::smem_u32(kd + kKImg), &tmK, tc::smem_u32(&kfull[st]), 0, t0 + kTB);
                ff_tma_load(tc::smem_u32(kd + 3 * kKImg), &tmK, tc::smem_u32(&kfull[st]), 128, t0 + kTB);
              }
            } else {
              tc::mbar_arrive_expect_tx(&kfull[st], kKBytes);
              for (int i = 0; i < 4; ++i) ff_tma_load(tc::smem_u32(kd + i * kKImg), &tmK, tc::smem_u32(&kfull[st]), i * 64, t0);
            }
          }
        }
      }
    }
  } else if (warp == 2 + kSoftWarps) {
    // =========================== TMA producer: V (pass 2 only) ===========================
    if (lane == 0) {
      uint32_t it = 0;
      for (int pc = 0; pc < npiece; ++pc) {
        const FFPiece p = piece_s[pc];
        const int vrow = p.item * kD;
        for (int j = 0; j < p.nblk; ++j, ++it) {
          const int st = it & 1;
          tc::mbar_wait(&vempty[st], ((it >> 1) & 1u) ^ 1u, err, 11);
          unsigned char* vd = vring + st * kVBytes;
          const int blk = p.blk0 + j;
          tc::mbar_arrive_expect_tx(&vfull[st], kVBytes);
          ff_tma_load(tc::smem_u32(vd), &tmV, tc::smem_u32(&vfull[st]), blk * 128, vrow);
          ff_tma_load(tc::smem_u32(vd + kVImg), &tmV, tc::smem_u32(&vfull[st]), blk * 128 + 64, vrow);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // Warp-uniform issue (operand addresses shuffled from lane 0, MMAs under elect.sync): the compiler then keeps the descriptors in
    // uniform registers and emits bare UTCHMMA instead of an ELECT / R2UR / branch loop around every instruction.
    constexpr uint32_t id_s2 = ff_idesc(kTQ, 2 * kTB), id_s1 = ff_idesc(kTQ, kTB);       // N = 128 / 64
    constexpr uint32_t id_o2 = ff_idesc(kTQ, 2 * kD), id_o1 = ff_idesc(kTQ, kD);          // N = 256 / 128
    const uint32_t q0 = __shfl_sync(0xffffffffu, tc::smem_u32(q_s), 0);
    const uint32_t kr0 = __shfl_sync(0xffffffffu, tc::smem_u32(kring), 0);
    const uint32_t vr0 = __shfl_sync(0xffffffffu, tc::smem_u32(vring), 0);
    const uint32_t p0 = __shfl_sync(0xffffffffu, tc::smem_u32(p_s), 0);
    const uint32_t ts = __shfl_sync(0xffffffffu, tmem_s, 0), to = __shfl_sync(0xffffffffu, tmem_o, 0);
    uint32_t it = 0;       // K ring uses consumed
    uint32_t vit = 0;      // V ring uses consumed
    uint32_t sit = 0;      // S buffers issued
    uint32_t pit = 0;      // P blocks consumed
    uint32_t pbuf = 0;     // running S count of the block whose P.V product is issued next (trace index)
    auto issue_s = [&](uint32_t st, bool full_product, bool pair) {      // pair (pass 1): the stage holds K_hi of two blocks -> N = 128
      const uint32_t sb = sit & 1u;
      tc::mbar_wait(&s_free[sb], ((sit >> 1) & 1u) ^ 1u, err, 4);            // softmax warps have read this S buffer's previous content
      tc::tc_fence_after_sync();
      FF_TRACE(lane == 0, sit, 0);
      const uint32_t k0 = kr0 + st * kKBytes;
      const uint32_t d = ts + sb * 128u;
      if (tc::elect_one()) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t qhi = q0 + (2 * c) * kQImg, qlo = qhi + kQImg, kb = k0 + (2 * c) * kKImg;     // [Khi_c ; Klo_c] = 128 rows
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t o = ks * 32;
            if (full_product) {
              ff_mma(d, tc::smem_desc_sw128(qhi + o), tc::smem_desc_sw128(kb + o), id_s2, (c | ks) ? 1u : 0u);
              ff_mma(d, tc::smem_desc_sw128(qlo + o), tc::smem_desc_sw128(kb + o), id_s1, 1u);
            } else {
              ff_mma(d, tc::smem_desc_sw128(qhi + o), tc::smem_desc_sw128(kb + o), pair ? id_s2 : id_s1, (c | ks) ? 1u : 0u);
            }
          }
        }
        tc::mma_commit(&s_ready[sb]);
        tc::mma_commit(&kempty[st]);              // the K block is only needed by this product
      }
      __syncwarp();
      FF_TRACE(lane == 0, sit, 1);
      ++sit;
    };
    for (int pc = 0; pc < npiece; ++pc) {
      const int nblk = piece_s[pc].nblk;
      tc::mbar_wait(&q_full, (uint32_t)pc & 1u, err, 3);
      // ---- pass 1: approximate scores (hi . hi) for the row maxima ----
      for (int j = 0; j < nblk; j += 2, ++it) {
        const uint32_t st = it & 1u;
        tc::mbar_wait(&kfull[st], (it >> 1) & 1u, err, 5);
        issue_s(st, false, j + 1 < nblk);
      }
      // ---- pass 2: S(j + 1) is issued before P(j) . V(j), so the tensor pipe works while the softmax warps handle block j ----
      tc::mbar_wait(&o_free, ((uint32_t)pc & 1u) ^ 1u, err, 6);                // the previous piece's O has been drained
      tc::mbar_wait(&kfull[it & 1u], (it >> 1) & 1u, err, 5);
      pbuf = sit;
      issue_s(it & 1u, true, false);
      ++it;
      for (int j = 0; j < nblk; ++j, ++vit, ++pbuf) {
        if (j + 1 < nblk) {
          tc::mbar_wait(&kfull[it & 1u], (it >> 1) & 1u, err, 5);
          issue_s(it & 1u, true, false);
          ++it;
        }
        const uint32_t st = vit & 1u;
        tc::mbar_wait(&p_ready, pit & 1u, err, 7);                              // P of block j is in shared memory
        tc::mbar_wait(&vfull[st], (vit >> 1) & 1u, err, 12);
        tc::tc_fence_after_sync();
        FF_TRACE(lane == 0, pbuf, 2);
        const uint32_t v0 = vr0 + st * kVBytes;                                // [Vhi ; Vlo] = 256 rows
        const uint32_t jj = (uint32_t)j;
        if (tc::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t o = ks * 32;
            ff_mma(to, tc::smem_desc_sw128(p0 + o), tc::smem_desc_sw128(v0 + o), id_o2, (jj | (uint32_t)ks) ? 1u : 0u);
            ff_mma(to, tc::smem_desc_sw128(p0 + kQImg + o), tc::smem_desc_sw128(v0 + o), id_o1, 1u);
          }
          tc::mma_commit(&vempty[st]);
          tc::mma_commit(&p_free);
          if (j == nblk - 1) tc::mma_commit(&o_done);
        }
        __syncwarp();
        FF_TRACE(lane == 0, pbuf, 3);
        ++pit;
      }
    }
  } else {
    // =========================== softmax (16 warps: TMEM lane quadrant = warp % 4, token quarter = (warp - 2) / 4) ===========================
    // Four warps per lane quadrant, 16 of a block's 64 tokens each: with eight warps (32 tokens per thread) the softmax of a block --
    // TMEM loads, 32 exponentials, the hi/lo split, the swizzled stores -- took longer than the block's MMAs and the tensor pipe idled.
    const int q = warp & 3;
    const int qt = (warp - 2) >> 2;                 // tokens [16 qt, 16 qt + 16) of a block
    const int r = q * 32 + lane;                    // row inside the tile
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float* xch = reinterpret_cast<float*>(p_s);     // [4][128] exchange area between the quarters (P is idle when it is used)
    uint32_t sct = 0, pct = 0;
    for (int pc = 0; pc < npiece; ++pc) {
      const FFPiece p = piece_s[pc];
      // ---- pass 1: row maximum of the approximate scores ----
      float row_max = -INFINITY;
      for (int j = 0; j < p.nblk; j += 2, ++sct) {          // block pairs: columns [0, 64) = block j, [64, 128) = block j + 1
        const uint32_t sb = sct & 1u;
        const bool pair = j + 1 < p.nblk;
        tc::mbar_wait(&s_ready[sb], (sct >> 1) & 1u, err, 8);
        tc::tc_fence_after_sync();
        FF_TRACE(tid == 64, sct, 4);
        const int t0 = (p.blk0 + j) * kTB + qt * 16;
        uint32_t a[16], b2[16];
        const uint32_t base = tmem_s + lane_addr + sb * 128u + (uint32_t)(qt * 16);
        tc::tmem_ld16_issue(base, a);
        if (pair) tc::tmem_ld16_issue(base + (uint32_t)kTB, b2);
        tc::tmem_ld_wait();
        tc::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&s_free[sb]);
        FF_TRACE(tid == 64, sct, 5);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (t0 + i < L) row_max = fmaxf(row_max, __uint_as_float(a[i]));
          if (pair && t0 + kTB + i < L) row_max = fmaxf(row_max, __uint_as_float(b2[i]));
        }
      }
      xch[qt * 128 + r] = row_max;
      ff_bar(1, kSoftWarps * 32);
      row_max = fmaxf(fmaxf(xch[r], xch[128 + r]), fmaxf(xch[256 + r], xch[384 + r]));
      ff_bar(1, kSoftWarps * 32);                   // everybody has read the exchange area before P is written again
      // ---- pass 2: exact scores, P = exp2(S - m), row sums ----
      float row_sum = 0.f;
      for (int j = 0; j < p.nblk; ++j, ++sct, ++pct) {
        const uint32_t sb = sct & 1u;
        tc::mbar_wait(&s_ready[sb], (sct >> 1) & 1u, err, 8);
        tc::tc_fence_after_sync();
        FF_TRACE(tid == 64, sct, 4);
        const int t0 = (p.blk0 + j) * kTB + qt * 16;
        uint32_t a[16], b2[16];
        const uint32_t base = tmem_s + lane_addr + sb * 128u + (uint32_t)(qt * 16);
        tc::tmem_ld16_issue(base, a);
        tc::tmem_ld16_issue(base + (uint32_t)kTB, b2);
        tc::tmem_ld_wait();
        tc::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&s_free[sb]);
        FF_TRACE(tid == 64, sct, 5);
        __half2 hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float s0 = __uint_as_float(a[2 * i]) + __uint_as_float(b2[2 * i]);
          const float s1 = __uint_as_float(a[2 * i + 1]) + __uint_as_float(b2[2 * i + 1]);
          const float p0v = (t0 + 2 * i < L) ? ff_exp2(s0 - row_max) : 0.f;
          const float p1v = (t0 + 2 * i + 1 < L) ? ff_exp2(s1 - row_max) : 0.f;
          row_sum += p0v + p1v;
          hi[i] = __floats2half2_rn(p0v, p1v);
          const float2 hf2 = __half22float2(hi[i]);
          lo[i] = __floats2half2_rn(p0v - hf2.x, p1v - hf2.y);
        }
        tc::mbar_wait(&p_free, (pct & 1u) ^ 1u, err, 9);                      // the previous block's P.V MMAs are done with p_s
        const int ch = qt * 2;                       // 16-byte chunk index of this quarter's first token inside the 64-token (128-byte) row
#pragma unroll
        for (int c4 = 0; c4 < 2; ++c4) {
          tc::st_shared_16(p_s + tc::sw128_offset(r, ch + c4), FHalf8{hi[4 * c4], hi[4 * c4 + 1], hi[4 * c4 + 2], hi[4 * c4 + 3]});
          tc::st_shared_16(p_s + kQImg + tc::sw128_offset(r, ch + c4), FHalf8{lo[4 * c4], lo[4 * c4 + 1], lo[4 * c4 + 2], lo[4 * c4 + 3]});
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&p_ready);
        FF_TRACE(tid == 64, sct, 6);
        FF_TRACE(tid == 64 + 15 * 32, sct, 7);
      }
      // ---- partial result of this piece: O (un-normalised), row maximum and row sum ----
      tc::mbar_wait(&o_done, (uint32_t)pc & 1u, err, 10);
      tc::tc_fence_after_sync();
      xch[qt * 128 + r] = row_sum;                  // (all P.V MMAs are complete: P is idle)
      float* op = Opart + ((size_t)p.slot * kTQ + r) * kD + qt * 32;
#pragma unroll 1
      for (int cb = 0; cb < 32; cb += 16) {
        uint32_t a[16], b2[16];
        tc::tmem_ld16_issue(tmem_o + lane_addr + (uint32_t)(qt * 32 + cb), a);
        tc::tmem_ld16_issue(tmem_o + lane_addr + (uint32_t)(kD + qt * 32 + cb), b2);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i)
          reinterpret_cast<float4*>(op + cb)[i] =
              make_float4(__uint_as_float(a[4 * i]) + __uint_as_float(b2[4 * i]), __uint_as_float(a[4 * i + 1]) + __uint_as_float(b2[4 * i + 1]),
                          __uint_as_float(a[4 * i + 2]) + __uint_as_float(b2[4 * i + 2]), __uint_as_float(a[4 * i + 3]) + __uint_as_float(b2[4 * i + 3]));
      }
      tc::tc_fence_before_sync();
      ff_bar(1, kSoftWarps * 32);
      if (qt == 0) {
        ml[((size_t)p.slot * kTQ + r) * 2] = row_max;
        ml[((size_t)p.slot * kTQ + r) * 2 + 1] = (xch[r] + xch[128 + r]) + (xch[256 + r] + xch[384 + r]);
      }
      ff_bar(1, kSoftWarps * 32);                   // the exchange area is free again; O is drained
      if (lane == 0) tc::mbar_arrive(&o_free);
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base_s, 512);
}

// out[row, :] = sum_p w_p O_p / sum_p w_p l_p,  w_p = exp2(m_p - max_p m_p), over the pieces p of the row's tile.
// grid = (tiles, 16 row groups): one warp per row, eight rows per CTA (a tile of a small problem has ~15 pieces whose loads are
// dependent: one CTA per tile left 9 CTAs crawling through 128 rows each, 137 us for a single C2 fragment); the piece <-> slot mapping
// is recomputed from the same quantities the attention kernel used.
__global__ void __launch_bounds__(256) k_flash_combine(const float* __restrict__ Opart, const float* __restrict__ ml, const int* __restrict__ seg,
                                                       const int* __restrict__ cnt, const int* __restrict__ m_ptr, int B, int M_max, int L,
                                                       int grid_att, float* __restrict__ out, int ldo, __half* __restrict__ out_h) {
  const int nb = (L + kTB - 1) / kTB;
  long long T = 0;
  for (int b = 0; b < B; ++b) T += (ff_item_rows(seg, cnt, m_ptr, M_max, b) + kTQ - 1) / kTQ;
  const long long tile = blockIdx.x;
  if (tile >= T) return;
  const long long total = T * nb;
  const long long per = (total + grid_att - 1) / grid_att;
  int b = 0;
  long long tb = 0;
  for (;;) {
    const int tiles_b = (ff_item_rows(seg, cnt, m_ptr, M_max, b) + kTQ - 1) / kTQ;
    if (tile < tb + tiles_b) break;
    tb += tiles_b;
    ++b;
  }
  const int s0 = seg ? seg[b] : 0;
  const int row0 = s0 + (int)(tile - tb) * kTQ, row_end = s0 + ff_item_rows(seg, cnt, m_ptr, M_max, b);
  const int c_first = (int)((tile * nb) / per), c_last = (int)(((tile + 1) * nb - 1) / per);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  {
    const int r = blockIdx.y * 8 + w;
    const int row = row0 + r;
    if (row >= row_end) return;
    float mx = -INFINITY;
    for (int c = c_first; c <= c_last; ++c) mx = fmaxf(mx, ml[((size_t)((int)tile + c) * kTQ + r) * 2]);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float den = 0.f;
    for (int c = c_first; c <= c_last; ++c) {
      const size_t slot = (size_t)((int)tile + c);
      const float m = ml[(slot * kTQ + r) * 2];
      if (m == -INFINITY) continue;                  // a piece whose tokens were all padding
      const float wgt = ff_exp2(m - mx);
      den = fmaf(wgt, ml[(slot * kTQ + r) * 2 + 1], den);
      const float4 o = *reinterpret_cast<const float4*>(Opart + (slot * kTQ + r) * kD + lane * 4);
      acc.x = fmaf(wgt, o.x, acc.x); acc.y = fmaf(wgt, o.y, acc.y); acc.z = fmaf(wgt, o.z, acc.z); acc.w = fmaf(wgt, o.w, acc.w);
    }
    const float inv = 1.0f / den;
    if (out_h) {          // h2 output (chunk width 64, ldo in halves): the operand of the to_out GEMM
      const float v[4] = {acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv};
      const __half2 h01 = __floats2half2_rn(v[0], v[1]), h23 = __floats2half2_rn(v[2], v[3]);
      const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
      const __half2 l01 = __floats2half2_rn(v[0] - b01.x, v[1] - b01.y), l23 = __floats2half2_rn(v[2] - b23.x, v[3] - b23.y);
      const int c = lane * 4;
      __half* p = out_h + (size_t)row * ldo + (c >> 6) * 128 + (c & 63);
      *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const unsigned*>(&h01), *reinterpret_cast<const unsigned*>(&h23));
      *reinterpret_cast<uint2*>(p + 64) = make_uint2(*reinterpret_cast<const unsigned*>(&l01), *reinterpret_cast<const unsigned*>(&l23));
      return;
    }
    *reinterpret_cast<float4*>(out + (size_t)row * ldo + lane * 4) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  }
}

}  // namespace
#ifdef IMF_FF_TRACE
extern "C" int imf_debug_flash_trace(long long* host_out, int n) {
  return cudaMemcpyFromSymbol(host_out, g_ff_trace, sizeof(long long) * (size_t)(n < 8 * 256 ? n : 8 * 256)) == cudaSuccess ? IMF_OK : IMF_ERR_CUDA;
}
#endif
namespace {
inline int ff_lpad(int L) { return (L + kTB - 1) / kTB * kTB; }
inline int ff_tiles_max(int M_max, int B) { return (M_max + kTQ - 1) / kTQ + B; }
inline int ff_grid(int M_max, int L, int B) {
  const long long nb = (L + kTB - 1) / kTB, tmax = ff_tiles_max(M_max, B);
  long long g = imf_sm_count();
  if (g > tmax * nb) g = tmax * nb;
  // a CTA holds at most kMaxPieces pieces (it spans at most ceil(per / nb) + 1 tiles): keep per <= nb * (kMaxPieces - 2)
  const long long gmin = (tmax + kMaxPieces - 3) / (kMaxPieces - 2);
  if (g < gmin) g = gmin;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace

// sizes of the h2 copies of K and V^T of B images with L tokens each, and of the attention workspace (partials of every piece)
size_t imf_flash_kv_h2_bytes(int L, int B) { return (size_t)B * 2 * ff_lpad(L) * kD * 2 * sizeof(__half); }
size_t imf_flash_workspace_bytes(int M_max, int L, int B) {
  const size_t slots = (size_t)ff_tiles_max(M_max, B) + (size_t)ff_grid(M_max, L, B) + 1;
  return slots * kTQ * (kD + 2) * sizeof(float);
}

// K / V (fp32) -> kvh2 = { Kh2 [B*Lpad, 256 halves], Vth2 [B*128, 2*Lpad halves] }.
// v_transposed == 0: K [B*L, ldk] and V [B*L, ldv] row-major (e.g. columns [0,128) and [128,256) of one [B*L, 256] matrix);
// v_transposed != 0 (B == 1): K [L, ldk] and V^T [128, ldv].
int imf_flash_pack_kv(const float* K, int ldk, const float* V, int ldv, int v_transposed, int L, int B, void* kvh2, cudaStream_t stream) {
  const int Lpad = ff_lpad(L);
  __half* Kh = reinterpret_cast<__half*>(kvh2);
  __half* Vh = Kh + (size_t)B * Lpad * 2 * kD;
  k_flash_pack_kv<<<dim3(Lpad / kTB, B, 2), 256, 0, stream>>>(K, ldk, V, ldv, v_transposed, L, Lpad, B, Kh, Vh);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

// o[row, 128] (fp32, row stride ldo) = softmax(q k_b^T) v_b for the rows [seg[b], seg[b] + cnt[b]) of every item b < B;
// qh2 = h2 matrix of the queries [M_max, 256 halves], already scaled by log2(e) / sqrt(d).  seg / cnt == NULL: one item with
// *m_dev (or M_max) rows starting at row 0.
static int ff_run(const void* qh2, int M_max, const int* seg_dev, const int* cnt_dev, const int* m_dev, int B, const void* kvh2, int L, float* o,
                  __half* o_h, int ldo, void* workspace, size_t workspace_bytes, int* err, cudaStream_t stream);
int imf_flash_attention(const void* qh2, int M_max, const int* seg_dev, const int* cnt_dev, const int* m_dev, int B, const void* kvh2, int L,
                        float* o, int ldo, void* workspace, size_t workspace_bytes, int* err, cudaStream_t stream) {
  return ff_run(qh2, M_max, seg_dev, cnt_dev, m_dev, B, kvh2, L, o, nullptr, ldo, workspace, workspace_bytes, err, stream);
}
// the same with the output written as an h2 matrix (chunk width 64, ldo_h halves)
int imf_flash_attention_h2(const void* qh2, int M_max, const int* seg_dev, const int* cnt_dev, const int* m_dev, int B, const void* kvh2, int L,
                           void* o_h2, int ldo_h, void* workspace, size_t workspace_bytes, int* err, cudaStream_t stream) {
  return ff_run(qh2, M_max, seg_dev, cnt_dev, m_dev, B, kvh2, L, nullptr, reinterpret_cast<__half*>(o_h2), ldo_h, workspace, workspace_bytes, err,
                stream);
}
static int ff_run(const void* qh2, int M_max, const int* seg_dev, const int* cnt_dev, const int* m_dev, int B, const void* kvh2, int L, float* o,
                  __half* o_h, int ldo, void* workspace, size_t workspace_bytes, int* err, cudaStream_t stream) {
  IMF_CHECK_ARG(qh2 && kvh2 && (o || o_h) && workspace && M_max > 0 && L > 0 && B >= 1 && B <= kMaxItems);
  IMF_CHECK_ARG((seg_dev == nullptr) == (cnt_dev == nullptr) && (cnt_dev != nullptr || B == 1));
  IMF_CHECK_ARG(workspace_bytes >= imf_flash_workspace_bytes(M_max, L, B));
  const int Lpad = ff_lpad(L);
  const __half* Kh = reinterpret_cast<const __half*>(kvh2);
  const __half* Vh = Kh + (size_t)B * Lpad * 2 * kD;
  CUtensorMap tmQ, tmK, tmV;
  int rc = tma::encode_2d_u16(&tmQ, qh2, (uint64_t)M_max, 2 * kD, 2 * kD, 64, kTQ);
  if (!rc) rc = tma::encode_2d_u16(&tmK, Kh, (uint64_t)B * Lpad, 2 * kD, 2 * kD, 64, kTB);
  if (!rc) rc = tma::encode_2d_u16(&tmV, Vh, (uint64_t)B * kD, (uint64_t)2 * Lpad, (uint64_t)2 * Lpad, 64, kD);
  if (rc) { imf_set_error("cuTensorMapEncodeTiled (flash attention) failed: %d", rc); return IMF_ERR_CUDA; }
  const int grid = ff_grid(M_max, L, B);
  const size_t slots = (size_t)ff_tiles_max(M_max, B) + (size_t)grid + 1;
  float* Opart = reinterpret_cast<float*>(workspace);
  float* ml = Opart + slots * kTQ * kD;
  IMF_CHECK_CUDA(imf_set_max_smem_once(reinterpret_cast<const void*>(&k_flash_fusion), kSmem + 1024));
  k_flash_fusion<<<grid, kThreads, kSmem + 1024, stream>>>(tmQ, tmK, tmV, seg_dev, cnt_dev, m_dev, B, M_max, L, Lpad, Opart, ml, err);
  IMF_CHECK_LAUNCH();
  k_flash_combine<<<dim3(ff_tiles_max(M_max, B), kTQ / 8), 256, 0, stream>>>(Opart, ml, seg_dev, cnt_dev, m_dev, B, M_max, L, grid, o, ldo, o_h);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}
