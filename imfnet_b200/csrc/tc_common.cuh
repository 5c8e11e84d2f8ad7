// imfnet_b200 -- sm_100a tensor-core plumbing: mbarrier, TMEM allocation, tcgen05.mma / commit / ld, UMMA descriptors,
// bulk (TMA) copies.  Inline PTX only; formats follow the PTX ISA "tcgen05" chapter (descriptor bit layouts
// cross-checked against CUTLASS cute/arch/mma_sm100_desc.hpp).
//
// Precision scheme used by every kernel built on this file ("3xTF32"): an fp32 value v is split into
//   hi = v with the 13 low mantissa bits cleared (exactly a TF32 number), lo = v - hi (exact in fp32),
// and a product a*b is accumulated as lo_a*hi_b + hi_a*lo_b + hi_a*hi_b in the fp32 TMEM accumulator.
// The dropped lo*lo term and the TF32 rounding of lo are ~2^-22 relative: fp32-class accuracy (SURVEY.md 7.3-1
// measured 5.1e-7 end to end for this scheme against 9.1e-4 for single-pass TF32).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- error reporting / watchdog ---------------------------------------------------------------------
// A wait that does not complete within ~2 s of SM clock writes a code to *err (if non-null) and traps, so a
// protocol bug surfaces as a failed launch instead of a hung GPU.
#define TC_WATCHDOG_CYCLES 4000000000ll

// ---- mbarrier ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > TC_WATCHDOG_CYCLES) {
      if (err) atomicExch(err, code);
      __threadfence_system();
      __trap();
    }
  }
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma / bulk copies read smem through it)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM -------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {   // one full warp; ncols power of 2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- descriptors ------------------------------------------------------------------------------------
// Shared-memory operand, K-major, 128-byte swizzle: rows of 128 B (32 fp32), 8-row groups 1024 B apart.
// Element (row r, 16-byte chunk c) of a tile based at a 1024-aligned address lives at
//   (r / 8) * 1024 + (r % 8) * 128 + ((c ^ (r % 8)) * 16).
// Successive K slices of one MMA (32 B each) are addressed by adding 32 B to the start address.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address        bits [0,14)
  d |= (uint64_t)1 << 16;                          // leading byte offset  bits [16,30)  (unused for swizzled K-major)
  d |= (uint64_t)(1024u >> 4) << 32;               // stride byte offset   bits [32,46)  = 1024 B per 8-row group
  d |= (uint64_t)1 << 46;                          // descriptor version   bits [46,48)  = 1 on sm_100
  d |= (uint64_t)2 << 61;                          // layout type          bits [61,64)  = SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

// Instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major, M x N tile.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) /*D = f32*/ | (2u << 7) /*A = tf32*/ | (2u << 10) /*B = tf32*/ | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every MMA issued so far by this thread has completed (implies fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: 32 lanes x 32 bit, 16 consecutive columns; thread t of warp w reads lane 32*(w%4)+t ----
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// the same load without the wait: issue several, then tmem_ld_wait() once, then read the registers
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- bulk copy global -> shared (TMA, 1-D), completion counted in bytes on an mbarrier -----------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- 3xTF32 split -------------------------------------------------------------------------------------
__device__ __forceinline__ void split_tf32(const float4& v, float4& hi, float4& lo) {
  hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  lo.x = v.x - hi.x;
  lo.y = v.y - hi.y;
  lo.z = v.z - hi.z;
  lo.w = v.w - hi.w;
}

// One 16-byte global store of a struct of four 32-bit members (e.g. four __half2).  Written as PTX because a struct assignment through a
// reinterpret_cast pointer compiles to FOUR 4-byte STG when the compiler cannot prove the alignment (seen in the scattered-row epilogue of
// the transposed convolutions: 32 M sector writes instead of 8 M, the launch was bound by them).
template <class V16>
__device__ __forceinline__ void st_global_16(void* p, const V16& v) {
  static_assert(sizeof(V16) == 16, "16-byte value expected");
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
}

// One 16-byte SHARED-memory store (STS.128) of a 16-byte value at a generic pointer into shared memory.  The epilogues' staging pointers
// are derived from the 1024-aligned dynamic shared base through integer arithmetic, so the compiler only knows them as generic pointers
// and emits ST.E.128 (generic store: tracked on the long scoreboard, the source registers are released late -- ncu showed the hi/lo
// conversions of the next 16 channels waiting for them); st.shared is explicit here.
template <class V16>
__device__ __forceinline__ void st_shared_16(void* p, const V16& v) {
  static_assert(sizeof(V16) == 16, "16-byte value expected");
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

}  // namespace tc
