// imfnet_b200 -- shared device/host helpers for the sm_100a kernels behind include/imfnet_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define IMF_OK 0
#define IMF_ERR_BAD_ARG (-1)
#define IMF_ERR_CUDA (-2)
#define IMF_ERR_UNSUPPORTED (-3)

// status word bits written by the coordinate kernels (device int32, read back by the host mirror)
#define IMF_STATUS_COORD_RANGE 1      // |x|,|y|,|z| >= 2^15 or batch index outside [0, 65534]
#define IMF_STATUS_DUPLICATE 2        // duplicate coordinate handed to imf_hash_build (rows must be unique)
#define IMF_STATUS_TABLE_FULL 4       // probe sequence exhausted (capacity too small)

void imf_set_error(const char* fmt, ...);
void imf_note_launch();   // bumps the process-wide kernel-launch counter reported by imf_launch_count()

#define IMF_CHECK_ARG(cond)                                                              \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      imf_set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, #cond);               \
      return IMF_ERR_BAD_ARG;                                                            \
    }                                                                                    \
  } while (0)

#define IMF_CHECK_LAUNCH()                                                               \
  do {                                                                                   \
    imf_note_launch();                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      imf_set_error("%s:%d: CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return IMF_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define IMF_CHECK_CUDA(expr)                                                             \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      imf_set_error("%s:%d: CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return IMF_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

// Opt a kernel into its dynamic shared-memory size once per (kernel, device): the attribute is per device, and the call costs a
// few microseconds, too much for every launch of a 150-launch forward.  Defined in coords.cu.
cudaError_t imf_set_max_smem_once(const void* kernel, int bytes);

// SM count of the current device, cached (coords.cu); 148 when no device is present
int imf_sm_count();

// tensor-core GEMM (tc_gemm.cu), used by the attention-fusion orchestration in dense.cu
extern "C" size_t imf_tc_gemm_workspace_bytes(int32_t M, int32_t N, int32_t K);
extern "C" int imf_tc_gemm(const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc, int32_t M, int32_t N,
                           int32_t K, float alpha, const float* bias, const float* R, int32_t ldr, int32_t geglu, void* workspace,
                           size_t workspace_bytes, int32_t* err, cudaStream_t stream);

extern "C" int imf_tc_gemm_m(const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc, int32_t M,
                             const int32_t* m_dev, int32_t N, int32_t K, float alpha, const float* bias, const float* R, int32_t ldr,
                             int32_t geglu, void* workspace, size_t workspace_bytes, int32_t* err, cudaStream_t stream);

// ---- coordinate keys + open-addressing hash table ------------------------------------------------
// One slot = 16 bytes so that a probe is a single 128-bit load.
struct __align__(16) ImfSlot {
  unsigned long long key;
  int val;
  int pad;
};
static_assert(sizeof(ImfSlot) == 16, "slot must be 16 bytes");

#define IMF_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull

__host__ __device__ __forceinline__ bool imf_coord_in_range(int b, int x, int y, int z) {
  return (unsigned)b < 65535u && (unsigned)(x + 32768) < 65536u && (unsigned)(y + 32768) < 65536u &&
         (unsigned)(z + 32768) < 65536u;
}

__host__ __device__ __forceinline__ unsigned long long imf_pack_key(int b, int x, int y, int z) {
  return ((unsigned long long)(unsigned)b << 48) | ((unsigned long long)(unsigned)(x + 32768) << 32) |
         ((unsigned long long)(unsigned)(y + 32768) << 16) | (unsigned long long)(unsigned)(z + 32768);
}

__host__ __device__ __forceinline__ unsigned long long imf_hash64(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

#ifdef __CUDACC__
// Look a key up; returns the stored value or -1.  `mask` = capacity-1 (capacity is a power of two).
__device__ __forceinline__ int imf_table_lookup(const ImfSlot* __restrict__ table, unsigned long long mask,
                                                unsigned long long key) {
  unsigned long long slot = imf_hash64(key) & mask;
  for (unsigned long long probes = 0; probes <= mask; ++probes) {
    const int4 raw = __ldg(reinterpret_cast<const int4*>(table + slot));
    const unsigned long long k = ((unsigned long long)(unsigned)raw.y << 32) | (unsigned)raw.x;
    if (k == key) return raw.z;
    if (k == IMF_EMPTY_KEY) return -1;
    slot = (slot + 1) & mask;
  }
  return -1;
}

__device__ __forceinline__ float imf_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float imf_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif
