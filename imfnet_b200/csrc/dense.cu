// imfnet_b200 -- attention-fusion dense path, fp32 SIMT tier (LayerNorm, GEMM with fused epilogues, softmax).
//
// Replaces, for the IMFNet descriptor path, AttentionFusion.forward with depth=0, one cross head, mask=None:
//   /root/reference/model/attention_fusion.py:132-154 (PreNorm :32-46, Attention :65-95, FeedForward :53-63, GEGLU :48-51)
// as called per batch item from ResUNet2.transformer (/root/reference/model/resunet.py:237-273):
//   x = P;  q = LN(x) Wq^T;  [k|v] = LN_c(I) Wkv^T;  A = softmax(q k^T / sqrt(d));  x = (A v) Wo^T + bo + x;
//   x = W2 . geglu(W1 . LN(x) + b1) + b2 + x
// Image tokens I arrive channel-major ([C, H*W], the NCHW feature map of the image encoder), so the context
// LayerNorm also performs the [C,L] -> [L,C] transposition of resunet.py:259-261.
#include <cuda_fp16.h>

#include "common.cuh"

// fused attention kernel (flash_fusion.cu), used when the head has 128 channels (the IMFNet configuration)
size_t imf_flash_kv_h2_bytes(int L, int B);
size_t imf_flash_workspace_bytes(int M_max, int L, int B);
int imf_flash_pack_kv(const float* K, int ldk, const float* V, int ldv, int v_transposed, int L, int B, void* kvh2, cudaStream_t stream);
int imf_flash_attention(const void* qh2, int M_max, const int* seg_dev, const int* cnt_dev, const int* m_dev, int B, const void* kvh2, int L,
                        float* o, int ldo, void* workspace, size_t workspace_bytes, int* err, cudaStream_t stream);
int imf_flash_attention_h2(const void* qh2, int M_max, const int* seg_dev, const int* cnt_dev, const int* m_dev, int B, const void* kvh2, int L,
                           void* o_h2, int ldo_h, void* workspace, size_t workspace_bytes, int* err, cudaStream_t stream);
extern "C" int imf_h2_pack_n(const float* X, int32_t ldx, int32_t n, const int32_t* n_dev, int32_t C, int32_t KC, void* H, int32_t ldh,
                             int32_t* err, cudaStream_t stream);
extern "C" int imf_h2_gemm(const void* A, int32_t lda, int32_t M_max, const int32_t* m_dev, const void* Wpacked, int32_t N, int32_t K,
                           float alpha, const float* bias, const float* R, int32_t ldr, int32_t mode, void* C, int32_t ldc, int32_t* err,
                           cudaStream_t stream);

namespace {

// ---- LayerNorm over the last dim of row-major rows; one warp per row, C <= 1024, C % 32 == 0 ----
// Yh (optional): the result is written as an h2 matrix (fp16 hi/lo, chunk width 64, row stride ldy halves) instead of fp32 rows.
__global__ void __launch_bounds__(256) k_layernorm_rows(const float* __restrict__ X, int ldx, int M, int C,
                                                        const float* __restrict__ g, const float* __restrict__ b, float eps,
                                                        float* __restrict__ Y, int ldy, const int* __restrict__ m_ptr,
                                                        __half* __restrict__ Yh = nullptr) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m_ptr) { const int v = *m_ptr; M = v < M ? v : M; }
  if (row >= M) return;
  const float* x = X + (size_t)row * ldx;
  const bool aligned8 = ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(b) |
                          reinterpret_cast<uintptr_t>(Y)) & 7) == 0 && (reinterpret_cast<uintptr_t>(Yh) & 3) == 0;
  if ((C & 63) == 0 && C <= 512 && (ldx & 1) == 0 && (ldy & 1) == 0 && aligned8) {
    // two adjacent channels per lane: 8-byte loads, 4-byte (half2) / 8-byte stores -- the 2-byte stores of the one-channel form made
    // the h2 output LSU-bound (77 % LSU, 22 us for 48 k x 128 tokens: profiles/r02/call43_ncu_other_kernels.txt)
    float2 w2[8];
    const int per2 = C >> 6;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < per2) { w2[i] = *reinterpret_cast<const float2*>(x + i * 64 + 2 * lane); s += w2[i].x + w2[i].y; }
    const float mean = imf_warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < per2) { const float d0 = w2[i].x - mean, d1 = w2[i].y - mean; q += d0 * d0 + d1 * d1; }
    const float rstd = 1.0f / sqrtf(imf_warp_sum(q) / (float)C + eps);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < per2) {
        const int c = i * 64 + 2 * lane;
        const float2 g2 = __ldg(reinterpret_cast<const float2*>(g + c)), b2 = __ldg(reinterpret_cast<const float2*>(b + c));
        const float o0 = (w2[i].x - mean) * rstd * g2.x + b2.x, o1 = (w2[i].y - mean) * rstd * g2.y + b2.y;
        if (Yh) {
          const __half2 h = __floats2half2_rn(o0, o1);
          const float2 hf = __half22float2(h);
          __half* p = Yh + (size_t)row * ldy + i * 128 + 2 * lane;          // chunk i = channels [64 i, 64 i + 64): [hi 64 | lo 64]
          *reinterpret_cast<__half2*>(p) = h;
          *reinterpret_cast<__half2*>(p + 64) = __floats2half2_rn(o0 - hf.x, o1 - hf.y);
        } else {
          *reinterpret_cast<float2*>(Y + (size_t)row * ldy + c) = make_float2(o0, o1);
        }
      }
    return;
  }
  float v[32];
  const int per = C / 32;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < per) { v[i] = x[i * 32 + lane]; s += v[i]; }
  const float mean = imf_warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < per) { const float d = v[i] - mean; q += d * d; }
  const float rstd = 1.0f / sqrtf(imf_warp_sum(q) / (float)C + eps);
  if (Yh) {
    __half* yh = Yh + (size_t)row * ldy;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < per) {
        const int c = i * 32 + lane;
        const float o = (v[i] - mean) * rstd * __ldg(g + c) + __ldg(b + c);
        const __half h = __float2half_rn(o);
        __half* p = yh + (c >> 6) * 128 + (c & 63);
        p[0] = h;
        p[64] = __float2half_rn(o - __half2float(h));
      }
    return;
  }
  float* y = Y + (size_t)row * ldy;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < per) { const int c = i * 32 + lane; y[c] = (v[i] - mean) * rstd * __ldg(g + c) + __ldg(b + c); }
}

// ---- LayerNorm of image tokens stored channel-major: X[C][L] -> Y[L][C] (C == 128) ----------------
__global__ void __launch_bounds__(256) k_layernorm_tokens_chw(const float* __restrict__ X, int L, const float* __restrict__ g,
                                                              const float* __restrict__ b, float eps, float* __restrict__ Y) {
  constexpr int C = 128;
  __shared__ float t[C][33];
  const int l0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < C; c += 8) {
    const int l = l0 + lane;
    t[c][lane] = (l < L) ? X[(size_t)c * L + l] : 0.f;
  }
  __syncthreads();
  for (int tok = warp; tok < 32; tok += 8) {
    const int l = l0 + tok;
    if (l >= L) break;
    float v[4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = t[i * 32 + lane][tok]; s += v[i]; }
    const float mean = imf_warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float d = v[i] - mean; q += d * d; }
    const float rstd = 1.0f / sqrtf(imf_warp_sum(q) / (float)C + eps);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = i * 32 + lane;
      Y[(size_t)l * C + c] = (v[i] - mean) * rstd * __ldg(g + c) + __ldg(b + c);
    }
  }
}

// ---- row softmax in place, one CTA per row ---------------------------------------------------------
__global__ void __launch_bounds__(256) k_softmax_rows(float* __restrict__ S, int lds, int M, int L, const int* __restrict__ m_ptr) {
  __shared__ float red[8];
  __shared__ float bcast;
  const int row = blockIdx.x;
  if (m_ptr) { const int v = *m_ptr; M = v < M ? v : M; }
  if (row >= M) return;
  float* s = S + (size_t)row * lds;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float mx = -INFINITY;
  for (int i = tid; i < L; i += 256) mx = fmaxf(mx, s[i]);
  mx = imf_warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (tid == 0) { float m = red[0]; for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]); bcast = m; }
  __syncthreads();
  mx = bcast;
  float sum = 0.f;
  for (int i = tid; i < L; i += 256) { const float e = expf(s[i] - mx); s[i] = e; sum += e; }
  sum = imf_warp_sum(sum);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (tid == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; bcast = t; }
  __syncthreads();
  const float inv = 1.0f / bcast;
  for (int i = tid; i < L; i += 256) s[i] *= inv;
  for (int i = L + tid; i < lds; i += 256) s[i] = 0.f;   // padding columns feed the PV GEMM as zeros
}

// ---- SGEMM: C[M,N] = alpha * A[M,K] . op(B) (+ bias[n]) (+ R[m,n]);  op(B) = B^T for B [N,K] (weights) or B for [K,N] ----
// GEGLU mode: B has 2N rows ([N,K] layout); C[m,n] = (acc_n + bias[n]) * gelu(acc_{n+N} + bias[n+N])  (exact erf GELU).
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <bool B_KN, bool GEGLU>
__global__ void __launch_bounds__(256) k_sgemm(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                               float* __restrict__ C, int ldc, int M, int N, int K, float alpha,
                                               const float* __restrict__ bias, const float* __restrict__ R, int ldr) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[GEGLU ? 2 : 1][BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4], acc2[GEGLU ? 4 : 1][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; if (GEGLU) acc2[i][j] = 0.f; }

  for (int k0 = 0; k0 < K; k0 += BK) {
    {   // A tile: 64 rows x 16 k; thread -> (row = tid/4, k4 = (tid%4)*4)
      const int r = tid >> 2, kk = (tid & 3) * 4;
      const int m = m0 + r;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + kk + i;
        As[kk + i][r] = (m < M && k < K) ? A[(size_t)m * lda + k] : 0.f;
      }
    }
    if (!B_KN) {
      const int r = tid >> 2, kk = (tid & 3) * 4;
      const int n = n0 + r;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + kk + i;
        Bs[0][kk + i][r] = (n < N && k < K) ? B[(size_t)n * ldb + k] : 0.f;
        if (GEGLU) Bs[GEGLU ? 1 : 0][kk + i][r] = (n < N && k < K) ? B[(size_t)(n + N) * ldb + k] : 0.f;
      }
    } else {   // B [K,N]: thread -> (k = tid/16, n4 = (tid%16)*4)
      const int kk = tid >> 4, nn = (tid & 15) * 4;
      const int k = k0 + kk;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = n0 + nn + i;
        Bs[0][kk][nn + i] = (n < N && k < K) ? B[(size_t)k * ldb + n] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[0][kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      if (GEGLU) {
        const float4 b2 = *reinterpret_cast<const float4*>(&Bs[GEGLU ? 1 : 0][kk][tx * 4]);
        const float b2v[4] = {b2.x, b2.y, b2.z, b2.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc2[GEGLU ? i : 0][j] = fmaf(av[i], b2v[j], acc2[GEGLU ? i : 0][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = alpha * acc[i][j];
      if (bias) v += __ldg(bias + n);
      if (GEGLU) {
        float gt = alpha * acc2[GEGLU ? i : 0][j];
        if (bias) gt += __ldg(bias + n + N);
        v = v * gelu_erf(gt);
      }
      if (R) v += R[(size_t)m * ldr + n];
      C[(size_t)m * ldc + n] = v;
    }
  }
}

template <bool B_KN, bool GEGLU>
int sgemm(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K, float alpha,
          const float* bias, const float* R, int ldr, cudaStream_t stream) {
  if (M == 0 || N == 0) return IMF_OK;
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  k_sgemm<B_KN, GEGLU><<<grid, 256, 0, stream>>>(A, lda, B, ldb, C, ldc, M, N, K, alpha, bias, R, ldr);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

inline size_t r256(size_t b) { return (b + 255) / 256 * 256; }

}  // namespace

struct imf_attn_weights_t {
  const float *ln_q_w, *ln_q_b;   // cross_attend_blocks.0.norm            [latent]
  const float *ln_c_w, *ln_c_b;   // cross_attend_blocks.0.norm_context    [dim]
  const float* wq;                // cross_attend_blocks.0.fn.to_q.weight  [inner, latent]
  const float* wkv;               // cross_attend_blocks.0.fn.to_kv.weight [2*inner, dim]
  const float *wo, *bo;           // cross_attend_blocks.0.fn.to_out       [latent, inner], [latent]
  const float *ln_f_w, *ln_f_b;   // cross_attend_blocks.1.norm            [latent]
  const float *w1, *b1;           // cross_attend_blocks.1.fn.net.0        [8*latent, latent], [8*latent]
  const float *w2, *b2;           // cross_attend_blocks.1.fn.net.2        [latent, 4*latent], [latent]
  int32_t latent, dim, inner;
};

static inline int round4(int v) { return (v + 3) / 4 * 4; }

// Layout of the projected context of one image ("kv" buffers): K [L, inner] row-major, then V^T [inner, Lp] (Lp = L rounded
// up to 4) so that both attention GEMMs read K-major operands.
extern "C" size_t imf_attention_kv_bytes(int32_t L, int32_t inner) {
  // K [L, inner] fp32, V^T [inner, Lp] fp32, then (128-channel head) their fp16 hi/lo copies for the fused attention kernel
  return r256((size_t)L * inner * 4) + r256((size_t)inner * round4(L) * 4) + (inner == 128 ? r256(imf_flash_kv_h2_bytes(L, 1)) : 0);
}
extern "C" size_t imf_attention_kv_workspace_bytes(int32_t L, int32_t dim) { return r256((size_t)L * dim * 4); }

// kv = { K = LN_c(tokens) . Wk^T,  V^T = Wv . LN_c(tokens)^T }; tokens are channel-major [dim][L] (channel_major != 0)
// or row-major [L][dim].  (to_kv.weight rows [0,inner) are Wk, rows [inner,2*inner) are Wv: attention_fusion.py:81-82.)
extern "C" int imf_attention_kv(const imf_attn_weights_t* w, const float* tokens, int32_t L, int32_t channel_major, float* kv,
                                void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  IMF_CHECK_ARG(w != nullptr && L >= 0 && w->dim % 32 == 0 && w->dim <= 1024 && (!channel_major || w->dim == 128));
  if (L == 0) return IMF_OK;
  IMF_CHECK_ARG(tokens != nullptr && kv != nullptr && workspace != nullptr);
  IMF_CHECK_ARG(workspace_bytes >= imf_attention_kv_workspace_bytes(L, w->dim));
  float* cn = reinterpret_cast<float*>(workspace);
  if (channel_major)
    k_layernorm_tokens_chw<<<(L + 31) / 32, 256, 0, stream>>>(tokens, L, w->ln_c_w, w->ln_c_b, 1e-5f, cn);
  else
    k_layernorm_rows<<<(L + 7) / 8, 256, 0, stream>>>(tokens, w->dim, L, w->dim, w->ln_c_w, w->ln_c_b, 1e-5f, cn, w->dim, nullptr);
  IMF_CHECK_LAUNCH();
  const int inner = w->inner, dim = w->dim, Lp = round4(L);
  float* Kmat = kv;
  float* Vt = reinterpret_cast<float*>(reinterpret_cast<char*>(kv) + r256((size_t)L * inner * 4));
  int rc;
  if ((rc = imf_tc_gemm(cn, dim, w->wkv, dim, Kmat, inner, L, inner, dim, 1.f, nullptr, nullptr, 0, 0, nullptr, 0, nullptr, stream))) return rc;
  if ((rc = imf_tc_gemm(w->wkv + (size_t)inner * dim, dim, cn, dim, Vt, Lp, inner, L, dim, 1.f, nullptr, nullptr, 0, 0, nullptr, 0, nullptr, stream))) return rc;
  if (inner == 128) {
    void* kvh2 = reinterpret_cast<char*>(kv) + r256((size_t)L * inner * 4) + r256((size_t)inner * Lp * 4);
    return imf_flash_pack_kv(Kmat, inner, Vt, Lp, 1, L, 1, kvh2, stream);
  }
  return IMF_OK;
}

extern "C" size_t imf_attention_workspace_bytes(int32_t M, int32_t L, int32_t latent, int32_t inner) {
  const size_t Lp = (size_t)round4(L);
  return r256((size_t)M * latent * 4)      // xn / reused for LN before FFN
         + r256((size_t)M * inner * 4)     // q
         + r256((size_t)M * Lp * 4)        // scores
         + r256((size_t)M * inner * 4)     // attention output
         + r256((size_t)M * latent * 4)    // x after attention residual
         + r256((size_t)M * latent * 4 * 4)   // GEGLU hidden [M, 4*latent]
         + r256(imf_tc_gemm_workspace_bytes(M, latent, 0))    // split-K partials (largest N used with split-K = latent)
         + (inner == 128 ? r256((size_t)M * inner * 4) + r256(imf_flash_workspace_bytes(M, L, 1)) : 0);   // q as h2 + flash partials
}

// out[M, latent] = cross-attention + GEGLU feed-forward of M point tokens P[M, latent] against kv (imf_attention_kv).
extern "C" int imf_attention_fusion_fwd_m(const imf_attn_weights_t* w, const float* P, int32_t ldp, int32_t M, const int32_t* m_dev,
                                          const float* kv, int32_t L, float* out, int32_t ldo, void* workspace, size_t workspace_bytes,
                                          cudaStream_t stream);
extern "C" int imf_attention_fusion_fwd(const imf_attn_weights_t* w, const float* P, int32_t ldp, int32_t M, const float* kv,
                                        int32_t L, float* out, int32_t ldo, void* workspace, size_t workspace_bytes,
                                        cudaStream_t stream) {
  return imf_attention_fusion_fwd_m(w, P, ldp, M, nullptr, kv, L, out, ldo, workspace, workspace_bytes, stream);
}

// Weights of the fusion module pre-packed for imf_h2_gemm (imf_sparse_conv_h2_pack of W^T as [1, K, N], chunk width 64), with the power-of-two
// scales they were packed with; w1 with its value / gate rows interleaved per 128-column tile (tile t = [value 64t.. | gate 64t..]).
struct imf_attn_packed_t {
  const void *wq, *wkv, *wo, *w1, *w2;
  float mq, mkv, mo, m1, m2;
};

namespace {

constexpr float kLog2e = 1.4426950408889634f;

// the h2 tier of the module body: LayerNorm writes fp16 hi/lo, every projection is a TMA-fed tcgen05 GEMM on pre-split operands
// (h2_gemm.cu), q goes straight into the attention kernel and its output comes back as h2 for to_out
int attention_body_h2(const imf_attn_weights_t* w, const imf_attn_packed_t* wp, const float* P, int ldp, int M, const int* m_dev,
                      const int* seg_dev, const int* cnt_dev, int B, const void* kvh2, int L, float* out, int ldo, void* workspace, int* err,
                      cudaStream_t stream) {
  const int latent = w->latent, inner = w->inner;
  char* ws = reinterpret_cast<char*>(workspace);
  __half* xn = reinterpret_cast<__half*>(ws);  ws += r256((size_t)M * latent * 4);
  __half* q = reinterpret_cast<__half*>(ws);   ws += r256((size_t)M * inner * 4);
  __half* o = reinterpret_cast<__half*>(ws);   ws += r256((size_t)M * inner * 4);
  float* x1 = reinterpret_cast<float*>(ws);    ws += r256((size_t)M * latent * 4);
  __half* hid = reinterpret_cast<__half*>(ws); ws += r256((size_t)M * latent * 4 * 4);
  ws += r256(imf_tc_gemm_workspace_bytes(M, latent, 0)) + r256((size_t)M * inner * 4);          // (regions of the 3xTF32 tier, unused here)
  void* fws = ws;
  const float sm_scale = 1.0f / sqrtf((float)inner);
  int rc;
  k_layernorm_rows<<<(M + 7) / 8, 256, 0, stream>>>(P, ldp, M, latent, w->ln_q_w, w->ln_q_b, 1e-5f, nullptr, 2 * latent, m_dev, xn);
  IMF_CHECK_LAUNCH();
  if ((rc = imf_h2_gemm(xn, 2 * latent, M, m_dev, wp->wq, inner, latent, sm_scale * kLog2e / wp->mq, nullptr, nullptr, 0, 1, q, 2 * inner, err, stream))) return rc;
  if ((rc = imf_flash_attention_h2(q, M, seg_dev, cnt_dev, m_dev, B, kvh2, L, o, 2 * inner, fws, imf_flash_workspace_bytes(M, L, B), err, stream))) return rc;
  if ((rc = imf_h2_gemm(o, 2 * inner, M, m_dev, wp->wo, latent, inner, 1.f / wp->mo, w->bo, P, ldp, 0, x1, latent, err, stream))) return rc;
  k_layernorm_rows<<<(M + 7) / 8, 256, 0, stream>>>(x1, latent, M, latent, w->ln_f_w, w->ln_f_b, 1e-5f, nullptr, 2 * latent, m_dev, xn);
  IMF_CHECK_LAUNCH();
  if ((rc = imf_h2_gemm(xn, 2 * latent, M, m_dev, wp->w1, 8 * latent, latent, 1.f / wp->m1, w->b1, nullptr, 0, 2, hid, 2 * 4 * latent, err, stream))) return rc;
  return imf_h2_gemm(hid, 2 * 4 * latent, M, m_dev, wp->w2, latent, 4 * latent, 1.f / wp->m2, w->b2, x1, latent, 0, out, ldo, err, stream);
}

// The module body for M_max rows (min(*m_dev, M_max) of them valid) that belong to B batch items: everything except the attention
// core is row-wise (LayerNorm, projections, GEGLU feed-forward), so ALL items go through ONE chain of launches; only the attention
// core needs the item structure (rows [seg[b], seg[b] + cnt[b]) attend to image b's tokens).  kvh2: fp16 hi/lo K / V^T of the B images
// (imf_flash_pack_kv); kv32 (B == 1, any head width): the fp32 K / V^T of imf_attention_kv for the unfused path.
int attention_body(const imf_attn_weights_t* w, const float* P, int ldp, int M, const int* m_dev, const int* seg_dev, const int* cnt_dev, int B,
                   const float* kv32, const void* kvh2, int L, float* out, int ldo, void* workspace, int* err, cudaStream_t stream) {
  const int latent = w->latent, inner = w->inner;
  const int Lp = round4(L);
  char* ws = reinterpret_cast<char*>(workspace);
  float* xn = reinterpret_cast<float*>(ws);  ws += r256((size_t)M * latent * 4);
  float* q = reinterpret_cast<float*>(ws);   ws += r256((size_t)M * inner * 4);
  float* S = reinterpret_cast<float*>(ws);   ws += (kvh2 ? 0 : r256((size_t)M * Lp * 4));
  float* o = reinterpret_cast<float*>(ws);   ws += r256((size_t)M * inner * 4);
  float* x1 = reinterpret_cast<float*>(ws);  ws += r256((size_t)M * latent * 4);
  float* hid = reinterpret_cast<float*>(ws); ws += r256((size_t)M * latent * 4 * 4);
  void* gws = ws;
  const size_t gws_bytes = imf_tc_gemm_workspace_bytes(M, latent, 0);
  const float sm_scale = 1.0f / sqrtf((float)inner);
  int rc;
  k_layernorm_rows<<<(M + 7) / 8, 256, 0, stream>>>(P, ldp, M, latent, w->ln_q_w, w->ln_q_b, 1e-5f, xn, latent, m_dev);
  IMF_CHECK_LAUNCH();
  // q = LN(P) . Wq^T   (fused attention: times log2(e) / sqrt(d), the kernel exponentiates with ex2)
  if ((rc = imf_tc_gemm_m(xn, latent, w->wq, latent, q, inner, M, m_dev, inner, latent, kvh2 ? sm_scale * kLog2e : 1.f, nullptr, nullptr, 0, 0,
                          gws, gws_bytes, err, stream))) return rc;
  if (kvh2) {
    // fused attention: q -> fp16 hi/lo -> one persistent tcgen05 kernel for q k^T, softmax and .v of every item (flash_fusion.cu)
    char* ws2 = reinterpret_cast<char*>(gws) + r256(gws_bytes);
    void* qh2 = ws2;
    void* fws = ws2 + r256((size_t)M * inner * 4);
    if ((rc = imf_h2_pack_n(q, inner, M, m_dev, inner, 64, qh2, 2 * inner, err, stream))) return rc;
    if ((rc = imf_flash_attention(qh2, M, seg_dev, cnt_dev, m_dev, B, kvh2, L, o, inner, fws, imf_flash_workspace_bytes(M, L, B), err, stream)))
      return rc;
  } else {
    const float* Kmat = kv32;
    const float* Vt = reinterpret_cast<const float*>(reinterpret_cast<const char*>(kv32) + r256((size_t)L * inner * 4));
    // S = (q . K^T) * scale
    if ((rc = imf_tc_gemm_m(q, inner, Kmat, inner, S, Lp, M, m_dev, L, inner, sm_scale, nullptr, nullptr, 0, 0, nullptr, 0, err, stream))) return rc;
    k_softmax_rows<<<M, 256, 0, stream>>>(S, Lp, M, L, m_dev);
    IMF_CHECK_LAUNCH();
    // o = A . V   (as A . (V^T)^T, split over the L tokens)
    if ((rc = imf_tc_gemm_m(S, Lp, Vt, Lp, o, inner, M, m_dev, inner, L, 1.f, nullptr, nullptr, 0, 0, gws, gws_bytes, err, stream))) return rc;
  }
  // x1 = o . Wo^T + bo + P
  if ((rc = imf_tc_gemm_m(o, inner, w->wo, inner, x1, latent, M, m_dev, latent, inner, 1.f, w->bo, P, ldp, 0, gws, gws_bytes, err, stream))) return rc;
  k_layernorm_rows<<<(M + 7) / 8, 256, 0, stream>>>(x1, latent, M, latent, w->ln_f_w, w->ln_f_b, 1e-5f, xn, latent, m_dev);
  IMF_CHECK_LAUNCH();
  // hid = geglu(LN(x1) . W1^T + b1)   [M, 4*latent]
  if ((rc = imf_tc_gemm_m(xn, latent, w->w1, latent, hid, 4 * latent, M, m_dev, 4 * latent, latent, 1.f, w->b1, nullptr, 0, 1, nullptr, 0, err, stream))) return rc;
  // out = hid . W2^T + b2 + x1
  if ((rc = imf_tc_gemm_m(hid, 4 * latent, w->w2, 4 * latent, out, ldo, M, m_dev, latent, 4 * latent, 1.f, w->b2, x1, latent, 0, gws, gws_bytes, err, stream))) return rc;
  return IMF_OK;
}

}  // namespace

// Same with an optional device-side token count: only min(*m_dev, M) rows are computed (M sizes launches and workspace).
extern "C" int imf_attention_fusion_fwd_m(const imf_attn_weights_t* w, const float* P, int32_t ldp, int32_t M, const int32_t* m_dev,
                                          const float* kv, int32_t L, float* out, int32_t ldo, void* workspace, size_t workspace_bytes,
                                          cudaStream_t stream) {
  IMF_CHECK_ARG(w != nullptr && M >= 0 && L >= 1);
  IMF_CHECK_ARG(w->latent % 32 == 0 && w->latent <= 1024 && w->inner % 4 == 0);
  if (M == 0) return IMF_OK;
  IMF_CHECK_ARG(P != nullptr && kv != nullptr && out != nullptr && workspace != nullptr && ldp >= w->latent && ldo >= w->latent);
  IMF_CHECK_ARG(workspace_bytes >= imf_attention_workspace_bytes(M, L, w->latent, w->inner));
  const void* kvh2 = nullptr;
  if (w->inner == 128)
    kvh2 = reinterpret_cast<const char*>(kv) + r256((size_t)L * w->inner * 4) + r256((size_t)w->inner * round4(L) * 4);
  return attention_body(w, P, ldp, M, m_dev, nullptr, nullptr, 1, kv, kvh2, L, out, ldo, workspace, nullptr, stream);
}

// ---- all batch items in one chain of launches (the batched captured plan; model/resunet.py:237-273 loops over the items) ----------
// kv of B images with L tokens each: the fp16 hi/lo K / V^T the attention kernel reads.
extern "C" size_t imf_attention_kv_batched_bytes(int32_t L, int32_t B) { return r256(imf_flash_kv_h2_bytes(L, B)); }
extern "C" size_t imf_attention_kv_batched_workspace_bytes(int32_t L, int32_t dim, int32_t inner, int32_t B) {
  return r256((size_t)B * L * dim * 4) + r256((size_t)B * L * 2 * inner * 4);          // LN(tokens), [K | V] fp32
}
// tokens [B*L, dim] row-major (image b = rows [b*L, (b+1)*L)) -> kv = fp16 hi/lo {K, V^T} per image, with
// [K | V] = LN_c(tokens) . Wkv^T as ONE GEMM over all images.  inner must be 128 (the IMFNet head).
extern "C" int imf_attention_kv_batched(const imf_attn_weights_t* w, const imf_attn_packed_t* wp, const float* tokens, int32_t L, int32_t B,
                                        void* kv, void* workspace, size_t workspace_bytes, int32_t* err, cudaStream_t stream) {
  IMF_CHECK_ARG(w != nullptr && L >= 1 && B >= 1 && B <= 256 && w->dim % 32 == 0 && w->dim <= 1024 && w->inner == 128);
  IMF_CHECK_ARG(tokens != nullptr && kv != nullptr && workspace != nullptr);
  IMF_CHECK_ARG(workspace_bytes >= imf_attention_kv_batched_workspace_bytes(L, w->dim, w->inner, B));
  const int n = B * L, dim = w->dim, inner = w->inner;
  float* cn = reinterpret_cast<float*>(workspace);
  float* KV = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + r256((size_t)n * dim * 4));
  int rc;
  if (wp != nullptr && dim % 64 == 0) {          // h2 tier: LayerNorm writes fp16 hi/lo, the projection is a TMA-fed tcgen05 GEMM
    k_layernorm_rows<<<(n + 7) / 8, 256, 0, stream>>>(tokens, dim, n, dim, w->ln_c_w, w->ln_c_b, 1e-5f, nullptr, 2 * dim, nullptr,
                                                      reinterpret_cast<__half*>(cn));
    IMF_CHECK_LAUNCH();
    if ((rc = imf_h2_gemm(cn, 2 * dim, n, nullptr, wp->wkv, 2 * inner, dim, 1.f / wp->mkv, nullptr, nullptr, 0, 0, KV, 2 * inner, err, stream))) return rc;
  } else {
    k_layernorm_rows<<<(n + 7) / 8, 256, 0, stream>>>(tokens, dim, n, dim, w->ln_c_w, w->ln_c_b, 1e-5f, cn, dim, nullptr);
    IMF_CHECK_LAUNCH();
    if ((rc = imf_tc_gemm(cn, dim, w->wkv, dim, KV, 2 * inner, n, 2 * inner, dim, 1.f, nullptr, nullptr, 0, 0, nullptr, 0, err, stream))) return rc;
  }
  return imf_flash_pack_kv(KV, 2 * inner, KV + inner, 2 * inner, 0, L, B, kv, stream);
}

extern "C" size_t imf_attention_batched_workspace_bytes(int32_t M, int32_t L, int32_t latent, int32_t inner, int32_t B) {
  return r256((size_t)M * latent * 4) + r256((size_t)M * inner * 4) + r256((size_t)M * inner * 4) + r256((size_t)M * latent * 4) +
         r256((size_t)M * latent * 4 * 4) + r256(imf_tc_gemm_workspace_bytes(M, latent, 0)) + r256((size_t)M * inner * 4) +
         r256(imf_flash_workspace_bytes(M, L, B));
}
// out[row] for the rows [seg[b], seg[b] + cnt[b]) of every item b < B of P [M, latent] (M = capacity; *m_dev = rows in use = the end of
// the last item): one LayerNorm / projection / feed-forward chain over all rows + one attention launch over all items.
// err (optional device int): in-kernel watchdog codes and the fp16-range flag of the query pack (bit 16).
extern "C" int imf_attention_fusion_fwd_batched(const imf_attn_weights_t* w, const imf_attn_packed_t* wp, const float* P, int32_t ldp, int32_t M,
                                                const int32_t* m_dev, const int32_t* seg_dev, const int32_t* cnt_dev, int32_t B, const void* kv,
                                                int32_t L, float* out, int32_t ldo, void* workspace, size_t workspace_bytes, int32_t* err,
                                                cudaStream_t stream) {
  IMF_CHECK_ARG(w != nullptr && M >= 0 && L >= 1 && B >= 1 && B <= 256);
  IMF_CHECK_ARG(w->latent % 32 == 0 && w->latent <= 1024 && w->inner == 128);
  if (M == 0) return IMF_OK;
  IMF_CHECK_ARG(P != nullptr && kv != nullptr && out != nullptr && workspace != nullptr && ldp >= w->latent && ldo >= w->latent);
  IMF_CHECK_ARG(seg_dev != nullptr && cnt_dev != nullptr);
  IMF_CHECK_ARG(workspace_bytes >= imf_attention_batched_workspace_bytes(M, L, w->latent, w->inner, B));
  if (wp != nullptr && w->latent % 128 == 0 && ldp % 4 == 0 && ldo % 4 == 0)
    return attention_body_h2(w, wp, P, ldp, M, m_dev, seg_dev, cnt_dev, B, kv, L, out, ldo, workspace, err, stream);
  return attention_body(w, P, ldp, M, m_dev, seg_dev, cnt_dev, B, nullptr, kv, L, out, ldo, workspace, err, stream);
}

// Plain dense helper for the 1x1 MinkowskiConvolution module path (kernel [Cin,Cout], optional bias [Cout]).
extern "C" int imf_linear_fwd(const float* X, int32_t ldx, const float* W_kn, const float* bias, int32_t M, int32_t Cin,
                              int32_t Cout, float* Y, int32_t ldy, cudaStream_t stream) {
  IMF_CHECK_ARG(M >= 0 && Cin > 0 && Cout > 0 && ldx >= Cin && ldy >= Cout);
  if (M == 0) return IMF_OK;
  IMF_CHECK_ARG(X != nullptr && W_kn != nullptr && Y != nullptr);
  return sgemm<true, false>(X, ldx, W_kn, Cout, Y, ldy, M, Cout, Cin, 1.f, bias, nullptr, 0, stream);
}
