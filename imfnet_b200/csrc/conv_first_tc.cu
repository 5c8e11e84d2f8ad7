// imfnet_b200 -- first layer (conv1: K^3 = 5^3 offsets, ONE input channel) on the tensor cores.
//
//   conv1 = ME.MinkowskiConvolution(in_channels=1, out_channels=32, kernel_size=5)  + norm1     /root/reference/model/resunet.py:42-49, 168-169
//   input feature = a column of ones at inference (util/misc.py:76-77), any [N,1] in general.
//
// Y[o, :] = sum_k f[nbr(o,k)] * W[k, 0, :] is a dense product E[N, K^3] . W[K^3, Cout] once the K^3 neighbour FEATURES of every voxel
// are laid out as a row: E[o, k] = f[nbr(o,k)] (0 where the neighbour is absent).  The old kernel (k_conv_first, sparse_conv.cu)
// probed the hash table 125 times per voxel and then walked the ~40 present offsets serially (shuffle + FMA per offset and lane):
// 965 us for 10 x 50 k voxels, 23 % of the sparse part of a batched forward.  Here
//   1. the FEATURES of every batch item's voxels are scattered into a DENSE grid over the item's bounding box (+ a halo of K/2 cells,
//      so neighbour reads need no range checks; empty cells hold 0, which contributes nothing): a neighbour's feature becomes one
//      4-byte load at a computed address -- no hash, no probe chain, no second gather through a row index -- and the K cells of an
//      x-run are contiguous, so one warp instruction covers 32 / K whole runs (6 cache lines instead of 25 for K = 5).  The grid
//      lives in the caller's workspace (budget: 512 cells per voxel -- a 3 m room at 2.5 cm is ~200^3 cells for 50 k voxels; measured
//      on B200: 341 us for conv1 of 10 x 50 k voxels against 1185 us through the hash); when the boxes do not fit the budget (sparse
//      outdoor scans at a fine voxel size) the same kernel probes the hash table and gathers the feature through the row index;
//   2. k_cf_expand writes E as an h2 matrix (fp16 hi/lo, K^3 padded to a multiple of 64 columns) + an identity "neighbour table";
//   3. the persistent tcgen05 convolution kernel (sparse_conv_g4.cu) runs the product as a one-offset convolution over E, with the
//      BatchNorm affine in its epilogue -- the accumulation over offsets happens in TMEM, not in a per-voxel loop.
#include <cuda_fp16.h>

#include <climits>
#include <cstdlib>

#include "common.cuh"
#include "dense_grid.cuh"

extern "C" int imf_sparse_conv_g4_fwd(const void* X, int32_t ldx, int32_t kc_in, const void* packed, const int32_t* nbr_t, int32_t ld_n,
                                      const uint32_t* tile_mask, const int32_t* n_out_dev, int32_t n_out_max, int32_t kernel_volume,
                                      int32_t Cin, int32_t Cout, const float* scale, const float* shift, const void* residual, int32_t ldr,
                                      int32_t kc_r, int32_t relu, void* Y, int32_t ldy, int32_t n_y_rows, int32_t kc_out, void* workspace,
                                      size_t workspace_bytes, int32_t* err, cudaStream_t stream);

namespace {

using imf_dense::CfMeta;
using imf_dense::cf_cell;
using imf_dense::kMaxItems;

__device__ __forceinline__ int cf_count(const int* n_ptr, int n_max) {
  if (!n_ptr) return n_max;
  const int v = *n_ptr;
  return v < n_max ? v : n_max;
}

__global__ void k_cf_init(CfMeta* m, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * 6) m->bbox[i / 6][i % 6] = (i % 6) < 3 ? INT_MAX : INT_MIN;
  if (i == 0) m->use_grid = 0;
}

// bounding box per batch item: block-level reduction when the block's voxels belong to one item (rows are batch-sorted, so almost
// every block), per-thread atomics otherwise
__global__ void __launch_bounds__(256) k_cf_bbox(const int4* __restrict__ coords, const int* __restrict__ n_ptr, int n_max, int B, CfMeta* m) {
  __shared__ int red[6][8];
  __shared__ int same_s;
  const int n = cf_count(n_ptr, n_max);
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int first = blockIdx.x * 256;
  if (first >= n) return;
  const int last = min(n, first + 256) - 1;
  const int b0 = coords[first].x, b1 = coords[last].x;
  int4 c = make_int4(b0, INT_MAX, INT_MAX, INT_MAX);
  const bool valid = i < n;
  if (valid) c = coords[i];
  if (threadIdx.x == 0) same_s = (b0 == b1) ? 1 : 0;
  __syncthreads();
  if (!same_s) {
    if (valid && (unsigned)c.x < (unsigned)B) {
      atomicMin(&m->bbox[c.x][0], c.y); atomicMin(&m->bbox[c.x][1], c.z); atomicMin(&m->bbox[c.x][2], c.w);
      atomicMax(&m->bbox[c.x][3], c.y); atomicMax(&m->bbox[c.x][4], c.z); atomicMax(&m->bbox[c.x][5], c.w);
    }
    return;
  }
  int v[6] = {valid ? c.y : INT_MAX, valid ? c.z : INT_MAX, valid ? c.w : INT_MAX, valid ? c.y : INT_MIN, valid ? c.z : INT_MIN, valid ? c.w : INT_MIN};
#pragma unroll
  for (int q = 0; q < 6; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int t = __shfl_xor_sync(0xffffffffu, v[q], o);
      v[q] = q < 3 ? min(v[q], t) : max(v[q], t);
    }
    if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v[q];
  }
  __syncthreads();
  if (threadIdx.x < 6 && (unsigned)b0 < (unsigned)B) {
    const int q = threadIdx.x;
    int r = red[q][0];
    for (int w = 1; w < 8; ++w) r = q < 3 ? min(r, red[q][w]) : max(r, red[q][w]);
    if (q < 3) atomicMin(&m->bbox[b0][q], r); else atomicMax(&m->bbox[b0][q], r);
  }
}

// boxes -> grid layout; decides whether the grid path is used (one thread: B <= 256 items)
__global__ void k_cf_layout(CfMeta* m, int B, int halo, long long budget_cells) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  long long total = 0;
  bool ok = true;
  for (int b = 0; b < B; ++b) {
    int* it = m->item[b];
    const int* bb = m->bbox[b];
    if (bb[0] > bb[3]) {                 // empty item
      it[0] = it[1] = it[2] = 0; it[3] = it[4] = it[5] = 0; it[6] = it[7] = 0;
      continue;
    }
    const long long dx = (long long)bb[3] - bb[0] + 1 + 2 * halo, dy = (long long)bb[4] - bb[1] + 1 + 2 * halo,
                    dz = (long long)bb[5] - bb[2] + 1 + 2 * halo;
    it[0] = bb[0] - halo; it[1] = bb[1] - halo; it[2] = bb[2] - halo;
    it[3] = (int)dx; it[4] = (int)dy; it[5] = (int)dz;
    it[6] = (int)(total & 0xffffffffll); it[7] = (int)(total >> 32);
    const long long cells = dx * dy * dz;
    if (dx > 65536 || dy > 65536 || dz > 65536 || cells > budget_cells) { ok = false; break; }
    total += cells;
    if (total > budget_cells) { ok = false; break; }
  }
  m->use_grid = ok ? 1 : 0;
  m->pad[0] = (int)(total & 0xffffffffll);
  m->pad[1] = (int)(total >> 32);
  m->pad[2] = B;
}

__global__ void __launch_bounds__(256) k_cf_clear(const CfMeta* __restrict__ m, int4* __restrict__ grid) {
  if (!m->use_grid) return;
  const long long total = ((long long)(unsigned)m->pad[0]) | ((long long)m->pad[1] << 32);
  const long long n4 = (total + 3) / 4;
  const int4 e = make_int4(0, 0, 0, 0);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) grid[i] = e;
}

// cell of voxel i <- its feature / i + 1 (release != 0: <- 0, which leaves the grids all-empty again without clearing ~200 MB each)
// Two grids of the same geometry: `grid` holds the voxel's FEATURE (what conv1's expansion reads: one load per neighbour, no second
// gather through a row index -- measured: +50 % on the expansion), `rows` holds row + 1 (what the neighbour tables read).
__global__ void __launch_bounds__(256) k_cf_scatter(const float* __restrict__ X, int ldx, const int4* __restrict__ coords,
                                                    const int* __restrict__ n_ptr, int n_max, int B, const CfMeta* __restrict__ m,
                                                    float* __restrict__ grid, int* __restrict__ rows, int release) {
  if (!m->use_grid) return;
  const int n = cf_count(n_ptr, n_max);
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int4 c = coords[i];
  if ((unsigned)c.x >= (unsigned)B) return;          // foreign batch index: such rows take the hash probe in k_cf_expand
  const long long cell = cf_cell(m->item[c.x], c.y, c.z, c.w);
  grid[cell] = release ? 0.f : X[(size_t)i * ldx];
  rows[cell] = release ? 0 : i + 1;
}

// E[row, k] = f[nbr(row, k)] as an h2 matrix of KP columns (chunk width 64); identity table + tile masks for the one-offset convolution.
// A warp owns 32 consecutive voxels and walks them one by one.  Offset k = kx + K ky + K^2 kz (x fastest): the K cells of an x-run are
// contiguous in the grid, so in instruction i lane l reads offset k = (K * RPI) i + l -- RPI = 32 / K whole runs per instruction -- and
// everything that depends only on (lane, i) (the offset's dx, dy, dz) is computed once per warp, not per voxel: per voxel and offset
// there is one address add, one load, the fp16 hi/lo split and two 2-byte stores into the warp's row buffer in shared memory, from
// which the row leaves as ONE 512-byte store (a first version issued ~400 warp instructions per voxel and was issue-bound at 850 us
// for 500 k voxels).  Grid cells are addressed with 32-bit indices (the budget is far below 2^31 cells).
template <int K>
__global__ void __launch_bounds__(256) k_cf_expand(const float* __restrict__ X, int ldx, const int4* __restrict__ coords,
                                                   const int* __restrict__ n_ptr, int n_max, int B, const CfMeta* __restrict__ m,
                                                   const float* __restrict__ grid, const ImfSlot* __restrict__ table, unsigned long long mask,
                                                   __half* __restrict__ E, int* __restrict__ ident, unsigned* __restrict__ tile_mask,
                                                   int ld_n) {
  constexpr int K3 = K * K * K, h = K / 2, KR = K * (32 / K), KP = (K3 + 63) / 64 * 64, NI = (KP + KR - 1) / KR;
  __shared__ __align__(16) __half rowbuf[8][2 * KP];
  const int n = cf_count(n_ptr, n_max);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int chunk0 = (blockIdx.x * 8 + w) * 32;
  if (chunk0 >= ld_n) return;
  const int n128 = (n + 127) / 128 * 128;
  // ---- this lane's voxel of the chunk: identity-table entry, tile mask, grid cell of its centre ----
  const int myrow = chunk0 + lane;
  int4 c = make_int4(-1, 0, 0, 0);
  if (myrow < n) c = coords[myrow];
  if (myrow < ld_n) {
    if (myrow < n) ident[myrow] = myrow;
    else if (myrow < n128) ident[myrow] = -1;           // padding rows up to the 128-row boundary the convolution kernel reads
    if ((myrow & 127) == 0) tile_mask[myrow >> 7] = myrow < n ? 1u : 0u;
  }
  const int grid_ok = m->use_grid;
  int my_ug = 0, my_base = 0, my_dx = 0, my_dxy = 0;
  if (grid_ok && myrow < n && (unsigned)c.x < (unsigned)B) {
    const int* it = m->item[c.x];
    my_ug = 1;
    my_base = (int)cf_cell(it, c.y, c.z, c.w);
    my_dx = it[3];
    my_dxy = it[3] * it[4];
  }
  // ---- this lane's offsets ----
  int odx[NI], ody[NI], odz[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int k = lane + KR * i;
    odx[i] = k % K - h;
    ody[i] = (k / K) % K - h;
    odz[i] = k / (K * K) - h;
  }
  __half* rb = rowbuf[w];
  const int vend = min(32, n - chunk0);
  for (int v = 0; v < vend; ++v) {
    const int ug = __shfl_sync(0xffffffffu, my_ug, v);
    const int base = __shfl_sync(0xffffffffu, my_base, v);
    const int DX = __shfl_sync(0xffffffffu, my_dx, v), DXY = __shfl_sync(0xffffffffu, my_dxy, v);
    int cb = 0, cx = 0, cy = 0, cz = 0;
    if (!ug) {                                           // hash-probe path (grid over budget, or a foreign batch index): warp-uniform
      cb = __shfl_sync(0xffffffffu, c.x, v); cx = __shfl_sync(0xffffffffu, c.y, v);
      cy = __shfl_sync(0xffffffffu, c.z, v); cz = __shfl_sync(0xffffffffu, c.w, v);
    }
    if (lane < KR) {
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int k = lane + KR * i;
        if (k >= KP) break;
        float f = 0.f;
        if (k < K3) {
          if (ug) {
            f = __ldg(grid + base + odz[i] * DXY + ody[i] * DX + odx[i]);
          } else {
            const int x = cx + odx[i], y = cy + ody[i], z = cz + odz[i];
            if (imf_coord_in_range(cb, x, y, z)) {
              const int r = imf_table_lookup(table, mask, imf_pack_key(cb, x, y, z));
              if (r >= 0) f = __ldg(X + (size_t)r * ldx);
            }
          }
        }
        const __half hi = __float2half_rn(f);
        __half* p = rb + (k >> 6) * 128 + (k & 63);     // chunk k / 64 holds [hi 64 | lo 64]
        p[0] = hi;
        p[64] = __float2half_rn(f - __half2float(hi));
      }
    }
    __syncwarp();
    if (lane * 8 < 2 * KP)
      *reinterpret_cast<uint4*>(E + (size_t)(chunk0 + v) * (2 * KP) + lane * 8) = *reinterpret_cast<const uint4*>(rb + lane * 8);
    __syncwarp();
  }
}

inline size_t r256(size_t b) { return (b + 255) / 256 * 256; }
inline int cf_kp(int K) { return (K * K * K + 63) / 64 * 64; }
inline long long cf_budget_cells(int n_max) {          // cells per voxel (default 512; IMF_CF_CELLS_PER_VOXEL overrides), below 2^31 (k_cf_expand addresses cells with 32-bit indices)
  static const long long per = [] { const char* e = getenv("IMF_CF_CELLS_PER_VOXEL"); const long long v = e ? atoll(e) : 0; return v > 0 ? v : 512LL; }();
  const long long c = per * (n_max > 0 ? n_max : 1) + (1 << 20);
  return c < 2000000000LL ? c : 2000000000LL;
}

struct CfLayout {
  size_t meta, grid, rows, E, ident, mask, conv_ws, total;
  int ld_n;
};
inline CfLayout cf_layout(int n_max, int K) {
  CfLayout L;
  const int KP = cf_kp(K);
  L.ld_n = (n_max + 127) / 128 * 128;
  size_t off = 0;
  L.meta = off; off += r256(sizeof(CfMeta));
  L.grid = off; off += r256((size_t)cf_budget_cells(n_max) * 4 + 16);
  L.rows = off; off += r256((size_t)cf_budget_cells(n_max) * 4 + 16);
  L.E = off;    off += r256((size_t)L.ld_n * 2 * KP * sizeof(__half));
  L.ident = off; off += r256((size_t)L.ld_n * 4);
  L.mask = off; off += r256((size_t)(L.ld_n / 128 + 2) * 4);
  L.total = off;
  return L;
}

}  // namespace

extern "C" int32_t imf_conv_first_tc_columns(int32_t kernel_size) { return cf_kp(kernel_size); }
extern "C" size_t imf_conv_first_tc_workspace_bytes(int32_t n_max, int32_t kernel_size) { return cf_layout(n_max, kernel_size).total; }

namespace {
int conv_first_tc_run(const float* X, int32_t ldx, const void* packed, const int32_t* coords, const int32_t* n_dev, int32_t n_max,
                      int32_t num_items, const void* table, long long capacity, int32_t kernel_size, int32_t Cout, const float* scale,
                      const float* shift, int32_t relu, void* Y, int32_t ldy, int32_t kc_out, void* workspace, size_t workspace_bytes,
                      int32_t* err, cudaStream_t stream, bool clear_first) {
  IMF_CHECK_ARG(n_max >= 0 && (kernel_size == 1 || kernel_size == 3 || kernel_size == 5) && num_items >= 1 && num_items <= kMaxItems);
  IMF_CHECK_ARG(ldx >= 1 && (Cout == 32 || Cout == 64 || Cout == 128) && scale != nullptr && shift != nullptr);
  IMF_CHECK_ARG(capacity > 0 && (capacity & (capacity - 1)) == 0);
  if (n_max == 0) return IMF_OK;
  IMF_CHECK_ARG(X != nullptr && packed != nullptr && coords != nullptr && table != nullptr && Y != nullptr && workspace != nullptr);
  const CfLayout L = cf_layout(n_max, kernel_size);
  IMF_CHECK_ARG(workspace_bytes >= L.total && ((uintptr_t)workspace % 256) == 0);
  char* ws = reinterpret_cast<char*>(workspace);
  CfMeta* meta = reinterpret_cast<CfMeta*>(ws + L.meta);
  float* grid = reinterpret_cast<float*>(ws + L.grid);
  int* rows = reinterpret_cast<int*>(ws + L.rows);
  __half* E = reinterpret_cast<__half*>(ws + L.E);
  int* ident = reinterpret_cast<int*>(ws + L.ident);
  unsigned* tmask = reinterpret_cast<unsigned*>(ws + L.mask);
  const int4* c4 = reinterpret_cast<const int4*>(coords);
  const int KP = cf_kp(kernel_size);
  const int blocks = (n_max + 255) / 256;
  k_cf_init<<<(num_items * 6 + 255) / 256, 256, 0, stream>>>(meta, num_items);
  IMF_CHECK_LAUNCH();
  k_cf_bbox<<<blocks, 256, 0, stream>>>(c4, n_dev, n_max, num_items, meta);
  IMF_CHECK_LAUNCH();
  k_cf_layout<<<1, 32, 0, stream>>>(meta, num_items, kernel_size / 2, cf_budget_cells(n_max));
  IMF_CHECK_LAUNCH();
  if (clear_first) {
    k_cf_clear<<<imf_sm_count() * 8, 256, 0, stream>>>(meta, reinterpret_cast<int4*>(grid));
    IMF_CHECK_LAUNCH();
    k_cf_clear<<<imf_sm_count() * 8, 256, 0, stream>>>(meta, reinterpret_cast<int4*>(rows));
    IMF_CHECK_LAUNCH();
  }
  k_cf_scatter<<<blocks, 256, 0, stream>>>(X, ldx, c4, n_dev, n_max, num_items, meta, grid, rows, 0);
  IMF_CHECK_LAUNCH();
  const ImfSlot* tab = reinterpret_cast<const ImfSlot*>(table);
  const unsigned long long hmask = (unsigned long long)capacity - 1;
  const int eblocks = (L.ld_n + 255) / 256;          // 8 warps x 32 voxels per block
  if (kernel_size == 5) k_cf_expand<5><<<eblocks, 256, 0, stream>>>(X, ldx, c4, n_dev, n_max, num_items, meta, grid, tab, hmask, E, ident, tmask, L.ld_n);
  else if (kernel_size == 3) k_cf_expand<3><<<eblocks, 256, 0, stream>>>(X, ldx, c4, n_dev, n_max, num_items, meta, grid, tab, hmask, E, ident, tmask, L.ld_n);
  else k_cf_expand<1><<<eblocks, 256, 0, stream>>>(X, ldx, c4, n_dev, n_max, num_items, meta, grid, tab, hmask, E, ident, tmask, L.ld_n);
  IMF_CHECK_LAUNCH();
  return imf_sparse_conv_g4_fwd(E, 2 * KP, 64, packed, ident, L.ld_n, tmask, n_dev, n_max, 1, KP, Cout, scale, shift, nullptr, 0, 0, relu, Y,
                                ldy, n_max, kc_out, nullptr, 0, err, stream);
}
}  // namespace

// conv1 (+ folded BatchNorm) for ONE input channel through the tensor-core tier.  packed = imf_sparse_conv_h2_pack of the kernel
// reshaped to ONE offset with K^3 (zero-padded to imf_conv_first_tc_columns) input channels; scale / shift as for imf_sparse_conv_g4_fwd.
// coords carry the batch index in column 0 (< num_items); table / capacity = the hash table of the same coordinate set (fallback when
// the items' bounding boxes exceed the workspace's dense-grid budget of 512 cells per voxel).  Y = h2 matrix (ldy halves, chunk kc_out).
// Self-contained: the grid region of the workspace is cleared first (any workspace content is fine) and left populated.
extern "C" int imf_conv_first_tc_h2_fwd(const float* X, int32_t ldx, const void* packed, const int32_t* coords, const int32_t* n_dev,
                                        int32_t n_max, int32_t num_items, const void* table, long long capacity, int32_t kernel_size,
                                        int32_t Cout, const float* scale, const float* shift, int32_t relu, void* Y, int32_t ldy,
                                        int32_t kc_out, void* workspace, size_t workspace_bytes, int32_t* err, cudaStream_t stream) {
  return conv_first_tc_run(X, ldx, packed, coords, n_dev, n_max, num_items, table, capacity, kernel_size, Cout, scale, shift, relu, Y, ldy, kc_out,
                           workspace, workspace_bytes, err, stream, true);
}

// The plans' form: the workspace was zero-initialised ONCE by the caller and every use is followed by imf_conv_first_tc_release, so
// the two ~200 MB grids are never cleared (45 us each per batch of ten fragments); between the two calls the populated row grid
// (cell = row + 1) also serves imf_kernel_map_t_batch (jobs' dense_meta / dense_cells, see imf_conv_first_tc_grid).
extern "C" int imf_conv_first_tc_h2_fwd_keep(const float* X, int32_t ldx, const void* packed, const int32_t* coords, const int32_t* n_dev,
                                             int32_t n_max, int32_t num_items, const void* table, long long capacity, int32_t kernel_size,
                                             int32_t Cout, const float* scale, const float* shift, int32_t relu, void* Y, int32_t ldy,
                                             int32_t kc_out, void* workspace, size_t workspace_bytes, int32_t* err, cudaStream_t stream) {
  return conv_first_tc_run(X, ldx, packed, coords, n_dev, n_max, num_items, table, capacity, kernel_size, Cout, scale, shift, relu, Y, ldy, kc_out,
                           workspace, workspace_bytes, err, stream, false);
}

// empties the cells the coordinates occupy (same coords / counts / num_items / workspace as the _keep call before)
extern "C" int imf_conv_first_tc_release(const int32_t* coords, const int32_t* n_dev, int32_t n_max, int32_t num_items, int32_t kernel_size,
                                         void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  IMF_CHECK_ARG(n_max >= 0 && num_items >= 1 && num_items <= kMaxItems);
  if (n_max == 0) return IMF_OK;
  const CfLayout L = cf_layout(n_max, kernel_size);
  IMF_CHECK_ARG(coords != nullptr && workspace != nullptr && workspace_bytes >= L.total);
  char* ws = reinterpret_cast<char*>(workspace);
  k_cf_scatter<<<(n_max + 255) / 256, 256, 0, stream>>>(nullptr, 0, reinterpret_cast<const int4*>(coords), n_dev, n_max, num_items,
                                                        reinterpret_cast<const CfMeta*>(ws + L.meta), reinterpret_cast<float*>(ws + L.grid),
                                                        reinterpret_cast<int*>(ws + L.rows), 1);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

// addresses of the grid's header and cells inside a workspace of imf_conv_first_tc_h2_fwd(_keep) (for imf_kmap_job_t)
extern "C" int imf_conv_first_tc_grid(void* workspace, int32_t n_max, int32_t kernel_size, const void** meta, const void** cells) {
  IMF_CHECK_ARG(workspace != nullptr && meta != nullptr && cells != nullptr && n_max >= 0);
  const CfLayout L = cf_layout(n_max, kernel_size);
  *meta = reinterpret_cast<char*>(workspace) + L.meta;
  *cells = reinterpret_cast<char*>(workspace) + L.rows;
  return IMF_OK;
}
