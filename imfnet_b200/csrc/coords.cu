// imfnet_b200 -- coordinate kernels: voxel hash, first-occurrence unique / stride maps, neighbour tables.
//
// Replaces (for the IMFNet descriptor path) what MinkowskiEngine 0.5.4's CoordinateManager does beneath
//   ME.SparseTensor(feats, coordinates=...)          /root/reference/util/misc.py:95
//   every stride-2 MinkowskiConvolution             /root/reference/model/resunet.py:54-85
//   kernel-map generation of every convolution      /root/reference/model/resunet.py:168-226
//   ME.utils.sparse_quantize(..., return_index=True) /root/reference/util/misc.py:83
// Semantics are those of oracle/sparse_ops.py (unique_first, stride_coords, neighbour_table); all results
// here are integers and must be bit-exact against it.
//
// Layout: a hash table is `capacity` 16-byte slots {u64 key, i32 val, i32 pad}; key = packed (b,x,y,z),
// val = row index in the coordinate set.  Coordinate sets are int32 [N,4] row-major.
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"
#include "dense_grid.cuh"
#include <stdarg.h>
#include <atomic>

// ------------------------------------------------------------------------------------------------
// error string (thread-local so concurrent host threads do not clobber each other)
static thread_local char g_err[512] = "";
void imf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* imf_last_error(void) { return g_err; }
extern "C" int imf_version(void) { return 100; }

static std::atomic<long long> g_launches{0};
cudaError_t imf_set_max_smem_once(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::vector<std::pair<const void*, int>> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  for (const auto& kd : done)
    if (kd.first == kernel && kd.second == dev) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done.emplace_back(kernel, dev);
  return e;
}

// Number of SMs of the current device (cached per device): sizes the persistent grids and the split heuristics; 148 on a full B200,
// fewer on other sm_100 SKUs or MIG slices.  Without a device (the CPU-side ABI tests) it reports the B200 count.
int imf_sm_count() {
  static std::mutex mu;
  static std::vector<std::pair<int, int>> known;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 148; }
  std::lock_guard<std::mutex> lock(mu);
  for (const auto& d : known)
    if (d.first == dev) return d.second;
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
  known.emplace_back(dev, n);
  return n;
}
extern "C" int imf_device_sm_count(void) { return imf_sm_count(); }

void imf_note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" long long imf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int imf_floor_div(int a, int s) { return (a >= 0) ? (a / s) : -((-a + s - 1) / s); }

__device__ __forceinline__ int imf_count(const int* n_ptr, int n_max) {
  if (n_ptr == nullptr) return n_max;
  int n = *n_ptr;
  return n < n_max ? n : n_max;
}

// Insert unique rows: val = row index.  Duplicates / out-of-range rows raise status bits.
__global__ void k_hash_insert_unique(const int4* __restrict__ coords, const int* __restrict__ n_ptr, int n_max,
                                     ImfSlot* table, unsigned long long mask, int* status) {
  const int n = imf_count(n_ptr, n_max);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 c = coords[i];
  if (!imf_coord_in_range(c.x, c.y, c.z, c.w)) {
    atomicOr(status, IMF_STATUS_COORD_RANGE);
    return;
  }
  const unsigned long long key = imf_pack_key(c.x, c.y, c.z, c.w);
  unsigned long long slot = imf_hash64(key) & mask;
  for (unsigned long long probes = 0; probes <= mask; ++probes) {
    const unsigned long long prev = atomicCAS(&table[slot].key, IMF_EMPTY_KEY, key);
    if (prev == IMF_EMPTY_KEY) {
      table[slot].val = i;
      return;
    }
    if (prev == key) {
      atomicOr(status, IMF_STATUS_DUPLICATE);
      return;
    }
    slot = (slot + 1) & mask;
  }
  atomicOr(status, IMF_STATUS_TABLE_FULL);
}

// Insert floor(c/stride)*stride keys, keeping the MINIMUM source row per key (first occurrence).
// slot_of[i] remembers where row i's key lives so later passes do not probe again.
__global__ void k_stride_insert_min(const int4* __restrict__ coords, const int* __restrict__ n_ptr, int n_max, int stride,
                                    ImfSlot* table, unsigned long long mask, int* __restrict__ slot_of, int* status) {
  const int n = imf_count(n_ptr, n_max);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int4 c = coords[i];
  c.y = imf_floor_div(c.y, stride) * stride;
  c.z = imf_floor_div(c.z, stride) * stride;
  c.w = imf_floor_div(c.w, stride) * stride;
  slot_of[i] = -1;
  if (!imf_coord_in_range(c.x, c.y, c.z, c.w)) {
    atomicOr(status, IMF_STATUS_COORD_RANGE);
    return;
  }
  const unsigned long long key = imf_pack_key(c.x, c.y, c.z, c.w);
  unsigned long long slot = imf_hash64(key) & mask;
  for (unsigned long long probes = 0; probes <= mask; ++probes) {
    const unsigned long long prev = atomicCAS(&table[slot].key, IMF_EMPTY_KEY, key);
    if (prev == IMF_EMPTY_KEY || prev == key) {
      atomicMin(reinterpret_cast<unsigned int*>(&table[slot].val), (unsigned int)i);   // val starts at 0xFFFFFFFF
      slot_of[i] = (int)slot;
      return;
    }
    slot = (slot + 1) & mask;
  }
  atomicOr(status, IMF_STATUS_TABLE_FULL);
}

// flag[i] = 1 iff row i is the first occurrence of its key; block_count[b] = number of flags in block b.
__global__ void __launch_bounds__(1024) k_flag_first(const int* __restrict__ slot_of, const ImfSlot* __restrict__ table,
                                                     const int* __restrict__ n_ptr, int n_max,
                                                     unsigned char* __restrict__ flag, int* __restrict__ block_count) {
  __shared__ int warp_cnt[32];
  const int n = imf_count(n_ptr, n_max);
  const int i = blockIdx.x * 1024 + threadIdx.x;
  bool first = false;
  if (i < n) {
    const int s = slot_of[i];
    first = (s >= 0) && (table[s].val == i);
    flag[i] = first ? 1 : 0;
  }
  const unsigned b = __ballot_sync(0xffffffffu, first);
  if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = __popc(b);
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = warp_cnt[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) block_count[blockIdx.x] = v;
  }
}

// Ordered compaction: rank(i) = #flags before i.  Writes the strided coordinates of each first row at its
// rank, rewrites the table value to that rank, and (optionally) first_idx[rank] = i.  The last block
// publishes the total.
__global__ void __launch_bounds__(1024) k_compact_first(const int4* __restrict__ coords, const int* __restrict__ slot_of,
                                                        const unsigned char* __restrict__ flag,
                                                        const int* __restrict__ block_count, const int* __restrict__ n_ptr,
                                                        int n_max, int stride, ImfSlot* table, int4* __restrict__ coords_out,
                                                        int* __restrict__ first_idx, int* __restrict__ n_out) {
  __shared__ int warp_cnt[32];
  __shared__ int red[32];
  __shared__ int block_base;
  const int n = imf_count(n_ptr, n_max);
  // base = sum of counts of all preceding blocks
  int part = 0;
  for (int j = threadIdx.x; j < (int)blockIdx.x; j += 1024) part += block_count[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;

  const int i = blockIdx.x * 1024 + threadIdx.x;
  const bool first = (i < n) && flag[i];
  const unsigned b = __ballot_sync(0xffffffffu, first);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_cnt[warp] = __popc(b);
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    // exclusive scan of warp counts
    int c = warp_cnt[threadIdx.x];
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (threadIdx.x >= o) incl += t;
    }
    warp_cnt[threadIdx.x] = incl - c;
    if (threadIdx.x == 31) {
      // incl = total flags in this block
      if (blockIdx.x == gridDim.x - 1) *n_out = v + incl;
    }
    if (threadIdx.x == 0) block_base = v;
  }
  __syncthreads();
  if (first) {
    const int rank = block_base + warp_cnt[warp] + __popc(b & ((1u << lane) - 1u));
    int4 c = coords[i];
    c.y = imf_floor_div(c.y, stride) * stride;
    c.z = imf_floor_div(c.z, stride) * stride;
    c.w = imf_floor_div(c.w, stride) * stride;
    coords_out[rank] = c;
    table[slot_of[i]].val = rank;
    if (first_idx) first_idx[rank] = i;
  }
}

// nbr[o*K3 + k] = row of the IN set at C_out[o] + off_k * scale (scale may be negative: transposed conv), or -1.
// k = kx + K*ky + K*K*kz, off = (kx,ky,kz) - K/2  (x fastest; oracle/sparse_ops.py::kernel_offsets).
__global__ void k_kernel_map(const int4* __restrict__ out_coords, const int* __restrict__ n_ptr, int n_max,
                             const ImfSlot* __restrict__ table, unsigned long long mask, int K, int scale,
                             int* __restrict__ nbr) {
  const int n = imf_count(n_ptr, n_max);
  const int K3 = K * K * K;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * K3) return;
  const int o = (int)(idx / K3), k = (int)(idx % K3);
  const int h = K / 2;
  const int kx = k % K - h, ky = (k / K) % K - h, kz = k / (K * K) - h;
  const int4 c = out_coords[o];
  const int x = c.y + kx * scale, y = c.z + ky * scale, z = c.w + kz * scale;
  int r = -1;
  if (imf_coord_in_range(c.x, x, y, z)) r = imf_table_lookup(table, mask, imf_pack_key(c.x, x, y, z));
  nbr[idx] = r;
}

// Offset-major neighbour tables for the persistent convolution (sparse_conv_g4.cu): nbr_t[k*ld_n + o]; rows o in
// [n, roundup128(n)) are -1 (they pad the last tile); tile_mask[o/128] has bit k set when any row of that 128-row tile has a
// neighbour at offset k.  One CTA = one 128-row tile of one table (blockIdx.y = job): every thread probes the K^3 offsets of
// its row, so the tile mask is a plain store (no memset, no atomics) and several tables share one launch.
__device__ __forceinline__ int imf_parity_class(const int4& c, int t) {
  return (((c.y / t) & 1)) | (((c.z / t) & 1) << 1) | (((c.w / t) & 1) << 2);
}

struct KmapJob {
  const int4* out_coords;
  const int* n_ptr;
  const ImfSlot* table;
  int* nbr_t;
  unsigned* tile_mask;
  const int* perm;       // optional: table row o describes output row perm[o] (parity-grouped transposed convolution)
  int scale;
  int pad;
  const imf_dense::CfMeta* dense_meta;      // optional: dense row-index grid over the INPUT coordinate set (see below)
  const int* dense_cells;
};
struct KmapJobs {
  KmapJob j[16];
};
// KT = kernel size as a compile-time constant (3: every 3x3x3 geometry of the network; 0: read K at run time).  With a run-time K
// the offset decomposition costs three integer divisions per neighbour and the loop cannot be unrolled: the launch (10 tables, ~25 M
// lookups per batch of ten fragments) was bound by ALU issue (sm 75 %, 182 us: profiles/r02/call43_ncu_other_kernels.txt).
template <int KT>
__global__ void __launch_bounds__(128) k_kernel_map_t(const __grid_constant__ KmapJobs jobs, int n_max, unsigned long long mask, int K_rt, int ld_n) {
  const int K = KT ? KT : K_rt;
  __shared__ unsigned wmask[4];
  const KmapJob& jb = jobs.j[blockIdx.y];
  const int n = imf_count(jb.n_ptr, n_max);
  const int tile = blockIdx.x;
  if (tile * 128 >= n) {
    if (tile * 128 < n_max + 128 && threadIdx.x == 0) jb.tile_mask[tile] = 0u;     // tiles past the data: empty masks
    return;
  }
  const int o = tile * 128 + threadIdx.x;
  const int K3 = K * K * K, h = K / 2;
  int4 c = make_int4(0, 0, 0, 0);
  if (o < n) c = jb.out_coords[jb.perm ? jb.perm[o] : o];
  const bool transposed = jb.perm != nullptr && jb.scale < 0 && K == 3;      // parity-grouped tables are stride-2 transposed by contract
  const int pc = transposed ? imf_parity_class(c, -jb.scale) : 0;
  // Dense path: when the input set's row-index grid exists (conv_first_tc.cu built it for conv1 and left it populated), a neighbour is
  // ONE 4-byte load at a computed cell -- x-neighbours share a 32-byte sector -- instead of a hash probe chain of 16-byte slots
  // (two thirds of a forward's ~25 M probes look up the stride-1 table; the launch was bound by L2 sectors).  A cell outside the
  // item's box (boxes include a halo of 2) has no voxel; rows with a batch index the grid does not know take the hash path.
  const bool dense = jb.dense_meta != nullptr && jb.dense_meta->use_grid && o < n && (unsigned)c.x < (unsigned)jb.dense_meta->pad[2];
  int gx0 = 0, gy0 = 0, gz0 = 0, gdx = 0, gdy = 0, gdz = 0;
  long long gbase = 0;
  if (dense) {
    const int* it = jb.dense_meta->item[c.x];
    gx0 = it[0]; gy0 = it[1]; gz0 = it[2]; gdx = it[3]; gdy = it[4]; gdz = it[5];
    gbase = ((long long)(unsigned)it[6]) | ((long long)it[7] << 32);
  }
  // centre cell of the dense grid (cells are addressed with 32-bit offsets relative to it: the grid budget is far below 2^31 cells)
  const int lx0 = c.y - gx0, ly0 = c.z - gy0, lz0 = c.w - gz0;
  const int* cell0 = dense ? jb.dense_cells + gbase + ((long long)lz0 * gdy + ly0) * gdx + lx0 : nullptr;
  const int sx = jb.scale, sy = jb.scale * gdx, sz = jb.scale * gdx * gdy;
  unsigned mine = 0u;
#pragma unroll
  for (int k = 0; k < (KT ? KT * KT * KT : K3); ++k) {
    int r = -1;
    if (o < n) {
      const int kx = k % K - h, ky = (k / K) % K - h, kz = k / (K * K) - h;
      // transposed table (scale < 0, stride-2 parents): a parent can only sit at offsets that are non-zero exactly on the odd axes
      const bool possible = !transposed || (((kx != 0) == ((pc & 1) != 0)) && ((ky != 0) == ((pc & 2) != 0)) && ((kz != 0) == ((pc & 4) != 0)));
      if (possible) {
        const int x = c.y + kx * jb.scale, y = c.z + ky * jb.scale, z = c.w + kz * jb.scale;
        if (dense) {
          const int lx = x - gx0, ly = y - gy0, lz = z - gz0;
          if ((unsigned)lx < (unsigned)gdx && (unsigned)ly < (unsigned)gdy && (unsigned)lz < (unsigned)gdz)
            r = __ldg(cell0 + (kz * sz + ky * sy + kx * sx)) - 1;
        } else if (imf_coord_in_range(c.x, x, y, z)) {
          r = imf_table_lookup(jb.table, mask, imf_pack_key(c.x, x, y, z));
        }
      }
    }
    if (o < ld_n) jb.nbr_t[(size_t)k * ld_n + o] = r;
    mine |= (r >= 0 ? 1u : 0u) << k;
  }
  mine = __reduce_or_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0) wmask[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x == 0) jb.tile_mask[tile] = wmask[0] | wmask[1] | wmask[2] | wmask[3];
}

// ---- parity grouping of a coordinate set (for transposed convolutions) ---------------------------------------------------
// A voxel f of the fine set (coordinates multiples of t) can only have coarse parents (multiples of 2t) at the offsets whose
// non-zero components sit exactly on the axes where f/t is odd: 8 parity classes with 1, 2, 2, 2, 4, 4, 4, 8 candidate offsets
// instead of 27.  perm lists the rows grouped by class (stable inside a class), so a 128-row tile of the permuted order walks
// 1-8 offsets instead of all 27.  Three small kernels: per-block histogram, scan, stable scatter.
__global__ void __launch_bounds__(256) k_parity_hist(const int4* __restrict__ coords, const int* __restrict__ n_ptr, int n_max, int t,
                                                     int* __restrict__ hist /*[blocks][8]*/) {
  __shared__ int h[8];
  const int n = imf_count(n_ptr, n_max);
  if (threadIdx.x < 8) h[threadIdx.x] = 0;
  __syncthreads();
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) atomicAdd(&h[imf_parity_class(coords[i], t)], 1);
  __syncthreads();
  if (threadIdx.x < 8) hist[blockIdx.x * 8 + threadIdx.x] = h[threadIdx.x];
}
// offsets[b][c] = rows of classes < c (all blocks) + rows of class c in blocks < b.  One block of 256 threads; thread t owns a
// contiguous range of blocks with all 8 classes (two 16-byte loads per block, all independent), then a block-wide scan per class.
// (The first form carried a warp scan and a global-load latency through every group of 32 blocks: 32 us for 2000 blocks.)
__global__ void __launch_bounds__(256) k_parity_scan(const int* __restrict__ hist, int blocks, int* __restrict__ offs) {
  __shared__ int wsum[8][8];      // [warp][class]
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int per = (blocks + 255) / 256;
  const int b0 = min(blocks, t * per), b1 = min(blocks, b0 + per);
  int sum[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) sum[c] = 0;
#pragma unroll 4
  for (int b = b0; b < b1; ++b) {
    const int4 v0 = __ldg(reinterpret_cast<const int4*>(hist + b * 8)), v1 = __ldg(reinterpret_cast<const int4*>(hist + b * 8 + 4));
    sum[0] += v0.x; sum[1] += v0.y; sum[2] += v0.z; sum[3] += v0.w;
    sum[4] += v1.x; sum[5] += v1.y; sum[6] += v1.z; sum[7] += v1.w;
  }
  int run[8];                     // exclusive prefix of this thread inside its class
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    int incl = sum[c];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    if (lane == 31) wsum[w][c] = incl;
    run[c] = incl - sum[c];
  }
  __syncthreads();
  int base = 0;                   // rows of all lower classes
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    int before = 0, total = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { const int v = wsum[k][c]; total += v; if (k < w) before += v; }
    run[c] += base + before;
    base += total;
  }
#pragma unroll 4
  for (int b = b0; b < b1; ++b) {
    const int4 v0 = __ldg(reinterpret_cast<const int4*>(hist + b * 8)), v1 = __ldg(reinterpret_cast<const int4*>(hist + b * 8 + 4));
    *reinterpret_cast<int4*>(offs + b * 8) = make_int4(run[0], run[1], run[2], run[3]);
    *reinterpret_cast<int4*>(offs + b * 8 + 4) = make_int4(run[4], run[5], run[6], run[7]);
    run[0] += v0.x; run[1] += v0.y; run[2] += v0.z; run[3] += v0.w;
    run[4] += v1.x; run[5] += v1.y; run[6] += v1.z; run[7] += v1.w;
  }
}
__global__ void __launch_bounds__(256) k_parity_scatter(const int4* __restrict__ coords, const int* __restrict__ n_ptr, int n_max, int t,
                                                        const int* __restrict__ offs, int* __restrict__ perm) {
  __shared__ int wcnt[8][8];      // [warp][class]
  const int n = imf_count(n_ptr, n_max);
  const int i = blockIdx.x * 256 + threadIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cls = i < n ? imf_parity_class(coords[i], t) : -1;
  int rank_in_warp = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const unsigned m = __ballot_sync(0xffffffffu, cls == c);
    if (cls == c) rank_in_warp = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) wcnt[warp][c] = __popc(m);
  }
  __syncthreads();
  if (cls >= 0) {
    int before = 0;
    for (int w = 0; w < warp; ++w) before += wcnt[w][cls];
    perm[offs[blockIdx.x * 8 + cls] + before + rank_in_warp] = i;
  }
}

// xyz (float64 [N,3]) -> int32 (b, floor(x/voxel), floor(y/voxel), floor(z/voxel)).  IEEE double division and
// floor are exact, so this equals numpy's np.floor(xyz / voxel_size) (/root/reference/util/misc.py:82).
__global__ void k_quantize_points(const double* __restrict__ xyz, int n, double voxel, int batch, int4* __restrict__ coords) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = floor(xyz[3 * (size_t)i + 0] / voxel), y = floor(xyz[3 * (size_t)i + 1] / voxel),
               z = floor(xyz[3 * (size_t)i + 2] / voxel);
  coords[i] = make_int4(batch, (int)x, (int)y, (int)z);
}

// float32 clouds: numpy computes np.floor(xyz / voxel_size) in the INPUT dtype (float32 array / Python float -> float32 quotient with
// the divisor rounded to float32), so points next to a voxel boundary can land in another voxel than in float64; __fdiv_rn is the
// correctly rounded IEEE division numpy performs (plain `/` may compile to an approximate sequence under fast-math flags).
__global__ void k_quantize_points_f32(const float* __restrict__ xyz, int n, float voxel, int batch, int4* __restrict__ coords) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = floorf(__fdiv_rn(xyz[3 * (size_t)i + 0], voxel)), y = floorf(__fdiv_rn(xyz[3 * (size_t)i + 1], voxel)),
              z = floorf(__fdiv_rn(xyz[3 * (size_t)i + 2], voxel));
  coords[i] = make_int4(batch, (int)x, (int)y, (int)z);
}

// seg[b] = first row whose batch index is >= b (rows are batch-sorted); seg[B] = n.
__global__ void k_batch_segments(const int4* __restrict__ coords, const int* __restrict__ n_ptr, int n_max, int B,
                                 int* __restrict__ seg) {
  const int n = imf_count(n_ptr, n_max);
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > B) return;
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (coords[mid].x < b) lo = mid + 1; else hi = mid;
  }
  seg[b] = lo;
}

// ------------------------------------------------------------------------------------------------
static bool is_pow2(long long v) { return v > 0 && (v & (v - 1)) == 0; }

extern "C" long long imf_hash_capacity(long long n) {
  long long c = 1024;
  while (c < 2 * n) c <<= 1;
  return c;
}
extern "C" size_t imf_hash_bytes(long long capacity) { return (size_t)capacity * sizeof(ImfSlot); }

extern "C" int imf_hash_clear(void* table, long long capacity, cudaStream_t stream) {
  IMF_CHECK_ARG(table != nullptr && is_pow2(capacity));
  IMF_CHECK_CUDA(cudaMemsetAsync(table, 0xFF, imf_hash_bytes(capacity), stream));
  return IMF_OK;
}

extern "C" int imf_hash_build(const int32_t* coords, const int32_t* n_dev, int32_t n_max, void* table, long long capacity,
                              int32_t* status, cudaStream_t stream) {
  IMF_CHECK_ARG(table != nullptr && status != nullptr && is_pow2(capacity) && n_max >= 0 && capacity >= 2LL * n_max);
  IMF_CHECK_CUDA(cudaMemsetAsync(table, 0xFF, imf_hash_bytes(capacity), stream));
  if (n_max == 0) return IMF_OK;
  IMF_CHECK_ARG(coords != nullptr);
  k_hash_insert_unique<<<(n_max + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const int4*>(coords), n_dev, n_max,
                                                                reinterpret_cast<ImfSlot*>(table),
                                                                (unsigned long long)capacity - 1, status);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

extern "C" size_t imf_stride_map_workspace_bytes(int32_t n_in_max) {
  const size_t n = (size_t)(n_in_max > 0 ? n_in_max : 1);
  const size_t blocks = (n + 1023) / 1024;
  // slot_of[int n] + block_count[int blocks] + flag[u8 n], each rounded to 256 B
  auto r = [](size_t b) { return (b + 255) / 256 * 256; };
  return r(n * 4) + r(blocks * 4) + r(n);
}

extern "C" int imf_stride_map(const int32_t* coords_in, const int32_t* n_in_dev, int32_t n_in_max, int32_t stride,
                              void* table_out, long long capacity, int32_t* coords_out, int32_t* n_out_dev,
                              int32_t* first_idx, void* workspace, size_t workspace_bytes, int32_t* status,
                              cudaStream_t stream) {
  IMF_CHECK_ARG(table_out != nullptr && status != nullptr && n_out_dev != nullptr && coords_out != nullptr);
  IMF_CHECK_ARG(is_pow2(capacity) && n_in_max >= 0 && capacity >= 2LL * n_in_max && stride >= 1);
  IMF_CHECK_ARG(workspace_bytes >= imf_stride_map_workspace_bytes(n_in_max));
  IMF_CHECK_CUDA(cudaMemsetAsync(table_out, 0xFF, imf_hash_bytes(capacity), stream));
  if (n_in_max == 0) {
    IMF_CHECK_CUDA(cudaMemsetAsync(n_out_dev, 0, sizeof(int32_t), stream));
    return IMF_OK;
  }
  IMF_CHECK_ARG(coords_in != nullptr && workspace != nullptr);
  const size_t n = (size_t)n_in_max;
  const int blocks = (int)((n + 1023) / 1024);
  auto r = [](size_t b) { return (b + 255) / 256 * 256; };
  char* ws = reinterpret_cast<char*>(workspace);
  int* slot_of = reinterpret_cast<int*>(ws);
  int* block_count = reinterpret_cast<int*>(ws + r(n * 4));
  unsigned char* flag = reinterpret_cast<unsigned char*>(ws + r(n * 4) + r((size_t)blocks * 4));
  ImfSlot* table = reinterpret_cast<ImfSlot*>(table_out);
  const unsigned long long mask = (unsigned long long)capacity - 1;
  const int4* cin = reinterpret_cast<const int4*>(coords_in);
  k_stride_insert_min<<<(n_in_max + 255) / 256, 256, 0, stream>>>(cin, n_in_dev, n_in_max, stride, table, mask, slot_of, status);
  IMF_CHECK_LAUNCH();
  k_flag_first<<<blocks, 1024, 0, stream>>>(slot_of, table, n_in_dev, n_in_max, flag, block_count);
  IMF_CHECK_LAUNCH();
  k_compact_first<<<blocks, 1024, 0, stream>>>(cin, slot_of, flag, block_count, n_in_dev, n_in_max, stride, table,
                                               reinterpret_cast<int4*>(coords_out), first_idx, n_out_dev);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

extern "C" int imf_kernel_map(const int32_t* out_coords, const int32_t* n_out_dev, int32_t n_out_max, const void* table_in,
                              long long capacity, int32_t kernel_size, int32_t scale, int32_t* nbr, cudaStream_t stream) {
  IMF_CHECK_ARG(is_pow2(capacity) && table_in != nullptr && n_out_max >= 0);
  IMF_CHECK_ARG(kernel_size >= 1 && (kernel_size & 1) == 1 && kernel_size <= 7);
  if (n_out_max == 0) return IMF_OK;
  IMF_CHECK_ARG(out_coords != nullptr && nbr != nullptr);
  const long long total = (long long)n_out_max * kernel_size * kernel_size * kernel_size;
  const long long blocks = (total + 255) / 256;
  IMF_CHECK_ARG(blocks < 0x7FFFFFFFLL);
  k_kernel_map<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const int4*>(out_coords), n_out_dev, n_out_max,
                                                     reinterpret_cast<const ImfSlot*>(table_in),
                                                     (unsigned long long)capacity - 1, kernel_size, scale, nbr);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

struct imf_kmap_job_t {      // mirrors include/imfnet_b200.h
  const int32_t* out_coords;
  const int32_t* n_out_dev;
  const void* table_in;
  int32_t* nbr_t;
  uint32_t* tile_mask;
  const int32_t* perm;
  int32_t scale;
  const void* dense_meta;
  const void* dense_cells;
};

extern "C" int imf_kernel_map_t_batch(const imf_kmap_job_t* jobs, int32_t njobs, int32_t n_out_max, long long capacity, int32_t kernel_size,
                                      int32_t ld_n, cudaStream_t stream) {
  IMF_CHECK_ARG(jobs != nullptr && njobs >= 1 && njobs <= 16 && is_pow2(capacity) && n_out_max >= 0);
  IMF_CHECK_ARG(kernel_size >= 1 && (kernel_size & 1) == 1 && kernel_size <= 3);
  IMF_CHECK_ARG(ld_n % 4 == 0 && ld_n >= ((n_out_max + 31) & ~31));
  if (n_out_max == 0) return IMF_OK;
  KmapJobs kj;
  for (int i = 0; i < njobs; ++i) {
    IMF_CHECK_ARG(jobs[i].out_coords != nullptr && jobs[i].table_in != nullptr && jobs[i].nbr_t != nullptr && jobs[i].tile_mask != nullptr);
    kj.j[i].out_coords = reinterpret_cast<const int4*>(jobs[i].out_coords);
    kj.j[i].n_ptr = jobs[i].n_out_dev;
    kj.j[i].table = reinterpret_cast<const ImfSlot*>(jobs[i].table_in);
    kj.j[i].nbr_t = jobs[i].nbr_t;
    kj.j[i].tile_mask = jobs[i].tile_mask;
    kj.j[i].perm = jobs[i].perm;
    kj.j[i].scale = jobs[i].scale;
    kj.j[i].pad = 0;
    IMF_CHECK_ARG((jobs[i].dense_meta == nullptr) == (jobs[i].dense_cells == nullptr));
    kj.j[i].dense_meta = reinterpret_cast<const imf_dense::CfMeta*>(jobs[i].dense_meta);
    kj.j[i].dense_cells = reinterpret_cast<const int*>(jobs[i].dense_cells);
  }
  const int tiles = (n_out_max + 127) / 128;
  dim3 grid(tiles + 1, njobs);          // + 1: the mask entry past the last tile is written (zero) too
  if (kernel_size == 3) k_kernel_map_t<3><<<grid, 128, 0, stream>>>(kj, n_out_max, (unsigned long long)capacity - 1, kernel_size, ld_n);
  else k_kernel_map_t<0><<<grid, 128, 0, stream>>>(kj, n_out_max, (unsigned long long)capacity - 1, kernel_size, ld_n);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

extern "C" int imf_kernel_map_t(const int32_t* out_coords, const int32_t* n_out_dev, int32_t n_out_max, const void* table_in,
                                long long capacity, int32_t kernel_size, int32_t scale, int32_t* nbr_t, int32_t ld_n,
                                uint32_t* tile_mask, cudaStream_t stream) {
  imf_kmap_job_t job{out_coords, n_out_dev, table_in, nbr_t, tile_mask, nullptr, scale, nullptr, nullptr};
  return imf_kernel_map_t_batch(&job, 1, n_out_max, capacity, kernel_size, ld_n, stream);
}

extern "C" size_t imf_parity_perm_workspace_bytes(int32_t n_max) { return (size_t)((n_max > 0 ? n_max : 1) + 255) / 256 * 8 * 2 * sizeof(int); }

extern "C" int imf_parity_perm(const int32_t* coords, const int32_t* n_dev, int32_t n_max, int32_t tensor_stride, int32_t* perm,
                               void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  IMF_CHECK_ARG(n_max >= 0 && tensor_stride >= 1);
  if (n_max == 0) return IMF_OK;
  IMF_CHECK_ARG(coords != nullptr && perm != nullptr && workspace != nullptr && workspace_bytes >= imf_parity_perm_workspace_bytes(n_max));
  const int blocks = (n_max + 255) / 256;
  int* hist = reinterpret_cast<int*>(workspace);
  int* offs = hist + (size_t)blocks * 8;
  k_parity_hist<<<blocks, 256, 0, stream>>>(reinterpret_cast<const int4*>(coords), n_dev, n_max, tensor_stride, hist);
  IMF_CHECK_LAUNCH();
  k_parity_scan<<<1, 256, 0, stream>>>(hist, blocks, offs);
  IMF_CHECK_LAUNCH();
  k_parity_scatter<<<blocks, 256, 0, stream>>>(reinterpret_cast<const int4*>(coords), n_dev, n_max, tensor_stride, offs, perm);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

extern "C" int imf_quantize_points(const double* xyz, int32_t n, double voxel_size, int32_t batch_index, int32_t* coords,
                                   cudaStream_t stream) {
  IMF_CHECK_ARG(n >= 0 && voxel_size > 0.0);
  if (n == 0) return IMF_OK;
  IMF_CHECK_ARG(xyz != nullptr && coords != nullptr);
  k_quantize_points<<<(n + 255) / 256, 256, 0, stream>>>(xyz, n, voxel_size, batch_index, reinterpret_cast<int4*>(coords));
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

extern "C" int imf_quantize_points_f32(const float* xyz, int32_t n, float voxel_size, int32_t batch_index, int32_t* coords,
                                       cudaStream_t stream) {
  IMF_CHECK_ARG(n >= 0 && voxel_size > 0.f);
  if (n == 0) return IMF_OK;
  IMF_CHECK_ARG(xyz != nullptr && coords != nullptr);
  k_quantize_points_f32<<<(n + 255) / 256, 256, 0, stream>>>(xyz, n, voxel_size, batch_index, reinterpret_cast<int4*>(coords));
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}

extern "C" int imf_batch_segments(const int32_t* coords, const int32_t* n_dev, int32_t n_max, int32_t num_batches,
                                  int32_t* seg, cudaStream_t stream) {
  IMF_CHECK_ARG(seg != nullptr && num_batches >= 0 && n_max >= 0);
  IMF_CHECK_ARG(coords != nullptr || n_max == 0);
  k_batch_segments<<<(num_batches + 1 + 127) / 128, 128, 0, stream>>>(reinterpret_cast<const int4*>(coords), n_dev, n_max,
                                                                      num_batches, seg);
  IMF_CHECK_LAUNCH();
  return IMF_OK;
}
