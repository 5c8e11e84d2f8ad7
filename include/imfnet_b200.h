/* imfnet_b200 -- C ABI of the B200 (sm_100a) descriptor-extraction kernels.
 *
 * The reference (XiaoshuiHuang/IMFNet) has no FFI of its own: its boundary for this path is the Python
 * nn.Module API of model/ over the MinkowskiEngine Python API.  The host-side mirror
 * in imfnet_b200/ keeps that Python API; everything below it goes through these entry points, which are
 * what a MinkowskiEngine-style backend would bind.  Each entry cites the reference call site it serves.
 *
 * Conventions
 *   - plain pointers + sizes; all pointers are DEVICE pointers unless stated; no allocation inside, the
 *     caller passes workspaces (size queries are provided); work is enqueued on `stream` and not synchronised;
 *   - return 0 on success, negative on failure (-1 bad argument, -2 CUDA error, -3 unsupported shape);
 *     imf_last_error() returns a thread-local message for the last failure;
 *   - coordinates are int32 [N,4] rows (batch, x, y, z), 16-byte aligned; features are fp32 row-major with an
 *     explicit leading dimension (elements) so concatenations are column windows of one buffer;
 *   - `*_dev` row counts are optional device int32 scalars: when non-NULL the kernels use
 *     min(*n_dev, n_max) rows, so a caller can chain levels without reading sizes back to the host;
 *   - coordinate range: batch in [0, 65534], |x|,|y|,|z| < 32768 (packed into one 64-bit key); violations,
 *     duplicate rows and hash overflow set bits in a device `status` word (IMF_STATUS_*).
 */
#ifndef IMFNET_B200_H_
#define IMFNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* imf_stream_t; /* == cudaStream_t */

#define IMF_STATUS_COORD_RANGE 1
#define IMF_STATUS_DUPLICATE 2
#define IMF_STATUS_TABLE_FULL 4

const char* imf_last_error(void);
int imf_version(void);
/* Number of kernels this library has launched in the calling process so far (bench.py's gpu_launches). */
long long imf_launch_count(void);
/* SMs of the current CUDA device (148 on a full B200; 148 is also reported when no device is present).  The persistent kernels size
 * their grids and split heuristics with it; callers use it for the same thresholds (e.g. "fewer 128-row tiles than SMs"). */
int imf_device_sm_count(void);

/* ---- coordinates: ME CoordinateManager work behind ME.SparseTensor(...) (util/misc.py:95) ------------- */

/* Slots (power of two, >= 2n) and bytes of a hash table for n coordinates. */
long long imf_hash_capacity(long long n);
size_t imf_hash_bytes(long long capacity);
int imf_hash_clear(void* table, long long capacity, imf_stream_t stream);

/* Build the coordinate -> row table of a set of UNIQUE coordinates (value = row index). */
int imf_hash_build(const int32_t* coords, const int32_t* n_dev, int32_t n_max, void* table, long long capacity,
                   int32_t* status, imf_stream_t stream);

/* Coarser coordinate set of a stride-2 convolution (model/resunet.py:54-85), or, with stride 1, the
 * first-occurrence de-duplication of ME.utils.sparse_quantize (util/misc.py:83):
 *   coords_out = unique_first(floor(coords_in / stride) * stride), rows in order of first appearance;
 *   table_out maps those coordinates to their row; *n_out_dev = number of rows;
 *   first_idx (optional, [n_in_max]) receives the source row of every output row. */
size_t imf_stride_map_workspace_bytes(int32_t n_in_max);
int imf_stride_map(const int32_t* coords_in, const int32_t* n_in_dev, int32_t n_in_max, int32_t stride, void* table_out,
                   long long capacity, int32_t* coords_out, int32_t* n_out_dev, int32_t* first_idx, void* workspace,
                   size_t workspace_bytes, int32_t* status, imf_stream_t stream);

/* Neighbour table of one convolution (ME kernel map, output-stationary form):
 *   nbr[o*K^3 + k] = row of the input set at out_coords[o] + off_k*scale, or -1;
 *   k = kx + K*ky + K^2*kz, off = (kx,ky,kz) - K/2.  Forward conv: scale = input tensor stride.
 *   Transposed conv (model/resunet.py:101-134): out = fine set, table = coarse set, scale = -(fine stride). */
int imf_kernel_map(const int32_t* out_coords, const int32_t* n_out_dev, int32_t n_out_max, const void* table_in,
                   long long capacity, int32_t kernel_size, int32_t scale, int32_t* nbr, imf_stream_t stream);

/* The same neighbour table in offset-major form for the TMA-gather convolution: nbr_t[k*ld_n + o] (ld_n % 4 == 0,
 * ld_n >= n_out_max rounded up to 32; rows in [n, roundup128(n)) are written as -1), plus tile_mask[o/128] (uint32, at least
 * ceil(n_out_max/128)+1 entries, all written here) whose bit k says that some row of that 128-row tile has a neighbour at offset k.
 * kernel_size in {1, 3}. */
int imf_kernel_map_t(const int32_t* out_coords, const int32_t* n_out_dev, int32_t n_out_max, const void* table_in,
                     long long capacity, int32_t kernel_size, int32_t scale, int32_t* nbr_t, int32_t ld_n, uint32_t* tile_mask,
                     imf_stream_t stream);

/* Several such tables (same n_out_max, capacity, kernel_size, ld_n) in ONE launch; `jobs` is a HOST array of njobs <= 16 entries. */
typedef struct {
  const int32_t* out_coords;
  const int32_t* n_out_dev;
  const void* table_in;
  int32_t* nbr_t;
  uint32_t* tile_mask;
  const int32_t* perm; /* optional: table row o describes output row perm[o] (see imf_parity_perm); only for the STRIDE-2 transposed
                          convolution (scale = -fine stride): offsets impossible for the row's parity class are not probed */
  int32_t scale;
  const void* dense_meta;  /* optional (both or neither): header and cells of the dense row-index grid over the INPUT coordinate set of */
  const void* dense_cells; /* this table (imf_conv_first_tc_grid of a workspace populated by imf_conv_first_tc_h2_fwd_keep): neighbours are
                              then read from the grid (one 4-byte load) instead of probing table_in; same result */
} imf_kmap_job_t;
int imf_kernel_map_t_batch(const imf_kmap_job_t* jobs, int32_t njobs, int32_t n_out_max, long long capacity, int32_t kernel_size,
                           int32_t ld_n, imf_stream_t stream);

/* Rows of a coordinate set at tensor stride t grouped by the parity class of (x/t, y/t, z/t) (stable inside a class): perm[v] = row.
 * A fine voxel can only have coarse parents at the offsets whose non-zero components sit on its odd axes, so tiles of the permuted
 * order need 1, 2, 4 or 8 of the 27 offsets of a stride-2 transposed convolution (model/resunet.py:101-134). */
size_t imf_parity_perm_workspace_bytes(int32_t n_max);
int imf_parity_perm(const int32_t* coords, const int32_t* n_dev, int32_t n_max, int32_t tensor_stride, int32_t* perm, void* workspace,
                    size_t workspace_bytes, imf_stream_t stream);

/* coords[i] = (batch_index, floor(xyz[i]/voxel_size)) in float64, as util/misc.py:82 computes on the host. */
int imf_quantize_points(const double* xyz, int32_t n, double voxel_size, int32_t batch_index, int32_t* coords,
                        imf_stream_t stream);
/* The same for float32 clouds: numpy evaluates np.floor(xyz / voxel_size) in the input dtype, i.e. a float32 division by
 * (float)voxel_size; points next to a voxel boundary may fall into another voxel than in float64. */
int imf_quantize_points_f32(const float* xyz, int32_t n, float voxel_size, int32_t batch_index, int32_t* coords,
                            imf_stream_t stream);

/* seg[b] = first row with batch index >= b, seg[num_batches] = n (rows are batch-sorted; resunet.py:240-255). */
int imf_batch_segments(const int32_t* coords, const int32_t* n_dev, int32_t n_max, int32_t num_batches, int32_t* seg,
                       imf_stream_t stream);

/* Device-side form of the per-item split of ResUNet2.transformer (model/resunet.py:240-255) for the batched captured plan
 * (imfnet_b200/batched.py): seg[0..num_batches] as above, cnt[b] = min(seg[b+1] - seg[b], cap_item).  *err (optional) gets
 * bit 17 when an item has more than cap_item rows and bit 18 when rows carry a batch index >= num_batches.  num_batches <= 255. */
int imf_batch_segments_n(const int32_t* coords, const int32_t* n_dev, int32_t n_max, int32_t num_batches, int32_t cap_item,
                         int32_t* seg, int32_t* cnt, int32_t* err, imf_stream_t stream);


/* ---- sparse convolution: ME.MinkowskiConvolution(+Transpose).forward (model/resunet.py:168-213,
 *      model/residual_block.py:40,44) with BatchNorm(eval)/residual/ReLU fused ------------------------- */

/* Y[o, :Cout] = act( (sum_k X[nbr[o,k], :Cin] . W[k]) * scale + shift (+ residual[o]) ), W = [K^3, Cin, Cout].
 * scale/shift (per output channel, both or neither) and residual are optional; relu != 0 applies max(.,0).
 * Requires Cin % 32 == 0, Cout % 32 == 0, K^3 <= 27, ldx % 4 == 0, X and W 16-byte aligned. */
int imf_sparse_conv_fwd(const float* X, int32_t ldx, const float* W, const int32_t* nbr, const int32_t* n_out_dev,
                        int32_t n_out_max, int32_t kernel_volume, int32_t Cin, int32_t Cout, const float* scale,
                        const float* shift, const float* residual, int32_t ldr, int32_t relu, float* Y, int32_t ldy,
                        imf_stream_t stream);

/* ---- "h2" tier: activations stored as two fp16 halves (v = hi + lo), products accumulated as lo*Whi + hi*Wlo + hi*Whi by
 *      kind::f16 tcgen05 MMAs into fp32 TMEM (fp32-class accuracy at twice the TF32 MMA rate and a copy-only gather).
 * h2 matrix of C channels with chunk width KC in {32,64} (C % KC == 0), row stride ld in HALVES (>= 2C, multiple of 8):
 *   channel c = q*KC + j  ->  hi at row*ld + q*2*KC + j, lo at row*ld + q*2*KC + KC + j.
 * `err` (optional device int): bit 16 is set if a stored value left the fp16 range (|v| > 60000); codes 1..3 = watchdog. */
int imf_h2_pack(const float* X, int32_t ldx, int32_t n, int32_t C, int32_t KC, void* H, int32_t ldh, int32_t* err, imf_stream_t stream);
int imf_h2_unpack(const void* H, int32_t ldh, int32_t n, int32_t C, int32_t KC, float* X, int32_t ldx, imf_stream_t stream);
/* The same with an optional device-side row count (min(*n_dev, n) rows are converted; n sizes the launch). */
int imf_h2_pack_n(const float* X, int32_t ldx, int32_t n, const int32_t* n_dev, int32_t C, int32_t KC, void* H, int32_t ldh, int32_t* err,
                  imf_stream_t stream);
int imf_h2_unpack_n(const void* H, int32_t ldh, int32_t n, const int32_t* n_dev, int32_t C, int32_t KC, float* X, int32_t ldx,
                    imf_stream_t stream);

/* The conversions with a multiplier: H = h2(X * mul), X = (hi + lo) * mul.  The plans store activations times a power-of-two scale chosen
 * from the BatchNorm affine parameters ("numeric range"); these move values across that boundary. */
int imf_h2_pack_scaled_n(const float* X, int32_t ldx, int32_t n, const int32_t* n_dev, int32_t C, int32_t KC, float mul, void* H, int32_t ldh,
                         int32_t* err, imf_stream_t stream);
int imf_h2_unpack_scaled_n(const void* H, int32_t ldh, int32_t n, const int32_t* n_dev, int32_t C, int32_t KC, float mul, float* X, int32_t ldx,
                           imf_stream_t stream);
/* h2 [n, C] (C <= 128) -> fp32 rows, divided by their L2 norm when normalize != 0 (model/resunet.py:228-231, no epsilon); out_row
 * (optional) scatters row i to Y[out_row[i]]. */
int imf_h2_unpack_l2norm(const void* H, int32_t ldh, int32_t n, const int32_t* n_dev, int32_t C, int32_t KC, int32_t normalize,
                         const int32_t* out_row, float* Y, int32_t ldy, imf_stream_t stream);
/* Identity neighbour table (+ tile masks) that turns imf_sparse_conv_g4_fwd with kernel_volume = 1 into a 1x1 convolution / dense
 * product over the rows [0, min(*n_dev, n_max)): nbr_t[i] = i, -1 from n up to the next 128-row boundary.  ld_n % 128 == 0, >= n_max;
 * tile_mask holds ld_n / 128 + 1 words. */
int imf_identity_table(const int32_t* n_dev, int32_t n_max, int32_t* nbr_t, int32_t ld_n, uint32_t* tile_mask, imf_stream_t stream);

/* Weights W[K^3,Cin,Cout] * wmul (a power of two that brings max|W| near 2^11; fold 1/wmul into `scale`) packed for the
 * input chunk width kc_in (consumed by imf_sparse_conv_g4_fwd). */
size_t imf_sparse_conv_h2_packed_bytes(int32_t kernel_volume, int32_t Cin, int32_t Cout, int32_t kc_in);
int imf_sparse_conv_h2_pack(const float* W, int32_t kernel_volume, int32_t Cin, int32_t Cout, int32_t kc_in, float wmul, void* packed,
                            imf_stream_t stream);

/* "g4" kernel of the h2 tier (csrc/sparse_conv_g4.cu): same operation and epilogue as imf_sparse_conv_fwd on h2 matrices and
 * packed weights (scale and shift are required):
 * persistent (one CTA per SM, launch shape independent of the row count, all sizes read from n_out_dev on the device), reading the
 * offset-major table + tile masks of imf_kernel_map_t, sharing each weight slab between the sub-tiles of a CTA, and writing the output
 * with tiled TMA stores.  n_y_rows = rows of the Y allocation (tensor-map extent, >= n_out_max; rows in [n, roundup32(n)) that exist
 * may be overwritten).  workspace (optional, imf_sparse_conv_g4_workspace_bytes): NULL = one CTA per 128-row tile when n < 128 * #SMs
 * (least SM time: best throughput with several fragments in flight); non-NULL = such a small level splits each tile's offsets over
 * several CTAs into fp32 partials in the workspace + a reduce launch (shortest latency of a lone forward).  Same result up to fp32
 * summation order; each setting is bit-reproducible. */
size_t imf_sparse_conv_g4_workspace_bytes(int32_t Cout);
int imf_sparse_conv_g4_fwd(const void* X, int32_t ldx, int32_t kc_in, const void* packed, const int32_t* nbr_t, int32_t ld_n,
                           const uint32_t* tile_mask, const int32_t* n_out_dev, int32_t n_out_max, int32_t kernel_volume, int32_t Cin,
                           int32_t Cout, const float* scale, const float* shift, const void* residual, int32_t ldr, int32_t kc_r,
                           int32_t relu, void* Y, int32_t ldy, int32_t n_y_rows, int32_t kc_out, void* workspace, size_t workspace_bytes,
                           int32_t* err, imf_stream_t stream);
/* imf_sparse_conv_g4_fwd with a row permutation: table row v (and its tile masks) describe output row out_row[v], whose result is
 * written to Y[out_row[v]].  With imf_parity_perm this makes the 128-row tiles of a TRANSPOSED convolution walk only the 1-8 offsets
 * their rows can have instead of all 27.  residual must be NULL. */
int imf_sparse_conv_g4_fwd_perm(const void* X, int32_t ldx, int32_t kc_in, const void* packed, const int32_t* nbr_t, int32_t ld_n,
                                const uint32_t* tile_mask, const int32_t* n_out_dev, int32_t n_out_max, int32_t kernel_volume, int32_t Cin,
                                int32_t Cout, const float* scale, const float* shift, const void* residual, int32_t ldr, int32_t kc_r,
                                int32_t relu, void* Y, int32_t ldy, int32_t n_y_rows, int32_t kc_out, const int32_t* out_row,
                                void* workspace, size_t workspace_bytes, int32_t* err, imf_stream_t stream);
/* Profiling hook: device int64 buffer (>= 1024 entries) filled by CTA 0 with clock64() stamps (slot map in the source), an
 * override of the CTAs per output-channel tile (0 = one per SM) and of the minimum stages per split CTA (`producer_warps`, 0 keeps
 * the current value); flags: bit 0 skips the weight copies, bit 1 the gathers, bit 2 the MMAs (results are then meaningless),
 * bit 3 = one MMA-issuing warp, bit 4 = stages dealt to the MMA warps by stage number instead of by accumulator (results correct
 * but not bit-reproducible: tools/conv_g4_check.py).  NULL / 0 switch everything off. */
int imf_debug_conv_g4_trace(long long* trace, int32_t grid, int32_t producer_warps, int32_t flags);

/* First layer (conv1, model/resunet.py:42-49,168): K in {1,3,5}, Cin in {1,3,6} (ones / rgb / rgb+normal,
 * util/misc.py:66-77), Cout in {32,64,128}; neighbours are
 * probed from the hash table of the same coordinate set, no neighbour table needed. */
int imf_conv_first_fwd(const float* X, int32_t ldx, int32_t Cin, const float* W, const int32_t* coords, const int32_t* n_dev,
                       int32_t n_max, const void* table, long long capacity, int32_t kernel_size, int32_t tensor_stride,
                       int32_t Cout, const float* scale, const float* shift, int32_t relu, float* Y, int32_t ldy,
                       imf_stream_t stream);

/* conv1 with ONE input channel on the tensor cores (csrc/conv_first_tc.cu): the K^3 neighbour features of every voxel are laid out as
 * a row of an h2 matrix E[n, KP] (KP = imf_conv_first_tc_columns(K) = K^3 rounded up to 64; neighbours found through a dense row-index
 * grid over every batch item's bounding box, or through the hash table when the boxes exceed the workspace's budget of 512 cells per
 * voxel), then Y = act((E . W) * scale + shift) runs as a one-offset convolution in imf_sparse_conv_g4_fwd.
 * packed = imf_sparse_conv_h2_pack(W', 1, KP, Cout, 64, wmul) with W'[0, k, :] = kernel[k, 0, :] (rows >= K^3 zero); scale already
 * divided by wmul.  coords column 0 = batch item (< num_items <= 256; other rows fall back to the hash probe).  workspace: 256-byte
 * aligned, imf_conv_first_tc_workspace_bytes(n_max, K).  Y: h2 matrix with >= n_max rows (ldy halves, chunk width kc_out). */
int32_t imf_conv_first_tc_columns(int32_t kernel_size);
size_t imf_conv_first_tc_workspace_bytes(int32_t n_max, int32_t kernel_size);
int imf_conv_first_tc_h2_fwd(const float* X, int32_t ldx, const void* packed, const int32_t* coords, const int32_t* n_dev, int32_t n_max,
                             int32_t num_items, const void* table, long long capacity, int32_t kernel_size, int32_t Cout,
                             const float* scale, const float* shift, int32_t relu, void* Y, int32_t ldy, int32_t kc_out, void* workspace,
                             size_t workspace_bytes, int32_t* err, imf_stream_t stream);
/* The captured plans' form of the same call: the workspace was zero-initialised ONCE by the caller and every use is followed by
 * imf_conv_first_tc_release, so the ~200 MB grid is never cleared; between the two calls the populated grid (cell = row + 1) also
 * serves imf_kernel_map_t_batch (imf_kmap_job_t.dense_meta / dense_cells from imf_conv_first_tc_grid). */
int imf_conv_first_tc_h2_fwd_keep(const float* X, int32_t ldx, const void* packed, const int32_t* coords, const int32_t* n_dev, int32_t n_max,
                             int32_t num_items, const void* table, long long capacity, int32_t kernel_size, int32_t Cout,
                             const float* scale, const float* shift, int32_t relu, void* Y, int32_t ldy, int32_t kc_out, void* workspace,
                             size_t workspace_bytes, int32_t* err, imf_stream_t stream);
int imf_conv_first_tc_release(const int32_t* coords, const int32_t* n_dev, int32_t n_max, int32_t num_items, int32_t kernel_size,
                              void* workspace, size_t workspace_bytes, imf_stream_t stream);
int imf_conv_first_tc_grid(void* workspace, int32_t n_max, int32_t kernel_size, const void** meta, const void** cells);

/* imf_conv_first_fwd writing an h2 matrix (ldy in halves, chunk width kc_out). */
int imf_conv_first_h2_fwd(const float* X, int32_t ldx, int32_t Cin, const float* W, const int32_t* coords, const int32_t* n_dev,
                          int32_t n_max, const void* table, long long capacity, int32_t kernel_size, int32_t tensor_stride,
                          int32_t Cout, const float* scale, const float* shift, int32_t relu, void* Y, int32_t ldy, int32_t kc_out,
                          imf_stream_t stream);

/* conv1_tr (1x1, no bias) -> ReLU -> final (1x1 + bias) -> optional row L2 normalisation
 * (model/resunet.py:224-233).  W1 = [C0, C1], W2 = [C1, C2], b2 = [C2] or NULL; C1 in {32,64,128}, C2 <= 32. */
int imf_pointwise_tail_fwd(const float* X, int32_t ldx, int32_t C0, const float* W1, int32_t C1, const float* W2,
                           const float* b2, int32_t C2, const int32_t* n_dev, int32_t n_max, int32_t normalize, float* Y,
                           int32_t ldy, imf_stream_t stream);

/* imf_pointwise_tail_fwd reading an h2 matrix made of two sections (model/resunet.py:219 concatenation): channels [0,Ca)
 * with chunk width kca, then [Ca,C0) with chunk width kcb; ldx in halves.  out_row (optional int32 [n]) writes result
 * row i to Y[out_row[i]] (internal row order -> caller row order). */
int imf_pointwise_tail_h2_fwd(const void* X, int32_t ldx, int32_t C0, int32_t Ca, int32_t kca, int32_t kcb, const float* W1, int32_t C1,
                              const float* W2, const float* b2, int32_t C2, const int32_t* n_dev, int32_t n_max, int32_t normalize,
                              const int32_t* out_row, float* Y, int32_t ldy, imf_stream_t stream);

/* The same tail in ONE tensor-core kernel (csrc/tail_fused.cu): per 128-row tile TMA load -> conv1_tr product -> ReLU -> final
 * product -> bias -> L2 norm -> TMA store; the hidden layer and the logits never touch HBM.
 *   out[i, :] = normalize( scale2 * (relu(scale1 * (X[i, :] . W1) + shift1) . W2) + bias2 ),  i < min(*n_dev, n_max)
 * X: h2 matrix of c0 channels, chunk width 32 (ldx halves); packed1 = imf_sparse_conv_h2_pack(W1 as [1, c0, 64], kc_in 32), packed2 =
 * (W2 as [1, 64, 32], kc_in 64), their power-of-two multipliers folded into scale1 / scale2; shift1, bias2 optional; out fp32 rows of
 * ldo floats.  c0 in {32, 64, 96}, c1 == 64, c2 == 32.  Replaces model/resunet.py:216-233 (conv1_tr, MEF.relu, final, L2 norm). */
int imf_tail_fused_h2_fwd(const void* X, int32_t ldx, int32_t n_max, const int32_t* n_dev, int32_t c0, int32_t c1, int32_t c2,
                          const void* packed1, const float* scale1, const float* shift1, const void* packed2, const float* scale2,
                          const float* bias2, int32_t normalize, float* out, int32_t ldo, int32_t* err, imf_stream_t stream);

/* Y = X . W (+ bias): a 1x1 ME.MinkowskiConvolution as a module (kernel [Cin, Cout]). */
int imf_linear_fwd(const float* X, int32_t ldx, const float* W_kn, const float* bias, int32_t M, int32_t Cin, int32_t Cout,
                   float* Y, int32_t ldy, imf_stream_t stream);

/* ---- image branch: ImageEncoder.forward (model/Img_Encoder.py:15-18 -> model/resnet.py:195-216) as pixel-major h2 matrices [H*W, C]
 *      run through imf_sparse_conv_g4_fwd with closed-form neighbour tables ------------------------------------------------ */

/* Offset-major neighbour table (+ tile masks, as imf_kernel_map_t) of a ksize x ksize / stride / zero-padded 2-D convolution on an
 * Hin x Win image: nbr_t[k*ld_n + oy*Wout + ox], k = kx + ksize*ky; ld_n >= Hout*Wout rounded up to 128. */
int imf_image_conv_table(int32_t Hin, int32_t Win, int32_t ksize, int32_t stride, int32_t pad, int32_t* nbr_t, int32_t ld_n,
                         uint32_t* tile_mask, imf_stream_t stream);
/* im2col of an fp32 [C,H,W] image for the stem convolution: row = output pixel, column = c + C*(kx + ksize*ky), zero-padded to Kpad
 * (multiple of 32) columns, written as an h2 matrix with chunk width 32 (ldy in halves). */
int imf_image_im2col_h2(const float* image, int32_t C, int32_t H, int32_t W, int32_t ksize, int32_t stride, int32_t pad, int32_t Kpad,
                        void* Y, int32_t ldy, imf_stream_t stream);
/* ResNet layer1 on a plane layout ("P8", csrc/image_conv_p8.cu): a 64-channel activation of num_images H x W images is stored as 16 planes
 * per image (8 chunks of 8 channels x {hi, lo} fp16 halves) of (H + 2) x (W + 2) zero-bordered pixels, 16 bytes per pixel and plane; a
 * 3x3 / stride-1 convolution then reads each input pixel once per 16 x 8 tile and forms the nine taps' operands as shifted un-swizzled
 * views of the same shared-memory patch.  imf_image_p8_bytes: size of such a buffer (zero-initialise it once).  imf_image_maxpool_p8:
 * model/resnet.py:203 from a pixel-major h2 matrix into P8.  imf_image_conv3x3_p8_fwd: Y = act(conv(X) * scale + shift (+ residual)),
 * model/resnet.py:60-76; packed = imf_sparse_conv_h2_pack(W as [9 (tap kx + 3 ky), 64, 64], kc_in 64); Y is P8 (y_pixel_major == 0) or a
 * pixel-major h2 matrix of chunk width 64 with ldy halves. */
size_t imf_image_p8_bytes(int32_t H, int32_t W, int32_t num_images);
int imf_image_maxpool_p8(const void* X, int32_t ldx, int32_t kc, int32_t Hin, int32_t Win, int32_t ksize, int32_t stride, int32_t pad, void* Y,
                         int32_t num_images, imf_stream_t stream);
int imf_image_conv3x3_p8_fwd(const void* X, int32_t H, int32_t W, int32_t num_images, const void* packed, const float* scale,
                             const float* shift, const void* residual, int32_t relu, void* Y, int32_t y_pixel_major, int32_t ldy,
                             int32_t* err, imf_stream_t stream);

/* The ResNet stem (conv 7x7 / stride 2 / padding 3 of the 3-channel frame + BatchNorm + ReLU, model/resnet.py:195-207) as a fused
 * implicit GEMM (csrc/stem_fused.cu): no im2col matrix in HBM.  image: fp32 [num_images, 3, H, W]; packed =
 * imf_sparse_conv_h2_pack of the kernel laid out as [4 (pairs of kernel rows), 64 (per row: 8 columns kx = -1..6 x 4 channels, zeros at
 * kx = -1, channel 3 and the missing 8th row), 64] with kc_in 64 (multiplier folded into scale); Y: h2 matrix of 64 channels, chunk width 64, rows = pixels of image 0, then 1, ...;
 * workspace >= imf_image_stem_workspace_bytes (the pre-split padded image set). */
size_t imf_image_stem_workspace_bytes(int32_t H, int32_t W, int32_t num_images);
int imf_image_stem_h2_fwd(const float* image, int32_t H, int32_t W, int32_t num_images, const void* packed, const float* scale,
                          const float* shift, void* workspace, size_t workspace_bytes, void* Y, int32_t ldy, int32_t* err,
                          imf_stream_t stream);

/* Max pooling (torch semantics: padding never wins) on a pixel-major h2 matrix of C channels, chunk width kc. */
int imf_image_maxpool_h2(const void* X, int32_t ldx, int32_t kc, int32_t C, int32_t Hin, int32_t Win, int32_t ksize, int32_t stride,
                         int32_t pad, void* Y, int32_t ldy, imf_stream_t stream);
/* The two kernels above for num_images images in ONE launch: contiguous NCHW images in, the pixel rows of image b following those of
 * image b - 1 in the input / output matrices (the batched image plan). */
int imf_image_im2col_h2_batch(const float* image, int32_t C, int32_t H, int32_t W, int32_t ksize, int32_t stride, int32_t pad, int32_t Kpad,
                              void* Y, int32_t ldy, int32_t num_images, imf_stream_t stream);
int imf_image_maxpool_h2_batch(const void* X, int32_t ldx, int32_t kc, int32_t C, int32_t Hin, int32_t Win, int32_t ksize, int32_t stride,
                               int32_t pad, void* Y, int32_t ldy, int32_t num_images, imf_stream_t stream);
/* Y[C][L] = X[L][C]^T (fp32): token-major result -> the NCHW feature map ImageEncoder.forward returns. */
int imf_transpose_tokens(const float* X, int32_t L, int32_t C, float* Y, imf_stream_t stream);

/* ---- dense GEMM on the tcgen05 tensor cores (3xTF32, fp32-class accuracy): every nn.Linear / einsum of
 *      model/attention_fusion.py:57-59,79-95 --------------------------------------------------------- */

/* C[M,N] = alpha * A[M,K] . B[N,K]^T (+ bias[n]) (+ R[m,n]); A and B row-major (K contiguous).
 * geglu != 0: B holds 2N rows (value rows then gate rows, bias likewise) and C[m,n] = (.)_n * gelu((.)_{n+N}).
 * workspace (optional, imf_tc_gemm_workspace_bytes) lets small tile grids split K over more SMs.
 * err (optional device int) receives a non-zero code if an in-kernel barrier wait times out (the kernel traps). */
size_t imf_tc_gemm_workspace_bytes(int32_t M, int32_t N, int32_t K);
int imf_tc_gemm(const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc, int32_t M, int32_t N, int32_t K,
                float alpha, const float* bias, const float* R, int32_t ldr, int32_t geglu, void* workspace,
                size_t workspace_bytes, int32_t* err, imf_stream_t stream);

/* imf_tc_gemm with an optional device-side row count: only min(*m_dev, M) rows are computed; M sizes the launch and the workspace. */
int imf_tc_gemm_m(const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc, int32_t M, const int32_t* m_dev, int32_t N,
                  int32_t K, float alpha, const float* bias, const float* R, int32_t ldr, int32_t geglu, void* workspace,
                  size_t workspace_bytes, int32_t* err, imf_stream_t stream);

/* ---- attention fusion: AttentionFusion.forward (model/attention_fusion.py:132-154), depth 0, 1 head ---- */
typedef struct {
  const float *ln_q_w, *ln_q_b; /* cross_attend_blocks.0.norm            [latent]                   */
  const float *ln_c_w, *ln_c_b; /* cross_attend_blocks.0.norm_context    [dim]                      */
  const float* wq;              /* cross_attend_blocks.0.fn.to_q.weight  [inner, latent]            */
  const float* wkv;             /* cross_attend_blocks.0.fn.to_kv.weight [2*inner, dim]             */
  const float *wo, *bo;         /* cross_attend_blocks.0.fn.to_out       [latent, inner], [latent]  */
  const float *ln_f_w, *ln_f_b; /* cross_attend_blocks.1.norm            [latent]                   */
  const float *w1, *b1;         /* cross_attend_blocks.1.fn.net.0        [8*latent, latent], [8*latent] */
  const float *w2, *b2;         /* cross_attend_blocks.1.fn.net.2        [latent, 4*latent], [latent]   */
  int32_t latent, dim, inner;
} imf_attn_weights_t; /* HOST struct of DEVICE pointers */

/* kv = { K = LayerNorm_c(tokens) . Wk^T, V^T } for one image (opaque buffer of imf_attention_kv_bytes).  channel_major != 0: tokens are the encoder's
 * feature map [dim][L] (NCHW, L = H'*W'), i.e. the view/permute of model/resunet.py:259-261 is folded in;
 * channel_major == 0: tokens are row-major [L][dim] (AttentionFusion.forward's `data` argument). */
size_t imf_attention_kv_bytes(int32_t L, int32_t inner); /* size of a kv buffer: K [L,inner] then V^T [inner, roundup4(L)] */
size_t imf_attention_kv_workspace_bytes(int32_t L, int32_t dim);
int imf_attention_kv(const imf_attn_weights_t* w, const float* tokens, int32_t L, int32_t channel_major, float* kv,
                     void* workspace, size_t workspace_bytes, imf_stream_t stream);

/* out[M, latent] = cross-attention + GEGLU feed-forward of M point tokens P[M, latent] against kv. */
size_t imf_attention_workspace_bytes(int32_t M, int32_t L, int32_t latent, int32_t inner);
int imf_attention_fusion_fwd(const imf_attn_weights_t* w, const float* P, int32_t ldp, int32_t M, const float* kv, int32_t L,
                             float* out, int32_t ldo, void* workspace, size_t workspace_bytes, imf_stream_t stream);

/* ---- descriptor matching (next row): nearest neighbour in descriptor space, the kernel under the mutual-NN matching of
 *      scripts/evaluation_3dmatch.py:207-217 (util/uio.py:245-258) and lib/eval.py:18-48 ---------------------------------------------- */
/* idx[i] = argmin_j ||A[i,:C] - B[j,:C]||^2 (first index wins ties; -1 if nb == 0); d2 (optional) = that squared distance; C in {16,32,64}. */
size_t imf_nn_search_workspace_bytes(int32_t na);
int imf_nn_search(const float* A, int32_t lda, int32_t na, const float* B, int32_t ldb, int32_t nb, int32_t C, int32_t* idx, float* d2,
                  void* workspace, size_t workspace_bytes, imf_stream_t stream);
/* The same search with the distance matrix on the tensor cores (csrc/matching_tc.cu): a tcgen05 product on fp16 hi/lo operands
 * filters, per query, the candidates within a proven error margin of the minimum; those (one or two) are re-evaluated exactly as
 * imf_nn_search does, so idx / d2 are identical to imf_nn_search bit for bit, ties included.  C in {16, 32}; workspace 256-byte
 * aligned, imf_nn_search_tc_workspace_bytes(na, nb). */
size_t imf_nn_search_tc_workspace_bytes(int32_t na, int32_t nb);
int imf_nn_search_tc(const float* A, int32_t lda, int32_t na, const float* B, int32_t ldb, int32_t nb, int32_t C, int32_t* idx, float* d2,
                     void* workspace, size_t workspace_bytes, imf_stream_t stream);

/* imf_attention_fusion_fwd with an optional device-side token count (min(*m_dev, M) rows; M sizes launches and workspace). */
int imf_attention_fusion_fwd_m(const imf_attn_weights_t* w, const float* P, int32_t ldp, int32_t M, const int32_t* m_dev, const float* kv,
                               int32_t L, float* out, int32_t ldo, void* workspace, size_t workspace_bytes, imf_stream_t stream);

/* ---- the same module for ALL batch items in one chain of launches (the batched captured plan) ----------------------------------
 * ResUNet2.transformer (model/resunet.py:237-273) calls the fusion module once per batch item; everything in it but the attention
 * core is row-wise, so here the point tokens of all items (rows [seg[b], seg[b] + cnt[b]) of P, imf_batch_segments_n) go through ONE
 * LayerNorm / projection / feed-forward chain and ONE attention launch in which item b attends to the L tokens of image b.
 * Requires the IMFNet head (inner == 128).
 *   imf_attention_kv_batched: tokens [B*L, dim] row-major (image b = rows [b*L, (b+1)*L)) -> kv (opaque: fp16 hi/lo K and V^T per image);
 *   imf_attention_fusion_fwd_batched: P [M, latent] (M = row capacity, min(*m_dev, M) rows in use) -> out, rows outside every item's
 *     range untouched.  err (optional device int32): watchdog codes, bit 16 = a query left the fp16 range.
 *   wp (optional): the module's matrices pre-packed for imf_h2_gemm; with it LayerNorm writes fp16 hi/lo and every projection is a
 *     TMA-fed tcgen05 GEMM on pre-split operands; NULL = the 3xTF32 GEMMs of imf_tc_gemm (same results to fp32 rounding). */
typedef struct imf_attn_packed {
  const void *wq, *wkv, *wo, *w1, *w2; /* imf_sparse_conv_h2_pack(W^T as [1, K, N], 1, K, N, 64, m*); w1: value / gate rows interleaved per
                                          128-column tile (tile t = [value 64t..64t+63 | gate 64t..64t+63]) */
  float mq, mkv, mo, m1, m2;           /* the power-of-two scales the matrices were packed with */
} imf_attn_packed_t;
size_t imf_attention_kv_batched_bytes(int32_t L, int32_t B);
size_t imf_attention_kv_batched_workspace_bytes(int32_t L, int32_t dim, int32_t inner, int32_t B);
int imf_attention_kv_batched(const imf_attn_weights_t* w, const imf_attn_packed_t* wp, const float* tokens, int32_t L, int32_t B, void* kv,
                             void* workspace, size_t workspace_bytes, int32_t* err, imf_stream_t stream);
size_t imf_attention_batched_workspace_bytes(int32_t M, int32_t L, int32_t latent, int32_t inner, int32_t B);
int imf_attention_fusion_fwd_batched(const imf_attn_weights_t* w, const imf_attn_packed_t* wp, const float* P, int32_t ldp, int32_t M,
                                     const int32_t* m_dev, const int32_t* seg_dev, const int32_t* cnt_dev, int32_t B, const void* kv, int32_t L,
                                     float* out, int32_t ldo, void* workspace, size_t workspace_bytes, int32_t* err, imf_stream_t stream);

/* Dense GEMM on fp16 hi/lo operands (csrc/h2_gemm.cu): C = alpha * A . W^T (+ bias) (+ R).  A: h2 matrix [M_max, K] (chunk width 64, lda
 * halves; min(*m_dev, M_max) rows are computed); Wpacked = imf_sparse_conv_h2_pack(W^T as [1, K, N], 1, K, N, 64, wmul), 1 / wmul folded
 * into alpha by the caller.  K % 64 == 0, N % 128 == 0.  mode 0: fp32 C (ldc floats) + optional fp32 residual R; mode 1: h2 C (ldc halves,
 * chunk width 64); mode 2: GEGLU h2 C [M, N / 2] (value / gate columns interleaved per 128-column tile by the packing; bias = the
 * un-permuted [N] vector).  err (optional): watchdog codes, bit 16 = an h2 output left the fp16 range. */
int imf_h2_gemm(const void* A, int32_t lda, int32_t M_max, const int32_t* m_dev, const void* Wpacked, int32_t N, int32_t K, float alpha,
                const float* bias, const float* R, int32_t ldr, int32_t mode, void* C, int32_t ldc, int32_t* err, imf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* IMFNET_B200_H_ */
