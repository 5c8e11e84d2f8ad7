// imfnet_b200 -- dense per-item grid over the bounding boxes of a coordinate set (built by conv_first_tc.cu, also read by the
// neighbour-table kernel of coords.cu).  A cell holds (row index + 1) of the voxel at that coordinate, 0 = empty.
#pragma once

namespace imf_dense {

constexpr int kMaxItems = 256;
constexpr int kItemInts = 8;          // per item: x0, y0, z0 (box origin incl. halo), DX, DY, DZ, base (two ints: 64-bit cell offset)

struct CfMeta {                        // head of the workspace
  int use_grid;                        // 1: every item's box fits the grid budget; 0: probe the hash table
  int pad[7];                          // [0], [1]: total cells (64 bit); [2]: number of items
  int bbox[kMaxItems][6];              // running min x,y,z / max x,y,z per item (k_cf_bbox)
  int item[kMaxItems][kItemInts];
};

__device__ __forceinline__ long long cf_cell(const int* it, int x, int y, int z) {
  const long long base = ((long long)(unsigned)it[6]) | ((long long)it[7] << 32);
  return base + ((long long)(z - it[2]) * it[4] + (y - it[1])) * it[3] + (x - it[0]);
}

}  // namespace imf_dense
