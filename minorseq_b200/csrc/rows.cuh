// rows.cuh -- the device layout of the packed reads ("tiles"), shared by every kernel that reads or writes them.
//
// A read is ceil(L/32) blocks of four u32 bit-planes (16 bytes, include/minorseq_b200.h).  In device memory the
// reads are grouped in tiles of 8 consecutive reads, stored block-major: the 8 copies of block b -- one per read of
// the tile -- are the 128 contiguous bytes at uint4 index (tile * nblk + b) * 8, read i of the tile sitting at
// position i ^ (b & 7) among them.  Why:
//   * the phasing kernels (K3) gather a few blocks per read: one 128-byte line now serves 8 reads, all of it useful
//     (row-major rows paid a 64-byte DRAM burst for 16 useful bytes);
//   * the pile-up kernel (K1) still stages a whole tile with ONE bulk copy, and a column segment of a tile is
//     contiguous too (one copy instead of eight), which is what makes column segments efficient for long rows;
//   * the XOR keeps K1's shared-memory reads conflict free: the 8 lanes of a quarter warp own consecutive blocks and
//     read the same read i, i.e. positions i ^ (b & 7) -- eight different 16-byte columns of the 128-byte rows.
// A buffer holds whole tiles: ceil(R/8) * nblk * 128 bytes; the slots of reads >= R in the last tile are never
// interpreted.  Host code sees plain rows (ms_expand_cigar, ms_pack_states); ms_tile_rows / ms_tile_rows_dev and the
// event expansion produce tiles.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace ms {

constexpr int kTileReads = 8;

// uint4 index of block b of read r
__host__ __device__ inline size_t tile_slot(int64_t r, int32_t b, int32_t nblk) {
    return ((static_cast<size_t>(r >> 3) * static_cast<size_t>(nblk) + static_cast<size_t>(b)) << 3) +
           static_cast<size_t>((r & 7) ^ (b & 7));
}

inline int64_t tiles_of(int64_t R) { return (R + kTileReads - 1) / kTileReads; }

}  // namespace ms
