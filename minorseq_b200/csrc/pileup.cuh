// pileup.cuh -- K1: juliet/fuse pileup over planar 4-bit packed reads (sm_100a).
//
// Replaces the per-column "MSA counts" and per-codon coverage/observed-codon
// counts juliet derives from the alignment (/root/reference/doc/JULIET.md:96-100,
// screenshot juliet_hiv-context.png) and fuse's per-column tallies
// (/root/reference/doc/FUSE.md:17-20).  SURVEY.md rows a4, a5.
//
// Design (DESIGN.md "K1"): at the HBM roofline one SM must absorb ~46 columns per
// clock, far beyond shared-memory atomic throughput, so nothing is counted with
// atomics.  A thread owns one 32-column block and walks down the reads; the four
// bit-planes of each read give 32-wide one-bit masks, which are summed with
// bit-sliced carry-save adders (Harley-Seal) into 12-plane vertical counters held
// in registers.  Codons: a per-column "pivot" base (majority of a read sample)
// turns "read carries the pivot codon" into three shifted ANDs; only the rare
// clean non-pivot codons go to the 64-bin histogram, with a global RED.
// Read tiles are staged into shared memory by cp.async.bulk (UBLKCP) behind an
// mbarrier full/empty ring driven by one producer lane.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ms {

constexpr int kPlanes = 12;            // vertical counter depth: up to 4095 reads between flushes
constexpr int kMaxReadsPerFlush = 4088;
constexpr int kMasks = 8;              // S0 S1 S2 S01 S02 S12 INS NOTPIVOT
constexpr int kPileupMaxThreads = 352; // 10 compute warps + 1 producer warp

struct PileupArgs {
    const uint32_t* packed;   // R rows of row_words u32
    int64_t R;
    int32_t L;
    int32_t nblk;             // ceil(L/32)
    int32_t warps_per_group;  // W = ceil(nblk/32)
    int32_t groups;           // G row-groups per CTA
    int32_t blocks8;          // U: 8-read blocks per group per tile
    int32_t stages;
    int32_t stage_bytes;      // G*U*8*row_bytes (+16 slack handled by host)
    const uint2* pivot;       // [nblk+1] {r0, r1} planes of the pivot base; may be null when !codon
    const uint32_t* start_mask; // [nblk] codon start columns; null when !codon
    uint32_t* codon;          // [L][64] global histogram (exceptions land here)
    uint32_t* part_col;       // [slices][nblk*32][8]
    uint32_t* part_piv;       // [slices][nblk*32]
    int32_t count_codons;
};

__global__ void pileup_csa_kernel(PileupArgs a);
__global__ void pivot_sample_kernel(const uint32_t* packed, int64_t R, int32_t nblk, int32_t L,
                                    uint2* pivot, uint8_t* pivot_state);
__global__ void pileup_finalize_kernel(const uint32_t* part_col, const uint32_t* part_piv, int32_t slices,
                                       int32_t nblk, int32_t L, const uint8_t* pivot_state,
                                       const uint32_t* start_mask, uint32_t* col, uint32_t* codon,
                                       int32_t count_codons);
__global__ void pileup_atomic_kernel(const uint32_t* packed, int64_t R, int32_t L, int32_t nblk,
                                     const uint32_t* start_mask, uint32_t* col, uint32_t* codon,
                                     int32_t count_codons);
__global__ void coverage_kernel(uint32_t* col, int32_t L);

}  // namespace ms
