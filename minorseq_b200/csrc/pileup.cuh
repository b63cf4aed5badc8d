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
// bit-sliced carry-save adders (Harley-Seal) into 11-plane vertical counters held
// in registers.  Codons: a per-column "pivot" base (majority of a read sample)
// turns "read carries the pivot codon" into three shifted ANDs; only the rare
// clean non-pivot codons go to the 64-bin histogram, with a global RED.
// Read tiles are staged into shared memory by cp.async.bulk (UBLKCP) behind an
// mbarrier full/empty ring driven by one producer lane.  At the end the row-groups
// of a CTA add their vertical counters bit-sliced through shared memory, one group
// bit-transposes them into per-column integers and stores the CTA's slice; a small
// finalize kernel sums the <=148 slices into the count tensor.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ms {

constexpr int kPlanes = 11;             // vertical counter planes held in registers
constexpr int kPlanesAll = 13;          // + two planes per mask in shared memory (HI instantiations)
constexpr int kMaxReadsPerFlush = 2047;     // reads of a row-group between flushes: register planes only ...
constexpr int kMaxReadsPerFlushHi = 8191;   // ... and with the two shared-memory planes
// shared memory in front of the chunk slots: mbarriers + release counters (2 KB); HI instantiations add the two upper
// counter planes (2 planes x up to 9 masks x 384 threads).  The plain kernel keeps the small header: the shared memory it
// does not claim stays L1, and K1 lost ~5 % at 1M x 3 kb with 27 KB less of it.
constexpr int kPileupHiOffset = 2048;
constexpr int kPileupSmemHeader = 2048;
constexpr int kPileupSmemHeaderHi = 2048 + 2 * 9 * 384 * 4;
constexpr int kPileupMaxThreads = 384;  // 12 warps = 3 per SM sub-partition -> 168 registers per thread

// which one-bit masks are counted
enum PileupMode { kModeJuliet = 0,  // 6 state sums + "not the pivot codon"
                  kModeFuse = 1,    // 6 state sums + insertion flag
                  kModeBoth = 2 };  // all 8

struct PileupArgs {
    const uint32_t* packed;   // R rows of row_words u32
    int64_t R;
    int32_t L;
    int32_t nblk;             // ceil(L/32)
    int32_t warps_per_group;  // W = ceil(nblk/32)
    int32_t groups;           // G row-groups per CTA
    int32_t nseg, seg_len;    // column segments: CTA c handles blocks [seg_len*(c%nseg), +seg_len) of its reads
    int32_t stages, stages_hi;   // ring depth of the plain and of the HI instantiation (smaller header / larger header)
    int32_t stage_bytes;      // G*8*row_bytes
    const uint2* pivot;       // [nblk+1] {r0, r1} planes of the pivot base
    const uint2* pivot2;      // [nblk+1] DENSE: planes of the designated base (runner-up of the sample where frequent, else the pivot)
    const uint32_t* start_mask; // [nblk] codon start columns
    uint32_t* codon;          // [L][64] global histogram (exceptions land here)
    uint32_t* part_col;       // [gridDim][nblk*32][8]
    uint32_t* part_piv;       // [gridDim][nblk*32]
    uint32_t* part_piv2;      // [gridDim][nblk*32] DENSE: designated-codon counts
    uint4* exc_list;          // [gridDim*blockDim][exc_cap] blocks of the reads logged for the exact codon pass (ExcLog, pileup.cu)
    uint32_t* exc_cnt;        // [gridDim*blockDim]
    uint32_t exc_cap;
    int64_t exc_lists;        // gridDim*blockDim of the pileup launch
};

template <int MODE, bool DENSE, bool SEG, bool HI>
__global__ void pileup_csa_kernel(PileupArgs a);

__global__ void pivot_sample_kernel(const uint32_t* packed, int64_t R, int32_t nblk, int32_t L,
                                    uint2* pivot, uint8_t* pivot_state, uint2* pivot2, uint8_t* pivot2_state, uint32_t* dense_stat);
__global__ void pileup_finalize_kernel(const uint32_t* part_col, const uint32_t* part_piv, const uint32_t* part_piv2, int32_t slices,
                                       int32_t nblk, int32_t L, const uint8_t* pivot_state, const uint8_t* pivot2_state,
                                       const uint32_t* start_mask, uint32_t* col, uint32_t* codon,
                                       int32_t count_codons, int32_t nseg, int32_t seg_len);
__global__ void pileup_atomic_kernel(const uint32_t* packed, int64_t R, int32_t L, int32_t nblk,
                                     const uint32_t* start_mask, uint32_t* col, uint32_t* codon,
                                     int32_t count_codons);
__global__ void coverage_kernel(uint32_t* col, int32_t L);

void pileup_set_smem_attr(int max_smem);
void pileup_launch(int mode, bool dense, bool hi, int grid, int threads, int smem, cudaStream_t s, const PileupArgs& a);
void pileup_exceptions_launch(bool dense, int pileup_grid, int pileup_threads, cudaStream_t s, const PileupArgs& a);

}  // namespace ms
