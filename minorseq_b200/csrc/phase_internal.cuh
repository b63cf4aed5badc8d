// phase_internal.cuh -- pieces of K3 shared by phase.cu (bit-vectors, local grouping table) and
// phase_order.cu (cross-rank merge + ordering on the device).
#pragma once
#include "handle.h"

namespace ms {

__device__ __forceinline__ uint64_t mix64d(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}

__device__ __forceinline__ uint64_t pattern_hash(const uint32_t* w, int32_t vwords, uint64_t seed) {
    uint64_t hsh = seed;
    for (int32_t i = 0; i < vwords; ++i) hsh = mix64d(hsh ^ (static_cast<uint64_t>(w[i]) + 0x9E3779B97F4A7C15ULL * (i + 1)));
    return hsh ? hsh : 1ULL;
}

// juliet's haplotype order among equal counts: ascending pattern words, word 0 first
__host__ __device__ __forceinline__ bool pattern_less(const uint32_t* a, const uint32_t* b, int32_t nw) {
    for (int32_t i = 0; i < nw; ++i)
        if (a[i] != b[i]) return a[i] < b[i];
    return false;
}

inline uint64_t phase_seed(int attempt) { return 0x6d696e6f72736571ULL + 0x9E3779B97F4A7C15ULL * static_cast<uint64_t>(attempt); }

// ctr layout (u64): [0] damaged [1] gaps [2] heteroduplex [3] partial [4] hash collisions [5] ngroups [6] table overflow [7] spare
inline unsigned long long* phase_ctr(ms_handle* h) { return h->b_ctr.as<unsigned long long>(); }

// Device-built phasing plan (phase_plan_kernel): the pooled variant list and the word stream of phase_bits_kernel made on the
// GPU from K2's output, so that the juliet pass needs no host round trip between the codon test and the bit-vectors.
// Covers the usual case -- at most 32 distinct (column, codon) keys from at most 64 calls; anything else sets `fallback`
// and the host builds the plan (ms_phase_begin) after all.
constexpr int kPlanMaxKeys = 32, kPlanMaxCalls = 64, kPlanMaxBlocks = 64, kPlanMaxLayers = 4;
constexpr int kPlanStreamWords = kPlanMaxBlocks * kPlanMaxLayers * 5 + kPlanMaxKeys;
struct PhasePlan {
    int32_t V, NB, nwords, partial_all, fallback, ncalls;
    int32_t key_col[kPlanMaxKeys], key_codon[kPlanMaxKeys];
};

int phase_ensure_stage(ms_handle* h, size_t bytes);
int phase_build_table(ms_handle* h, int attempt);
// enqueue: distinct patterns of the local table -> g_cnt/g_pat (and the table slot of each in g_slot, may be null); count in ctr[5]
int phase_compact(ms_handle* h, uint32_t* g_cnt, uint32_t* g_pat, int32_t* g_slot, int64_t cap);
// after a host look at ctr[6]: grow the table (x8) for the next build; fails at the maximum size
int phase_grow_table(ms_handle* h);

}  // namespace ms
