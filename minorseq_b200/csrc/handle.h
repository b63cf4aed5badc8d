// handle.h -- internal state behind the opaque ms_handle of include/minorseq_b200.h
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/minorseq_b200.h"

struct ms_handle {
    int device = 0;
    int num_sms = 0;
    int max_smem = 0;
    cudaStream_t own_stream = nullptr, copy_stream = nullptr, stream = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    cudaEvent_t ev_k1[2] = {nullptr, nullptr};  // around the last K1 launch when timing is on
    bool timing = false;
    int64_t k1_reads = 0;
    std::string err;
    int64_t launches = 0;

    // layout
    int32_t L = 0, nblk = 0;
    bool count_codons = false;
    int variant = 0;
    uint32_t* d_counts = nullptr;  // [L*8 | L*64]
    uint32_t* d_start = nullptr;   // [nblk]
    uint2* d_pivot = nullptr;      // [nblk+1]
    uint8_t* d_pivot_state = nullptr;  // [nblk*32 + 2]
    uint32_t *d_part_col = nullptr, *d_part_piv = nullptr;
    int32_t groups = 1, wpg = 1, blocks8 = 1, stages = 2, stage_bytes = 0, smem_bytes = 0;
    bool have_pivot = false;
    std::vector<uint32_t> h_start;

    // host-upload staging (ms_pileup_host keeps the rows for phasing)
    uint32_t* d_upload = nullptr;
    size_t upload_cap = 0;

    // call
    void* d_call_buf = nullptr;
    size_t call_cap = 0;

    // phasing
    int32_t V = 0, vwords = 0;
    int64_t phase_cap = 0, phase_n = 0;
    int32_t* d_var = nullptr;       // [V] {col, codon, ...} packed
    int32_t* d_blocklist = nullptr; // distinct blocks touched by the variants
    int32_t nblocklist = 0;
    uint32_t* d_bits = nullptr;
    uint8_t* d_flags = nullptr;
    uint64_t* d_hash = nullptr;     // per read
    int32_t* d_slot = nullptr;      // per read -> table slot
    uint64_t* d_tab_key = nullptr;  // open-addressing table
    uint32_t* d_tab_cnt = nullptr;
    int64_t* d_tab_rep = nullptr;
    int64_t tab_size = 0;
    uint64_t* d_ctr = nullptr;      // damage counters + collision flag
    int32_t* d_cooc = nullptr;
    uint32_t* d_bits_t = nullptr;

    // fuse
    char* d_seq = nullptr;
};

#define MS_CUDA(h, call)                                                                    \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                  \
            return MS_ERR_CUDA;                                                             \
        }                                                                                   \
    } while (0)

#define MS_FAIL(h, code, msg) \
    do {                      \
        (h)->err = (msg);     \
        return (code);        \
    } while (0)
