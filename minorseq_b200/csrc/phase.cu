// phase.cu -- K3: read-level haplotype phasing.
//
// Replaces juliet --mode-phasing (/root/reference/doc/JULIET.md:192-211): for every
// read, which called variants it carries; reads with identical patterns form a
// haplotype; a read with a deletion (:278-288), a QV-filtered 'N' (:256-259) or
// missing coverage in any variant codon is unsuitable and only tallied in the three
// marginals of the tooltip (:372-381, screenshot juliet_haplotype-tooltip.png).
// SURVEY.md rows a11-a13.
//
// Kernels: phase_bits (one warp per read: stage the touched 32-column blocks in shared
// memory, one lane per variant, ballot -> bit-vector word), phase_insert (64-bit hash of
// the bit-vector into an open-addressing table with atomicCAS), phase_verify (exact compare
// against the slot's representative; a 64-bit collision triggers a re-hash with a new seed),
// compaction, and the popcount-AND co-occurrence over the transposed bit matrix.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>
#include "handle.h"

namespace ms {

struct VarDev {
    int32_t slotA, slotB;  // index into the block list for column col and col+2
    int32_t shift;         // col & 31
    int32_t codon;         // 0..63, or -1: variant lies outside the reference (always partial)
};

__device__ __forceinline__ uint64_t mix64d(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}

constexpr int kPhaseWarps = 8;

__global__ void __launch_bounds__(kPhaseWarps * 32) phase_bits_kernel(
    const uint4* __restrict__ packed, int64_t R, int32_t nblk, const int32_t* __restrict__ blocklist, int32_t NB,
    const VarDev* __restrict__ vars, int32_t V, int32_t vwords, uint32_t* __restrict__ bits, uint8_t* __restrict__ flags,
    unsigned long long* __restrict__ ctr) {
    extern __shared__ uint4 sm[];  // [kPhaseWarps][NB]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4* my = sm + static_cast<size_t>(warp) * NB;
    unsigned long long c_dam = 0, c_gap = 0, c_het = 0, c_par = 0;
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * kPhaseWarps + warp; r < R;
         r += static_cast<int64_t>(gridDim.x) * kPhaseWarps) {
        const uint4* row = packed + static_cast<size_t>(r) * nblk;
        for (int i = lane; i < NB; i += 32) my[i] = row[blocklist[i]];
        __syncwarp();
        uint32_t f = 0;
        for (int w = 0; w < vwords; ++w) {
            const int v = w * 32 + lane;
            bool bit = false;
            if (v < V) {
                const VarDev vd = vars[v];
                if (vd.codon < 0) {
                    f |= MS_FLAG_PARTIAL;
                } else {
                    const uint4 a = my[vd.slotA], b = my[vd.slotB];
                    const uint32_t b0 = __funnelshift_r(a.x, b.x, vd.shift) & 7u;
                    const uint32_t b1 = __funnelshift_r(a.y, b.y, vd.shift) & 7u;
                    const uint32_t z = __funnelshift_r(a.z, b.z, vd.shift) & 7u;
                    if (z & ~b0 & ~b1) f |= MS_FLAG_GAP;      // 100
                    if (z & b0 & ~b1) f |= MS_FLAG_HET;       // 101
                    if (z & b1) f |= MS_FLAG_PARTIAL;         // 11x
                    const uint32_t cod = ((b0 & 1u) << 4) | ((b1 & 1u) << 5) | ((b0 & 2u) << 1) | ((b1 & 2u) << 2) |
                                         ((b0 & 4u) >> 2) | ((b1 & 4u) >> 1);
                    bit = (z == 0u) && (cod == static_cast<uint32_t>(vd.codon));
                }
            }
            const uint32_t word = __ballot_sync(0xffffffffu, bit);
            if (lane == 0) bits[static_cast<size_t>(r) * vwords + w] = word;
        }
        f = __reduce_or_sync(0xffffffffu, f);
        if (lane == 0) {
            flags[r] = static_cast<uint8_t>(f);
            if (f) {
                ++c_dam;
                if (f & MS_FLAG_GAP) ++c_gap;
                if (f & MS_FLAG_HET) ++c_het;
                if (f & MS_FLAG_PARTIAL) ++c_par;
            }
        }
        __syncwarp();
    }
    if (lane == 0 && c_dam) {
        atomicAdd(ctr + 0, c_dam);
        atomicAdd(ctr + 1, c_gap);
        atomicAdd(ctr + 2, c_het);
        atomicAdd(ctr + 3, c_par);
    }
}

__device__ __forceinline__ uint64_t pattern_hash(const uint32_t* w, int32_t vwords, uint64_t seed) {
    uint64_t hsh = seed;
    for (int32_t i = 0; i < vwords; ++i) hsh = mix64d(hsh ^ (static_cast<uint64_t>(w[i]) + 0x9E3779B97F4A7C15ULL * (i + 1)));
    return hsh ? hsh : 1ULL;
}

__global__ void phase_insert_kernel(const uint32_t* __restrict__ bits, const uint8_t* __restrict__ flags, int64_t R,
                                    int32_t vwords, uint64_t seed, unsigned long long* tab_key, uint32_t* tab_cnt,
                                    long long* tab_rep, int64_t mask, int32_t* __restrict__ slot) {
    const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= R) return;
    if (flags[r]) { slot[r] = -1; return; }
    const uint64_t key = pattern_hash(bits + static_cast<size_t>(r) * vwords, vwords, seed);
    int64_t idx = static_cast<int64_t>(key) & mask;
    for (;;) {
        const unsigned long long prev = atomicCAS(tab_key + idx, 0ULL, static_cast<unsigned long long>(key));
        if (prev == 0ULL || prev == key) break;
        idx = (idx + 1) & mask;
    }
    atomicAdd(tab_cnt + idx, 1u);
    atomicMin(tab_rep + idx, static_cast<long long>(r));
    slot[r] = static_cast<int32_t>(idx);
}

__global__ void phase_verify_kernel(const uint32_t* __restrict__ bits, int64_t R, int32_t vwords,
                                    const long long* __restrict__ tab_rep, const int32_t* __restrict__ slot,
                                    unsigned long long* collision) {
    const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= R || slot[r] < 0) return;
    const long long rep = tab_rep[slot[r]];
    if (rep == r) return;
    const uint32_t* a = bits + static_cast<size_t>(r) * vwords;
    const uint32_t* b = bits + static_cast<size_t>(rep) * vwords;
    for (int32_t i = 0; i < vwords; ++i)
        if (a[i] != b[i]) { atomicAdd(collision, 1ULL); return; }
}

__global__ void phase_compact_kernel(const uint32_t* __restrict__ tab_cnt, const long long* __restrict__ tab_rep,
                                     int64_t tab_size, unsigned long long* ngroups, int32_t* __restrict__ g_slot,
                                     uint32_t* __restrict__ g_cnt, long long* __restrict__ g_rep, int64_t cap) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= tab_size || tab_cnt[i] == 0) return;
    const unsigned long long k = atomicAdd(ngroups, 1ULL);
    if (static_cast<int64_t>(k) < cap) {
        g_slot[k] = static_cast<int32_t>(i); g_cnt[k] = tab_cnt[i]; g_rep[k] = tab_rep[i];
    }
}

__global__ void phase_gather_kernel(const uint32_t* __restrict__ bits, const long long* __restrict__ g_rep, int64_t H,
                                    int32_t vwords, uint32_t* __restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= H * vwords) return;
    const int64_t g = i / vwords;
    out[i] = bits[static_cast<size_t>(g_rep[g]) * vwords + (i - g * vwords)];
}

// rank lookup for ms_phase_assign: ordered pattern -> its slot in this rank's table
__global__ void phase_rank_kernel(const uint32_t* __restrict__ ordered, int64_t H, int32_t vwords, uint64_t seed,
                                  const unsigned long long* __restrict__ tab_key, const long long* __restrict__ tab_rep,
                                  int64_t mask, const uint32_t* __restrict__ bits, int32_t* __restrict__ slot_rank) {
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= H) return;
    const uint32_t* pat = ordered + static_cast<size_t>(g) * vwords;
    const uint64_t key = pattern_hash(pat, vwords, seed);
    int64_t idx = static_cast<int64_t>(key) & mask;
    for (;;) {
        const unsigned long long k = tab_key[idx];
        if (k == 0ULL) return;  // pattern not present on this rank
        if (k == key) {
            const uint32_t* b = bits + static_cast<size_t>(tab_rep[idx]) * vwords;
            bool same = true;
            for (int32_t i = 0; i < vwords; ++i) same = same && (b[i] == pat[i]);
            if (same) { slot_rank[idx] = static_cast<int32_t>(g); return; }
        }
        idx = (idx + 1) & mask;
    }
}

__global__ void phase_assign_kernel(const int32_t* __restrict__ slot, const int32_t* __restrict__ slot_rank, int64_t R,
                                    int32_t* __restrict__ hap) {
    const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= R) return;
    hap[r] = slot[r] < 0 ? -1 : slot_rank[slot[r]];
}

// ---- co-occurrence: C = B^T B by popcount-AND over the transposed bit matrix ----
// transpose 32 reads x 32 variants per warp with ballots
__global__ void bits_transpose_kernel(const uint32_t* __restrict__ bits, int64_t R, int32_t vwords, int64_t rwords,
                                      uint32_t* __restrict__ bt) {
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t t = warp / vwords;   // read word
    const int32_t w = static_cast<int32_t>(warp - t * vwords);
    if (t >= rwords) return;
    const int64_t r = t * 32 + lane;
    const uint32_t x = r < R ? bits[static_cast<size_t>(r) * vwords + w] : 0u;
    uint32_t mine = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const uint32_t b = __ballot_sync(0xffffffffu, (x >> i) & 1u);
        if (lane == i) mine = b;
    }
    bt[static_cast<size_t>(w * 32 + lane) * rwords + t] = mine;
}

constexpr int kCoTile = 32, kCoChunk = 32;
__global__ void __launch_bounds__(256) cooccurrence_kernel(const uint32_t* __restrict__ bt, int32_t V, int64_t rwords,
                                                           int32_t* __restrict__ C) {
    __shared__ uint32_t A[kCoTile][kCoChunk + 1], B[kCoTile][kCoChunk + 1];
    const int tv = blockIdx.y * kCoTile, tw = blockIdx.x * kCoTile;
    if (tw < tv) return;  // symmetric: compute the upper triangle, mirror on write
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // ty 0..7
    int32_t acc[4] = {0, 0, 0, 0};
    for (int64_t t0 = 0; t0 < rwords; t0 += kCoChunk) {
        for (int i = ty; i < kCoTile; i += 8) {
            const int64_t t = t0 + tx;
            A[i][tx] = (tv + i < V && t < rwords) ? bt[static_cast<size_t>(tv + i) * rwords + t] : 0u;
            B[i][tx] = (tw + i < V && t < rwords) ? bt[static_cast<size_t>(tw + i) * rwords + t] : 0u;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < kCoChunk; ++k) {
            const uint32_t b = B[tx][k];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] += __popc(A[ty + 8 * q][k] & b);
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int v = tv + ty + 8 * q, w = tw + tx;
        if (v < V && w < V) {
            C[static_cast<size_t>(v) * V + w] = acc[q];
            C[static_cast<size_t>(w) * V + v] = acc[q];
        }
    }
}

static inline bool pattern_less(const uint32_t* a, const uint32_t* b, int32_t nw) {
    for (int32_t i = 0; i < nw; ++i)
        if (a[i] != b[i]) return a[i] < b[i];
    return false;
}

}  // namespace ms

extern "C" {

void ms_phase_free_internal(ms_handle* h) {
    cudaFree(h->d_var); cudaFree(h->d_blocklist); cudaFree(h->d_bits); cudaFree(h->d_flags); cudaFree(h->d_hash);
    cudaFree(h->d_slot); cudaFree(h->d_tab_key); cudaFree(h->d_tab_cnt); cudaFree(h->d_tab_rep); cudaFree(h->d_ctr);
    cudaFree(h->d_cooc); cudaFree(h->d_bits_t);
    h->d_var = nullptr; h->d_blocklist = nullptr; h->d_bits = nullptr; h->d_flags = nullptr; h->d_hash = nullptr;
    h->d_slot = nullptr; h->d_tab_key = nullptr; h->d_tab_cnt = nullptr; h->d_tab_rep = nullptr; h->d_ctr = nullptr;
    h->d_cooc = nullptr; h->d_bits_t = nullptr;
    h->phase_cap = h->phase_n = 0; h->V = 0; h->vwords = 0;
}

int ms_phase_begin(ms_handle* h, const int32_t* var_col, const int32_t* var_codon, int32_t V, int64_t max_reads) {
    if (!h || h->L <= 0 || V < 0 || max_reads < 0 || (V > 0 && (!var_col || !var_codon))) return MS_ERR_ARG;
    MS_CUDA(h, cudaSetDevice(h->device));
    MS_CUDA(h, cudaStreamSynchronize(h->stream));
    ms_phase_free_internal(h);
    h->V = V;
    h->vwords = std::max(1, (V + 31) / 32);
    h->phase_cap = std::max<int64_t>(1, max_reads);
    // distinct 32-column blocks the variants touch
    std::vector<int32_t> blocks;
    for (int32_t v = 0; v < V; ++v) {
        if (var_col[v] < 0 || var_codon[v] < 0 || var_codon[v] > 63) MS_FAIL(h, MS_ERR_ARG, "bad variant");
        if (var_col[v] + 2 < h->L) { blocks.push_back(var_col[v] >> 5); blocks.push_back((var_col[v] + 2) >> 5); }
    }
    std::sort(blocks.begin(), blocks.end());
    blocks.erase(std::unique(blocks.begin(), blocks.end()), blocks.end());
    if (blocks.empty()) blocks.push_back(0);
    h->nblocklist = static_cast<int32_t>(blocks.size());
    std::vector<ms::VarDev> vd(std::max(1, V));
    for (int32_t v = 0; v < V; ++v) {
        ms::VarDev d;
        if (var_col[v] + 2 < h->L) {
            d.slotA = static_cast<int32_t>(std::lower_bound(blocks.begin(), blocks.end(), var_col[v] >> 5) - blocks.begin());
            d.slotB = static_cast<int32_t>(std::lower_bound(blocks.begin(), blocks.end(), (var_col[v] + 2) >> 5) - blocks.begin());
            d.shift = var_col[v] & 31; d.codon = var_codon[v];
        } else { d.slotA = d.slotB = 0; d.shift = 0; d.codon = -1; }
        vd[v] = d;
    }
    MS_CUDA(h, cudaMalloc(&h->d_var, vd.size() * sizeof(ms::VarDev)));
    MS_CUDA(h, cudaMalloc(&h->d_blocklist, blocks.size() * 4));
    MS_CUDA(h, cudaMemcpyAsync(h->d_var, vd.data(), vd.size() * sizeof(ms::VarDev), cudaMemcpyHostToDevice, h->stream));
    MS_CUDA(h, cudaMemcpyAsync(h->d_blocklist, blocks.data(), blocks.size() * 4, cudaMemcpyHostToDevice, h->stream));
    MS_CUDA(h, cudaMalloc(&h->d_bits, static_cast<size_t>(h->phase_cap) * h->vwords * 4));
    MS_CUDA(h, cudaMalloc(&h->d_flags, static_cast<size_t>(h->phase_cap)));
    MS_CUDA(h, cudaMalloc(&h->d_slot, static_cast<size_t>(h->phase_cap) * 4));
    int64_t ts = 1024;
    while (ts < 2 * h->phase_cap) ts <<= 1;
    h->tab_size = ts;
    MS_CUDA(h, cudaMalloc(&h->d_tab_key, static_cast<size_t>(ts) * 8));
    MS_CUDA(h, cudaMalloc(&h->d_tab_cnt, static_cast<size_t>(ts) * 4));
    MS_CUDA(h, cudaMalloc(&h->d_tab_rep, static_cast<size_t>(ts) * 8));
    MS_CUDA(h, cudaMalloc(&h->d_ctr, 8 * 8));
    MS_CUDA(h, cudaMemsetAsync(h->d_ctr, 0, 8 * 8, h->stream));
    MS_CUDA(h, cudaStreamSynchronize(h->stream));
    return MS_OK;
}

int ms_phase_dev(ms_handle* h, const uint32_t* d_packed, int64_t R) {
    if (!h || !h->d_bits || R < 0 || (R > 0 && !d_packed)) return MS_ERR_ARG;
    if (h->phase_n + R > h->phase_cap) MS_FAIL(h, MS_ERR_CAPACITY, "more reads than ms_phase_begin(max_reads)");
    if (R == 0) return MS_OK;
    MS_CUDA(h, cudaSetDevice(h->device));
    const size_t smem = static_cast<size_t>(ms::kPhaseWarps) * h->nblocklist * sizeof(uint4);
    if (smem > 48 * 1024)
        MS_CUDA(h, cudaFuncSetAttribute(ms::phase_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const int64_t want = (R + ms::kPhaseWarps - 1) / ms::kPhaseWarps;
    const int grid = static_cast<int>(std::min<int64_t>(want, static_cast<int64_t>(h->num_sms) * 8));
    ms::phase_bits_kernel<<<grid, ms::kPhaseWarps * 32, smem, h->stream>>>(
        reinterpret_cast<const uint4*>(d_packed), R, h->nblk, h->d_blocklist, h->nblocklist,
        reinterpret_cast<const ms::VarDev*>(h->d_var), h->V, h->vwords,
        h->d_bits + static_cast<size_t>(h->phase_n) * h->vwords, h->d_flags + h->phase_n,
        reinterpret_cast<unsigned long long*>(h->d_ctr));
    h->launches++;
    MS_CUDA(h, cudaGetLastError());
    h->phase_n += R;
    return MS_OK;
}

static uint64_t phase_seed(int attempt) { return 0x6d696e6f72736571ULL + 0x9E3779B97F4A7C15ULL * static_cast<uint64_t>(attempt); }

// builds the table (re-hashing on a 64-bit collision); leaves h->d_tab_* valid
static int phase_build_table(ms_handle* h, int* attempt_out) {
    const int64_t R = h->phase_n;
    unsigned long long* d_coll = reinterpret_cast<unsigned long long*>(h->d_ctr) + 4;
    for (int attempt = 0; attempt < 4; ++attempt) {
        MS_CUDA(h, cudaMemsetAsync(h->d_tab_key, 0, static_cast<size_t>(h->tab_size) * 8, h->stream));
        MS_CUDA(h, cudaMemsetAsync(h->d_tab_cnt, 0, static_cast<size_t>(h->tab_size) * 4, h->stream));
        MS_CUDA(h, cudaMemsetAsync(h->d_tab_rep, 0x7f, static_cast<size_t>(h->tab_size) * 8, h->stream));
        MS_CUDA(h, cudaMemsetAsync(d_coll, 0, 8, h->stream));
        if (R > 0) {
            const int grid = static_cast<int>((R + 255) / 256);
            ms::phase_insert_kernel<<<grid, 256, 0, h->stream>>>(h->d_bits, h->d_flags, R, h->vwords, phase_seed(attempt),
                                                                 reinterpret_cast<unsigned long long*>(h->d_tab_key), h->d_tab_cnt,
                                                                 reinterpret_cast<long long*>(h->d_tab_rep), h->tab_size - 1, h->d_slot);
            ms::phase_verify_kernel<<<grid, 256, 0, h->stream>>>(h->d_bits, R, h->vwords, reinterpret_cast<long long*>(h->d_tab_rep),
                                                                 h->d_slot, d_coll);
            h->launches += 2;
        }
        unsigned long long coll = 0;
        MS_CUDA(h, cudaMemcpyAsync(&coll, d_coll, 8, cudaMemcpyDeviceToHost, h->stream));
        MS_CUDA(h, cudaStreamSynchronize(h->stream));
        if (coll == 0) { *attempt_out = attempt; return MS_OK; }
    }
    MS_FAIL(h, MS_ERR_CUDA, "haplotype hash collided under four seeds");
}

int ms_phase_groups(ms_handle* h, uint32_t* patterns, uint64_t* counts, int64_t cap, int64_t* H, ms_phase_counters* ctr) {
    if (!h || !h->d_bits || !H) return MS_ERR_ARG;
    MS_CUDA(h, cudaSetDevice(h->device));
    int attempt = 0;
    int rc = phase_build_table(h, &attempt);
    if (rc != MS_OK) return rc;
    // compact
    const int64_t maxg = std::max<int64_t>(1, h->phase_n);
    int32_t* g_slot = nullptr; uint32_t* g_cnt = nullptr; long long* g_rep = nullptr; uint32_t* g_pat = nullptr;
    unsigned long long* d_ng = reinterpret_cast<unsigned long long*>(h->d_ctr) + 5;
    MS_CUDA(h, cudaMalloc(&g_slot, maxg * 4)); MS_CUDA(h, cudaMalloc(&g_cnt, maxg * 4)); MS_CUDA(h, cudaMalloc(&g_rep, maxg * 8));
    MS_CUDA(h, cudaMemsetAsync(d_ng, 0, 8, h->stream));
    ms::phase_compact_kernel<<<static_cast<int>((h->tab_size + 255) / 256), 256, 0, h->stream>>>(
        h->d_tab_cnt, reinterpret_cast<long long*>(h->d_tab_rep), h->tab_size, d_ng, g_slot, g_cnt, g_rep, maxg);
    h->launches++;
    unsigned long long ng = 0;
    uint64_t hc[4];
    MS_CUDA(h, cudaMemcpyAsync(&ng, d_ng, 8, cudaMemcpyDeviceToHost, h->stream));
    MS_CUDA(h, cudaMemcpyAsync(hc, h->d_ctr, 32, cudaMemcpyDeviceToHost, h->stream));
    MS_CUDA(h, cudaStreamSynchronize(h->stream));
    const int32_t nw = h->vwords;
    std::vector<uint32_t> pat(static_cast<size_t>(ng) * nw), cnt(ng);
    if (ng) {
        MS_CUDA(h, cudaMalloc(&g_pat, static_cast<size_t>(ng) * nw * 4));
        const int64_t tot = static_cast<int64_t>(ng) * nw;
        ms::phase_gather_kernel<<<static_cast<int>((tot + 255) / 256), 256, 0, h->stream>>>(h->d_bits, g_rep, static_cast<int64_t>(ng), nw, g_pat);
        h->launches++;
        MS_CUDA(h, cudaMemcpyAsync(pat.data(), g_pat, pat.size() * 4, cudaMemcpyDeviceToHost, h->stream));
        MS_CUDA(h, cudaMemcpyAsync(cnt.data(), g_cnt, cnt.size() * 4, cudaMemcpyDeviceToHost, h->stream));
        MS_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    cudaFree(g_slot); cudaFree(g_cnt); cudaFree(g_rep); cudaFree(g_pat);
    std::vector<int64_t> order(ng);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        return ms::pattern_less(pat.data() + static_cast<size_t>(a) * nw, pat.data() + static_cast<size_t>(b) * nw, nw);
    });
    for (int64_t i = 0; i < std::min<int64_t>(cap, static_cast<int64_t>(ng)); ++i) {
        if (patterns) memcpy(patterns + static_cast<size_t>(i) * nw, pat.data() + static_cast<size_t>(order[i]) * nw, static_cast<size_t>(nw) * 4);
        if (counts) counts[i] = cnt[order[i]];
    }
    *H = static_cast<int64_t>(ng);
    if (ctr) {
        ctr->reported = 0; ctr->insufficient = 0;
        ctr->damaged = hc[0]; ctr->gaps = hc[1]; ctr->heteroduplex = hc[2]; ctr->partial = hc[3];
    }
    return MS_OK;
}

int ms_haplotype_order(uint32_t* patterns, uint64_t* counts, int64_t H, int32_t V, int32_t min_reads, int64_t* Hmerged,
                       int64_t* nreported, ms_phase_counters* ctr) {
    if (H < 0 || (H > 0 && (!patterns || !counts)) || V < 0) return MS_ERR_ARG;
    const int32_t nw = std::max(1, (V + 31) / 32);
    std::vector<int64_t> order(H);
    std::iota(order.begin(), order.end(), 0);
    auto pat = [&](int64_t i) { return patterns + static_cast<size_t>(i) * nw; };
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return ms::pattern_less(pat(a), pat(b), nw); });
    // merge equal patterns (same haplotype seen on several ranks)
    std::vector<uint32_t> mp; std::vector<uint64_t> mc;
    for (int64_t i = 0; i < H; ++i) {
        const uint32_t* p = pat(order[i]);
        if (!mc.empty() && memcmp(mp.data() + (mc.size() - 1) * nw, p, static_cast<size_t>(nw) * 4) == 0) mc.back() += counts[order[i]];
        else { mp.insert(mp.end(), p, p + nw); mc.push_back(counts[order[i]]); }
    }
    const int64_t M = static_cast<int64_t>(mc.size());
    std::vector<int64_t> ord2(M);
    std::iota(ord2.begin(), ord2.end(), 0);
    std::stable_sort(ord2.begin(), ord2.end(), [&](int64_t a, int64_t b) { return mc[a] > mc[b]; });  // ties keep ascending pattern
    int64_t nrep = 0; uint64_t rep = 0, ins = 0;
    for (int64_t i = 0; i < M; ++i) {
        memcpy(patterns + static_cast<size_t>(i) * nw, mp.data() + static_cast<size_t>(ord2[i]) * nw, static_cast<size_t>(nw) * 4);
        counts[i] = mc[ord2[i]];
        if (counts[i] >= static_cast<uint64_t>(min_reads)) { ++nrep; rep += counts[i]; } else ins += counts[i];
    }
    if (Hmerged) *Hmerged = M;
    if (nreported) *nreported = nrep;
    if (ctr) { ctr->reported = rep; ctr->insufficient = ins; }
    return MS_OK;
}

void ms_haplotype_name(int64_t rank, char buf[3]) {
    if (rank < 26) { buf[0] = static_cast<char>('A' + rank); buf[1] = 0; return; }
    rank -= 26;
    buf[0] = static_cast<char>('A' + (rank / 26) % 26); buf[1] = static_cast<char>('a' + rank % 26); buf[2] = 0;
}

int ms_phase_assign(ms_handle* h, const uint32_t* ordered_patterns, int64_t H, int32_t* hap_id) {
    if (!h || !h->d_bits || H < 0 || (H > 0 && !ordered_patterns) || !hap_id) return MS_ERR_ARG;
    MS_CUDA(h, cudaSetDevice(h->device));
    int attempt = 0;
    int rc = phase_build_table(h, &attempt);
    if (rc != MS_OK) return rc;
    const int64_t R = h->phase_n;
    const int32_t nw = h->vwords;
    int32_t *d_rank = nullptr, *d_hap = nullptr; uint32_t* d_pat = nullptr;
    MS_CUDA(h, cudaMalloc(&d_rank, static_cast<size_t>(h->tab_size) * 4));
    MS_CUDA(h, cudaMalloc(&d_hap, static_cast<size_t>(std::max<int64_t>(1, R)) * 4));
    MS_CUDA(h, cudaMalloc(&d_pat, static_cast<size_t>(std::max<int64_t>(1, H)) * nw * 4));
    MS_CUDA(h, cudaMemsetAsync(d_rank, 0xff, static_cast<size_t>(h->tab_size) * 4, h->stream));
    if (H > 0) {
        MS_CUDA(h, cudaMemcpyAsync(d_pat, ordered_patterns, static_cast<size_t>(H) * nw * 4, cudaMemcpyHostToDevice, h->stream));
        ms::phase_rank_kernel<<<static_cast<int>((H + 127) / 128), 128, 0, h->stream>>>(
            d_pat, H, nw, phase_seed(attempt), reinterpret_cast<unsigned long long*>(h->d_tab_key),
            reinterpret_cast<long long*>(h->d_tab_rep), h->tab_size - 1, h->d_bits, d_rank);
        h->launches++;
    }
    if (R > 0) {
        ms::phase_assign_kernel<<<static_cast<int>((R + 255) / 256), 256, 0, h->stream>>>(h->d_slot, d_rank, R, d_hap);
        h->launches++;
        MS_CUDA(h, cudaMemcpyAsync(hap_id, d_hap, static_cast<size_t>(R) * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    MS_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(d_rank); cudaFree(d_hap); cudaFree(d_pat);
    MS_CUDA(h, cudaGetLastError());
    return MS_OK;
}

int ms_phase_device(ms_handle* h, uint32_t** d_bits, uint8_t** d_flags, int64_t* R) {
    if (!h || !h->d_bits) return MS_ERR_ARG;
    if (d_bits) *d_bits = h->d_bits;
    if (d_flags) *d_flags = h->d_flags;
    if (R) *R = h->phase_n;
    return MS_OK;
}

int ms_cooccurrence(ms_handle* h, int32_t** d_C) {
    if (!h || !h->d_bits || !d_C) return MS_ERR_ARG;
    MS_CUDA(h, cudaSetDevice(h->device));
    const int32_t V = h->V, nw = h->vwords;
    const int64_t R = h->phase_n, rwords = std::max<int64_t>(1, (R + 31) / 32);
    cudaFree(h->d_cooc); cudaFree(h->d_bits_t);
    h->d_cooc = nullptr; h->d_bits_t = nullptr;
    MS_CUDA(h, cudaMalloc(&h->d_cooc, std::max<size_t>(4, static_cast<size_t>(V) * V * 4)));
    MS_CUDA(h, cudaMalloc(&h->d_bits_t, static_cast<size_t>(nw) * 32 * rwords * 4));
    MS_CUDA(h, cudaMemsetAsync(h->d_cooc, 0, std::max<size_t>(4, static_cast<size_t>(V) * V * 4), h->stream));
    if (V > 0) {
        const int64_t nwarps = rwords * nw;
        ms::bits_transpose_kernel<<<static_cast<unsigned>((nwarps * 32 + 255) / 256), 256, 0, h->stream>>>(h->d_bits, R, nw, rwords, h->d_bits_t);
        dim3 grid((V + ms::kCoTile - 1) / ms::kCoTile, (V + ms::kCoTile - 1) / ms::kCoTile);
        ms::cooccurrence_kernel<<<grid, 256, 0, h->stream>>>(h->d_bits_t, V, rwords, h->d_cooc);
        h->launches += 2;
    }
    MS_CUDA(h, cudaGetLastError());
    *d_C = h->d_cooc;
    return MS_OK;
}

}  // extern "C"
