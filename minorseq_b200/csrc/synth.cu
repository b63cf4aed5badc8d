// synth.cu -- seeded synthetic amplicon reads written directly in the packed format.
//
// Workload generator for bench.py and the full-size tests: mixdata-style
// strain mixing (/root/reference/doc/MIXDATA.md:12-13: first strain is the major, the
// others are minors at given percentages) plus per-base noise.  The noise is a pure
// function of (seed, read, column) -- see minorseq_b200/synth.py for the numpy twin the
// tests compare against bit for bit.
#include "handle.h"
#include "rows.cuh"

namespace ms {

__host__ __device__ inline uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}
constexpr uint64_t kK1 = 0x9E3779B97F4A7C15ULL, kK2 = 0xD1B54A32D192ED03ULL;

__global__ void synth_kernel(ms_synth_params p, const uint8_t* __restrict__ strain_base,
                             const uint32_t* __restrict__ thr_del, const uint32_t* __restrict__ strain_cum,
                             int64_t read0, int64_t R, int32_t nblk, uint4* __restrict__ out) {
    // one thread per 16-byte slot of the tile layout (rows.cuh): consecutive threads write consecutive slots
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= ((R + 7) >> 3) * 8 * nblk) return;
    const int64_t tb = idx >> 3;
    const int64_t tile = tb / nblk;
    const int32_t blk = static_cast<int32_t>(tb - tile * nblk);
    const int64_t rl = tile * 8 + ((idx & 7) ^ (blk & 7));
    if (rl >= R) { out[idx] = make_uint4(~0u, ~0u, ~0u, 0u); return; }   // padding of the last tile: not spanned
    const uint64_t r = static_cast<uint64_t>(read0 + rl);
    const uint64_t y = mix64(p.seed + r * kK1);
    const uint32_t us = static_cast<uint32_t>(y);
    int32_t strain = p.nstrains - 1;
    for (int32_t s = 0; s < p.nstrains; ++s)
        if (us < strain_cum[s]) { strain = s; break; }
    int32_t begin = 0, end = p.L;
    if (((y >> 32) & 0xffffu) < p.thr_trunc16) {
        const uint32_t z = static_cast<uint32_t>(y >> 48);
        const int32_t amount = static_cast<int32_t>((static_cast<uint64_t>(z >> 1) * static_cast<uint64_t>(p.L / 2)) >> 15);
        if (z & 1u) begin = amount; else end = p.L - amount;
    }
    const uint8_t* sb = strain_base + static_cast<size_t>(strain) * p.L;
    uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0;
    for (int32_t j = 0; j < 32; ++j) {
        const int32_t c = blk * 32 + j;
        uint32_t st = 7, ins = 0;
        if (c < p.L && c >= begin && c < end) {
            const uint64_t x = mix64(p.seed + r * kK1 + static_cast<uint64_t>(c + 1) * kK2);
            const uint64_t u = x & 0xffffffffULL;
            const uint64_t tN = p.thr_N, tD = tN + thr_del[c], tS = tD + p.thr_sub;
            const uint32_t base = sb[c];
            if (u < tN) st = 5;
            else if (u < tD) st = 4;
            else if (u < tS) st = (base + 1u + static_cast<uint32_t>((x >> 32) % 3ULL)) & 3u;
            else st = base;
            ins = ((x >> 44) & 0xfffffULL) < p.thr_ins20 ? 1u : 0u;
        }
        p0 |= (st & 1u) << j;
        p1 |= ((st >> 1) & 1u) << j;
        p2 |= ((st >> 2) & 1u) << j;
        p3 |= ins << j;
    }
    out[idx] = make_uint4(p0, p1, p2, p3);
}

}  // namespace ms

extern "C" int ms_synth_dev(ms_handle* h, const ms_synth_params* p, const uint8_t* strain_base, const uint32_t* thr_del,
                            const uint32_t* strain_cum, int64_t read0, int64_t R, uint32_t* d_packed) {
    if (!h || !p || !strain_base || !thr_del || !strain_cum || !d_packed || R < 0 || p->L <= 0 || p->nstrains <= 0)
        return MS_ERR_ARG;
    if (R == 0) return MS_OK;
    MS_CUDA(h, cudaSetDevice(h->device));
    const int32_t nblk = (p->L + 31) / 32;
    uint8_t* d_sb = nullptr;
    uint32_t *d_td = nullptr, *d_sc = nullptr;
    MS_CUDA(h, cudaMalloc(&d_sb, static_cast<size_t>(p->nstrains) * p->L));
    MS_CUDA(h, cudaMalloc(&d_td, static_cast<size_t>(p->L) * 4));
    MS_CUDA(h, cudaMalloc(&d_sc, static_cast<size_t>(p->nstrains) * 4));
    MS_CUDA(h, cudaMemcpyAsync(d_sb, strain_base, static_cast<size_t>(p->nstrains) * p->L, cudaMemcpyHostToDevice, h->stream));
    MS_CUDA(h, cudaMemcpyAsync(d_td, thr_del, static_cast<size_t>(p->L) * 4, cudaMemcpyHostToDevice, h->stream));
    MS_CUDA(h, cudaMemcpyAsync(d_sc, strain_cum, static_cast<size_t>(p->nstrains) * 4, cudaMemcpyHostToDevice, h->stream));
    const int64_t total = ms::tiles_of(R) * 8 * nblk;
    ms::synth_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, h->stream>>>(
        *p, d_sb, d_td, d_sc, read0, R, nblk, reinterpret_cast<uint4*>(d_packed));
    h->launches++;
    MS_CUDA(h, cudaGetLastError());
    MS_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(d_sb); cudaFree(d_td); cudaFree(d_sc);
    return MS_OK;
}
