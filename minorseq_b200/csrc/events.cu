// events.cu -- compact "event rows": the host->device form of an aligned read, and its expansion on the GPU.
//
// A PacBio CCS alignment is "the reference, except at a few columns" (CIGAR `=` runs with sparse X / D / I ops,
// /root/reference/doc/JULIET.md:49-58, plus the QV-filtered bases that become N, :256-259).  The planar rows K1 and K3
// read cost L/2 bytes per read on the PCIe link (1504 B at 3 kb), which is what bounds the end-to-end pass.  Here the
// host ships, per read, its span and two sorted event lists (one byte per QV-filtered base, 12 bits -- column delta + new
// 4-bit column value -- per other event) against a base sequence both sides hold (~85 events = ~113 B per 3 kb read at
// CCS error rates), and expand_events_kernel rebuilds the packed reads in HBM (as tiles, rows.cuh), where the pile-up and
// the phasing kernels run unchanged.  SURVEY.md rows a2/a3 (host CIGAR walk) and 8f-2 ("GPU-side CIGAR expansion is the
// next real speed-up").
//
// Format (include/minorseq_b200.h):  ms_read_hdr hdr[R+1] = {ev_off, begin, end}; the byte string of read r is
// events[hdr[r].ev_off .. hdr[r+1].ev_off): empty when the read equals the base on its whole span, else
//   [nN: u16] [N list: nN bytes] [rest list: 12-bit entries, entry k in bits [12k, 12k+12) of what follows].
// Both lists walk the columns from the read's `begin`: an entry's column is the previous entry's column + delta.  N list (the
// QV-filtered bases, /root/reference/doc/JULIET.md:256-259 -- two thirds of all events on CCS data): one byte per entry, delta
// 0..254 = the column holds 'N' (nibble 5), 255 = move on by 255 columns.  Rest list: delta << 4 | nibble with delta 0..254 = the
// column holds `nibble` (state | insertion-follows << 3), 255 = move on by 255 columns.  Every spanned column without an
// entry holds the base sequence's base, columns outside [begin, end) are "not spanned"; no column appears twice.  hdr[R] is a
// sentinel: ev_off = total number of event bytes, begin | end << 16 = a 32-bit hash of the base sequence, so that rows
// encoded against another base are rejected instead of silently mis-expanded.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>
#include "handle.h"
#include "rows.cuh"

namespace ms {

static uint32_t base_hash(const uint8_t* base, int32_t L) {   // FNV-1a over (L, base)
    uint32_t hsh = 2166136261u;
    auto eat = [&](uint8_t b) { hsh ^= b; hsh *= 16777619u; };
    for (int k = 0; k < 4; ++k) eat(static_cast<uint8_t>(static_cast<uint32_t>(L) >> (8 * k)));
    for (int32_t i = 0; i < L; ++i) eat(base[i] & 3u);
    return hsh;
}

// base sequence as two bit-planes per 32-column block (same layout as planes P0/P1 of a row)
static void base_planes(const uint8_t* base, int32_t L, std::vector<uint32_t>& pl) {
    const int32_t nblk = (L + 31) / 32;
    pl.assign(static_cast<size_t>(nblk) * 2, 0u);
    for (int32_t c = 0; c < L; ++c) {
        pl[2 * (c >> 5)] |= static_cast<uint32_t>(base[c] & 1u) << (c & 31);
        pl[2 * (c >> 5) + 1] |= static_cast<uint32_t>((base[c] >> 1) & 1u) << (c & 31);
    }
}

static inline uint32_t span_mask(int32_t blk, int32_t begin, int32_t end) {   // columns of block blk inside [begin, end)
    const int32_t lo = std::max(0, begin - 32 * blk), hi = std::min(32, end - 32 * blk);
    if (hi <= lo) return 0u;
    const uint32_t upto_hi = hi == 32 ? 0xffffffffu : ((1u << hi) - 1u);
    return upto_hi & ~((1u << lo) - 1u);
}

static inline uint32_t nibble_at(const uint32_t* row, int32_t c) {
    const uint32_t* w = row + 4 * (c >> 5);
    const int sh = c & 31;
    return ((w[0] >> sh) & 1u) | (((w[1] >> sh) & 1u) << 1) | (((w[2] >> sh) & 1u) << 2) | (((w[3] >> sh) & 1u) << 3);
}

constexpr int32_t kSkip = 255;      // delta value that only moves on (by 255 columns)

// appends 12-bit entries to a read's byte string
struct EventWriter {
    uint8_t* ev;
    int64_t cap, o = 0;   // o: bytes completely or partly written
    bool half = false;    // an odd number of entries so far: the high nibble of ev[o-1] is still free
    bool put(uint32_t e) {
        if (!half) {
            if (o + 2 > cap) return false;
            ev[o] = static_cast<uint8_t>(e & 0xffu);
            ev[o + 1] = static_cast<uint8_t>(e >> 8);          // low nibble; the next entry fills the high one
            o += 2; half = true;
        } else {
            if (o + 1 > cap) return false;
            ev[o - 1] = static_cast<uint8_t>(ev[o - 1] | ((e & 0xfu) << 4));
            ev[o] = static_cast<uint8_t>(e >> 4);
            o += 1; half = false;
        }
        return true;
    }
};

// one planar row -> span + event bytes.  Returns the number of bytes written, or -1 when cap is too small.
static int64_t encode_row(const uint32_t* row, int32_t L, const uint32_t* bpl, uint8_t* ev, int64_t cap, int32_t& begin, int32_t& end) {
    const int32_t nblk = (L + 31) / 32;
    begin = end = 0;
    int32_t first = -1, last = -1;
    for (int32_t b = 0; b < nblk; ++b) {
        const uint32_t spanned = ~(row[4 * b] & row[4 * b + 1] & row[4 * b + 2]) | row[4 * b + 3];   // any column that is not plain "not spanned"
        if (spanned) {
            if (first < 0) first = 32 * b + __builtin_ctz(spanned);
            last = 32 * b + 31 - __builtin_clz(spanned);
        }
    }
    if (first < 0) return 0;
    begin = first; end = last + 1;
    // pass 1: the 'N' columns (nibble exactly 5), one byte each, behind the 2-byte count
    int64_t o = 2;
    int32_t prev = begin;
    bool any = false;
    for (int32_t b = begin >> 5; b <= (end - 1) >> 5; ++b) {
        const uint32_t* q = row + 4 * b;
        uint32_t isN = (q[0] & ~q[1] & q[2] & ~q[3]) & span_mask(b, begin, end);
        while (isN) {
            const int32_t c = 32 * b + __builtin_ctz(isN);
            isN &= isN - 1;
            while (c - prev >= kSkip) {
                if (o >= cap) return -1;
                ev[o++] = static_cast<uint8_t>(kSkip);
                prev += kSkip;
            }
            if (o >= cap) return -1;
            ev[o++] = static_cast<uint8_t>(c - prev);
            prev = c;
            any = true;
        }
    }
    const int64_t nN = o - 2;
    if (nN > 0xffff) return -1;      // cannot happen for L <= 65535: an entry moves on by >= 1 column (the first by >= 0), a skip by 255
    // pass 2: every other column that differs from the base, 12 bits each
    EventWriter w{ev + o, cap - o};
    prev = begin;
    for (int32_t b = begin >> 5; b <= (end - 1) >> 5; ++b) {
        const uint32_t* q = row + 4 * b;
        const uint32_t isN = q[0] & ~q[1] & q[2] & ~q[3];
        uint32_t diff = ((q[0] ^ bpl[2 * b]) | (q[1] ^ bpl[2 * b + 1]) | q[2] | q[3]) & ~isN & span_mask(b, begin, end);
        while (diff) {
            const int32_t c = 32 * b + __builtin_ctz(diff);
            diff &= diff - 1;
            while (c - prev >= kSkip) {
                if (!w.put(static_cast<uint32_t>(kSkip) << 4)) return -1;
                prev += kSkip;
            }
            if (!w.put((static_cast<uint32_t>(c - prev) << 4) | nibble_at(row, c))) return -1;
            prev = c;
            any = true;
        }
    }
    if (!any) return 0;              // the read is the base on its whole span: no bytes at all
    if (cap < 2) return -1;
    ev[0] = static_cast<uint8_t>(nN & 0xff);
    ev[1] = static_cast<uint8_t>(nN >> 8);
    return o + w.o;
}

// ---------------------------------------------------------------- device side
__device__ __forceinline__ uint32_t ev_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

constexpr int kExpandWarps = 8;           // one tile (rows.cuh) per CTA and iteration
constexpr int kStageBytes = 512;          // per warp: the window of a read's event bytes that is staged in shared memory

struct EvHdr { uint32_t off0, off1, span; };     // span = begin | end << 16

__device__ __forceinline__ EvHdr ev_load_hdr(const ms_read_hdr* __restrict__ hdr, int64_t r, int64_t R) {
    EvHdr h{0u, 0u, 0u};
    if (r < R) {
        const uint2 a = *reinterpret_cast<const uint2*>(hdr + r);
        h.off0 = a.x; h.span = a.y;
        h.off1 = hdr[r + 1].ev_off;
    }
    return h;
}

// the first / last 16 bytes of the whole event array: never read outside of it
__device__ __noinline__ uint4 ev_load_edge(const uint8_t* p, const uint8_t* events, const uint8_t* ev_end) {
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    for (int k = 0; k < 16; ++k)
        if (p + k >= events && p + k < ev_end) w[k >> 2] |= static_cast<uint32_t>(p[k]) << (8 * (k & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// this lane's 16 bytes of the 512-byte window that starts at the 16-byte boundary below the read's first event byte
__device__ __forceinline__ uint4 ev_load_window(const uint8_t* __restrict__ events, const uint8_t* ev_end, const EvHdr& h, int lane) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    const uint8_t* first = events + h.off0;
    const uint8_t* p = first - (reinterpret_cast<uintptr_t>(first) & 15u) + 16 * lane;
    if (p < events + h.off1) {
        if (p + 16 <= ev_end && p >= events) v = *reinterpret_cast<const uint4*>(p);
        else v = ev_load_edge(p, events, ev_end);
    }
    return v;
}

// bits [0, k) set, k in 0..32 (PTX shifts clamp at the register width)
__device__ __forceinline__ uint32_t ev_mask_below(int32_t k) {
    uint32_t d;
    asm("shr.b32 %0, %1, %2;" : "=r"(d) : "r"(0xffffffffu), "r"(32 - k));
    return d;
}

// inclusive warp prefix sum; the shuffle's own predicate says whether the source lane exists
#define MS_SCAN_STEP(d) asm volatile("{ .reg .pred p; .reg .s32 t; shfl.sync.up.b32 t|p, %0, " #d ", 0, 0xffffffff; @p add.s32 %0, %0, t; }" : "+r"(s))
__device__ __forceinline__ int32_t ev_warp_scan(int32_t s) {
    MS_SCAN_STEP(1); MS_SCAN_STEP(2); MS_SCAN_STEP(4); MS_SCAN_STEP(8); MS_SCAN_STEP(16);
    return s;
}
#undef MS_SCAN_STEP

// row word ^= v when x has any bit of `bits`
__device__ __forceinline__ void ev_xor_if(uint32_t x, uint32_t bits, uint32_t addr, uint32_t v) {
    asm volatile("{ .reg .pred q; .reg .b32 t; and.b32 t, %0, %1; setp.ne.u32 q, t, 0; @q red.shared.xor.b32 [%2], %3; }"
                 :: "r"(x), "r"(bits), "r"(addr), "r"(v) : "memory");
}

// One list of a read, 32 entries at a time: a warp scan turns the deltas into columns, every lane XORs its column's difference
// from the base into the four plane words of the row (shared-memory reductions: XOR commutes, and two lanes hit the same word
// only when two events share a 32-column block).  NLIST: one byte per entry, nibble 5; else 12 bits per entry.  `src` is a
// generic pointer: the staged window in shared memory, or global memory for strings longer than the window.
template <bool NLIST>
__device__ __forceinline__ void ev_apply_list(const uint8_t* __restrict__ src, uint32_t n, int32_t begin, int32_t end, int32_t nblk,
                                              const uint2* __restrict__ base_sm, uint32_t row_addr, int lane) {
    int32_t carry = begin;
#pragma unroll 1
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
        const uint32_t i = i0 + lane;
        uint32_t delta = 0u, nib = 5u;
        if (i < n) {
            if (NLIST) {
                delta = src[i];
            } else {                                       // entry i sits in bits [12 i, 12 i + 12) of the list
                const uint32_t o = (3u * i) >> 1;
                const uint32_t two = static_cast<uint32_t>(src[o]) | (static_cast<uint32_t>(src[o + 1]) << 8);
                const uint32_t e = (i & 1u) ? two >> 4 : two & 0xfffu;
                delta = e >> 4; nib = e & 15u;
            }
        }
        const bool event = i < n && delta != 255u;         // 255 only moves on
        const int32_t c = carry + ev_warp_scan(static_cast<int32_t>(delta));
        carry = __shfl_sync(0xffffffffu, c, 31);
        const int32_t blk = c >> 5;
        const int bit = c & 31;
        if (event && blk < nblk) {                         // a column past the row can only come from a corrupt list
            const uint2 bp = base_sm[blk];
            const uint32_t basenib = (c >= begin && c < end) ? (((bp.x >> bit) & 1u) | (((bp.y >> bit) & 1u) << 1)) : 7u;
            const uint32_t x = nib ^ basenib;
            const uint32_t w = row_addr + 16u * static_cast<uint32_t>(blk);
            const uint32_t v = 1u << bit;
            ev_xor_if(x, 1u, w, v);
            ev_xor_if(x, 2u, w + 4u, v);
            ev_xor_if(x, 4u, w + 8u, v);
            if (!NLIST) ev_xor_if(x, 8u, w + 12u, v);
        }
    }
}

// One warp per read, eight warps = one tile (rows.cuh) per CTA and iteration; persistent CTAs stride over the tiles.
// Global-memory latency is taken out of the loop: a warp loads the header of its read two iterations ahead and the first
// 512 bytes of its event string (nearly always all of it) one iteration ahead, into registers, and decodes from shared
// memory.  The row is built in the warp's shared-memory slice: the base row (shared-memory copy) masked by the span, then
// the two lists are applied (ev_apply_list).  COOP: the eight warps then write the tile together, 512 contiguous bytes per
// warp instruction (rows are kept an odd number of 16-byte words apart, so these reads are bank-conflict free); rows too
// long for eight of them to fit in shared memory leave with per-warp scattered 16-byte stores instead.
template <bool COOP>
__global__ void __launch_bounds__(kExpandWarps * 32) expand_events_kernel(const ms_read_hdr* __restrict__ hdr,
                                                                        const uint8_t* __restrict__ events, int64_t R,
                                                                        int32_t nblk, int32_t rstride, const uint2* __restrict__ basepl,
                                                                        uint4* __restrict__ out) {
    extern __shared__ __align__(16) uint4 rows_sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    uint4* row = rows_sm + warp * rstride;
    const uint32_t row_addr = ev_smem_u32(row);
    uint4* stage = rows_sm + wpc * rstride + warp * (kStageBytes / 16);
    const uint8_t* stage_b = reinterpret_cast<const uint8_t*>(stage);
    uint2* base_sm = reinterpret_cast<uint2*>(rows_sm + wpc * rstride + wpc * (kStageBytes / 16));
    for (int32_t b = threadIdx.x; b < nblk; b += blockDim.x) base_sm[b] = basepl[b];
    const uint8_t* ev_end = events + hdr[R].ev_off;
    const int64_t Rpad = ((R + 7) >> 3) << 3;                // whole tiles (COOP: every warp of the CTA takes part in the barriers)
    const int64_t step = static_cast<int64_t>(gridDim.x) * wpc;
    const int64_t iters = (Rpad + step - 1) / step;
    const int64_t rfirst = static_cast<int64_t>(blockIdx.x) * wpc + warp;
    EvHdr h0 = ev_load_hdr(hdr, rfirst, R), h1 = ev_load_hdr(hdr, rfirst + step, R);
    uint4 win = ev_load_window(events, ev_end, h0, lane);
    __syncthreads();
    int64_t r = rfirst;
    for (int64_t it = 0; it < iters; ++it, r += step) {
        // in flight during this iteration: the next read's event bytes, the header of the one after it
        const uint4 win_next = ev_load_window(events, ev_end, h1, lane);
        const EvHdr h2 = ev_load_hdr(hdr, r + 2 * step, R);
        if (r < R) {
            const int32_t begin = static_cast<int32_t>(h0.span & 0xffffu), end = static_cast<int32_t>(h0.span >> 16);
            stage[lane] = win;
            for (int32_t b = lane; b < nblk; b += 32) {
                const uint2 bp = base_sm[b];
                const int32_t c0 = 32 * b;
                uint4 v = make_uint4(bp.x, bp.y, 0u, 0u);
                if (c0 < begin || c0 + 32 > end) {           // a block at the edge of (or outside) the span
                    const uint32_t m = ev_mask_below(min(32, max(0, end - c0))) & ~ev_mask_below(min(32, max(0, begin - c0)));
                    v = make_uint4((bp.x & m) | ~m, (bp.y & m) | ~m, ~m, 0u);
                }
                row[b] = v;
            }
            __syncwarp();
            // (a header whose offsets run backwards or past the array can only be corrupt: the read keeps the base on its span)
            const uint32_t nbytes = (h0.off1 >= h0.off0 && events + h0.off1 <= ev_end) ? h0.off1 - h0.off0 : 0u;
            if (nbytes >= 2u) {
                const uint8_t* ev = events + h0.off0;
                const uint32_t mis = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(ev) & 15u);
                const bool staged = mis + nbytes <= static_cast<uint32_t>(kStageBytes);     // (warp-uniform)
                const uint8_t* str = staged ? stage_b + mis : ev;
                uint32_t nN = static_cast<uint32_t>(str[0]) | (static_cast<uint32_t>(str[1]) << 8);
                nN = min(nN, nbytes - 2u);                                   // a corrupt count cannot run past the string
                const uint32_t nrest = (2u * (nbytes - 2u - nN)) / 3u;      // ceil(1.5 n) bytes hold n 12-bit entries
                ev_apply_list<true>(str + 2, nN, begin, end, nblk, base_sm, row_addr, lane);
                ev_apply_list<false>(str + 2 + nN, nrest, begin, end, nblk, base_sm, row_addr, lane);
            }
            __syncwarp();
        }
        h0 = h1; h1 = h2; win = win_next;
        if (COOP) {
            __syncthreads();
            // slot i of the tile holds block i >> 3 of read (i & 7) ^ (block & 7)
            const int64_t r0 = r - warp;                                     // first read of this CTA's tile (a multiple of 8)
            if (r0 < Rpad) {
                const int nslots = nblk * 8;
                uint4* dst = out + static_cast<size_t>(r0) * nblk + threadIdx.x;
                if (r0 + 8 <= R) {
                    for (int i = threadIdx.x; i < nslots; i += kExpandWarps * 32, dst += kExpandWarps * 32) {
                        const int b = i >> 3;
                        *dst = rows_sm[((i ^ b) & 7) * rstride + b];
                    }
                } else {                                                     // the last tile: its padding is "not spanned"
                    for (int i = threadIdx.x; i < nslots; i += kExpandWarps * 32, dst += kExpandWarps * 32) {
                        const int b = i >> 3;
                        const int w = (i ^ b) & 7;
                        *dst = r0 + w < R ? rows_sm[w * rstride + b] : make_uint4(~0u, ~0u, ~0u, 0u);
                    }
                }
            }
            __syncthreads();
        } else if (r < Rpad) {     // the padding of the last tile is "not spanned" here as well
            for (int32_t b = lane; b < nblk; b += 32) out[tile_slot(r, b, nblk)] = r < R ? row[b] : make_uint4(~0u, ~0u, ~0u, 0u);
            __syncwarp();
        }
    }
}

}  // namespace ms

static int expand_smem_bytes(int wpc, int nblk) {
    return wpc * ((nblk | 1) * 16 + ms::kStageBytes) + nblk * 8;
}

static int expand_launch(ms_handle* h, const ms_read_hdr* d_hdr, const uint8_t* d_events, int64_t R, uint32_t* d_packed) {
    const int rstride = h->nblk | 1;                      // uint4 between the rows in shared memory: odd
    const int smem_cap = std::min(h->max_smem, 100 << 10);
    const bool coop = expand_smem_bytes(ms::kExpandWarps, h->nblk) <= smem_cap;
    int wpc = ms::kExpandWarps;
    while (!coop && wpc > 1 && expand_smem_bytes(wpc, h->nblk) > smem_cap) --wpc;
    const int smem = expand_smem_bytes(wpc, h->nblk);
    if (smem > smem_cap) MS_FAIL(h, MS_ERR_ARG, "reference too long for the event expansion's shared memory");
    if (h->expand_ctas_nblk != h->nblk) {                 // resident CTAs per SM for this row length
        int n = 0;
        cudaError_t e = coop ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ms::expand_events_kernel<true>, wpc * 32, smem)
                             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ms::expand_events_kernel<false>, wpc * 32, smem);
        MS_CUDA(h, e);
        h->expand_ctas = std::max(1, n);
        h->expand_ctas_nblk = h->nblk;
    }
    const int64_t want = (R + wpc - 1) / wpc;
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(want, static_cast<int64_t>(h->num_sms) * h->expand_ctas)));
    MS_STAGE_BEGIN(h, MS_STAGE_EXPAND);
    if (coop)
        ms::expand_events_kernel<true><<<grid, wpc * 32, smem, h->stream>>>(d_hdr, d_events, R, h->nblk, rstride, h->b_base.as<uint2>(),
                                                                           reinterpret_cast<uint4*>(d_packed));
    else
        ms::expand_events_kernel<false><<<grid, wpc * 32, smem, h->stream>>>(d_hdr, d_events, R, h->nblk, rstride, h->b_base.as<uint2>(),
                                                                            reinterpret_cast<uint4*>(d_packed));
    MS_STAGE_END(h, MS_STAGE_EXPAND);
    h->launches++;
    MS_CUDA(h, cudaGetLastError());
    return MS_OK;
}

void ms_events_set_smem_attr(int max_smem) {
    cudaFuncSetAttribute(ms::expand_events_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, std::min(max_smem, 100 << 10));
    cudaFuncSetAttribute(ms::expand_events_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, std::min(max_smem, 100 << 10));
}

extern "C" {

int64_t ms_events_bound(int32_t L) { return L > 0 ? 4 + ((static_cast<int64_t>(L) + L / 254 + 2) * 3 + 1) / 2 : 0; }

int ms_encode_rows(const uint32_t* packed, int64_t R, int32_t L, const uint8_t* base, ms_read_hdr* hdr, uint8_t* events,
                   int64_t cap, int64_t* nevents) {
    if (!packed || !base || !hdr || (!events && cap > 0) || R < 0 || L <= 0 || L > 65535 || cap < 0 || !nevents) return MS_ERR_ARG;
    std::vector<uint32_t> bpl;
    ms::base_planes(base, L, bpl);
    const int32_t rw = 4 * ((L + 31) / 32);
    int64_t n = 0;
    for (int64_t r = 0; r < R; ++r) {
        int32_t b, e;
        const int64_t k = ms::encode_row(packed + static_cast<size_t>(r) * rw, L, bpl.data(), events + n, cap - n, b, e);
        if (k < 0) return MS_ERR_CAPACITY;
        if (n + k > 0xffffffffLL) return MS_ERR_CAPACITY;
        hdr[r].ev_off = static_cast<uint32_t>(n);
        hdr[r].begin = static_cast<uint16_t>(b);
        hdr[r].end = static_cast<uint16_t>(e);
        n += k;
    }
    *nevents = n;
    return ms_events_seal(hdr, R, n, base, L);
}

int ms_encode_states(const uint8_t* states, int64_t R, int32_t L, const uint8_t* base, ms_read_hdr* hdr, uint8_t* events,
                     int64_t cap, int64_t* nevents) {
    if (!states || !base || !hdr || R < 0 || L <= 0 || L > 65535 || cap < 0 || !nevents) return MS_ERR_ARG;
    std::vector<uint32_t> bpl;
    ms::base_planes(base, L, bpl);
    const int32_t rw = 4 * ((L + 31) / 32);
    std::vector<uint32_t> row(static_cast<size_t>(rw));
    int64_t n = 0;
    for (int64_t r = 0; r < R; ++r) {
        int rc = ms_pack_states(states + static_cast<size_t>(r) * L, 1, L, row.data());
        if (rc != MS_OK) return rc;
        int32_t b, e;
        const int64_t k = ms::encode_row(row.data(), L, bpl.data(), events + n, cap - n, b, e);
        if (k < 0 || n + k > 0xffffffffLL) return MS_ERR_CAPACITY;
        hdr[r].ev_off = static_cast<uint32_t>(n);
        hdr[r].begin = static_cast<uint16_t>(b);
        hdr[r].end = static_cast<uint16_t>(e);
        n += k;
    }
    *nevents = n;
    return ms_events_seal(hdr, R, n, base, L);
}

int ms_encode_row(const uint32_t* row, int32_t L, const uint32_t* base_planes, ms_read_hdr* hdr, uint8_t* events, int64_t cap,
                  int64_t* nevents) {
    if (!row || !base_planes || !hdr || !nevents || L <= 0 || L > 65535 || *nevents < 0 || cap < *nevents) return MS_ERR_ARG;
    int32_t b, e;
    const int64_t k = ms::encode_row(row, L, base_planes, events + *nevents, cap - *nevents, b, e);
    if (k < 0 || *nevents + k > 0xffffffffLL) return MS_ERR_CAPACITY;
    hdr->ev_off = static_cast<uint32_t>(*nevents);
    hdr->begin = static_cast<uint16_t>(b);
    hdr->end = static_cast<uint16_t>(e);
    *nevents += k;
    return MS_OK;
}

int ms_base_planes(const uint8_t* base, int32_t L, uint32_t* planes) {
    if (!base || !planes || L <= 0) return MS_ERR_ARG;
    std::vector<uint32_t> pl;
    ms::base_planes(base, L, pl);
    memcpy(planes, pl.data(), pl.size() * 4);
    return MS_OK;
}

int ms_events_seal(ms_read_hdr* hdr, int64_t R, int64_t nevents, const uint8_t* base, int32_t L) {
    if (!hdr || !base || R < 0 || nevents < 0 || nevents > 0xffffffffLL || L <= 0 || L > 65535) return MS_ERR_ARG;
    const uint32_t hsh = ms::base_hash(base, L);
    hdr[R].ev_off = static_cast<uint32_t>(nevents);
    hdr[R].begin = static_cast<uint16_t>(hsh & 0xffffu);
    hdr[R].end = static_cast<uint16_t>(hsh >> 16);
    return MS_OK;
}

int ms_set_base(ms_handle* h, const uint8_t* base) {
    if (!h || !base || !h->d_counts) return MS_ERR_ARG;
    if (h->L > 65535) MS_FAIL(h, MS_ERR_ARG, "event rows need a reference of at most 65535 columns");
    MS_CUDA(h, cudaSetDevice(h->device));
    std::vector<uint32_t> pl;
    ms::base_planes(base, h->L, pl);
    MS_CUDA(h, cudaStreamSynchronize(h->stream));
    MS_CUDA(h, h->b_base.ensure(pl.size() * 4));
    MS_CUDA(h, cudaMemcpyAsync(h->b_base.p, pl.data(), pl.size() * 4, cudaMemcpyHostToDevice, h->stream));
    MS_CUDA(h, cudaStreamSynchronize(h->stream));
    h->base_hash = ms::base_hash(base, h->L);
    h->have_base = true;
    return MS_OK;
}

int ms_expand_events_dev(ms_handle* h, const ms_read_hdr* d_hdr, const uint8_t* d_events, int64_t R, uint32_t* d_packed) {
    MsRange nvtx_range("expand events");
    if (!h || !h->d_counts || R < 0 || (R > 0 && (!d_hdr || !d_packed))) return MS_ERR_ARG;
    if (!h->have_base) MS_FAIL(h, MS_ERR_ARG, "ms_set_base has not been called for this layout");
    if (R == 0) return MS_OK;
    MS_CUDA(h, cudaSetDevice(h->device));
    return expand_launch(h, d_hdr, d_events, R, d_packed);
}

// Event rows from host memory (pinned recommended): the header and event arrays go up in a few chunks on the copy
// stream; behind each chunk the main stream expands it into the handle's row buffer and piles it up, so that only the
// last chunk's kernels are not hidden behind the PCIe transfer.
int ms_pileup_events_host(ms_handle* h, const ms_read_hdr* hdr, const uint8_t* events, int64_t R, const uint32_t** keep_dev) {
    MsRange nvtx_range("H2D events + expand + K1");
    if (!h || !h->d_counts || R < 0 || !hdr || (R > 0 && !events && hdr[R].ev_off > 0)) return MS_ERR_ARG;
    if (!h->have_base) MS_FAIL(h, MS_ERR_ARG, "ms_set_base has not been called for this layout");
    if ((static_cast<uint32_t>(hdr[R].begin) | (static_cast<uint32_t>(hdr[R].end) << 16)) != h->base_hash)
        MS_FAIL(h, MS_ERR_FORMAT, "event rows were encoded against a different base sequence (or are not sealed)");
    MS_CUDA(h, cudaSetDevice(h->device));
    const size_t row_bytes = static_cast<size_t>(h->nblk) * 16;
    const size_t need = std::max<size_t>(16, static_cast<size_t>(ms::tiles_of(R)) * 8 * row_bytes);   // whole tiles (rows.cuh)
    const int64_t total_ev = hdr[R].ev_off;
    if (need > h->upload_cap || static_cast<size_t>(R + 1) * sizeof(ms_read_hdr) > h->b_ev_hdr.cap || static_cast<size_t>(total_ev) + 64 > h->b_ev.cap) {
        MS_CUDA(h, cudaStreamSynchronize(h->stream));
        if (need > h->upload_cap) {
            cudaFree(h->d_upload);
            h->d_upload = nullptr; h->upload_cap = 0;
            MS_CUDA(h, cudaMalloc(&h->d_upload, need));
            h->upload_cap = need;
        }
        MS_CUDA(h, h->b_ev_hdr.ensure(static_cast<size_t>(R + 1) * sizeof(ms_read_hdr)));
        MS_CUDA(h, h->b_ev.ensure(static_cast<size_t>(total_ev) + 64));
    }
    ms_read_hdr* d_hdr = h->b_ev_hdr.as<ms_read_hdr>();
    uint8_t* d_ev = h->b_ev.as<uint8_t>();
    // chunking: ~12 MB of payload per chunk (0.2 ms on the link; 12 vs 24 MB: 3.2 vs 3.3 ms per 1M reads), at most 16 chunks, at least 64 Ki reads per chunk.  Only the LAST
    // chunk's expansion + pile-up is not hidden behind a transfer, so the chunks shrink towards the end: the last one holds a
    // quarter of an even share (but no less than 64 Ki reads), the ones before it make up for it.
    const double payload = static_cast<double>(total_ev) + static_cast<double>(R) * 8;
    static const double chunk_mb = getenv("MS_EVENTS_CHUNK_MB") ? atof(getenv("MS_EVENTS_CHUNK_MB")) : 12.0;
    int64_t nchunks = std::max<int64_t>(1, std::min<int64_t>(16, static_cast<int64_t>(payload / (chunk_mb * 1048576.0) + 0.5)));
    nchunks = std::max<int64_t>(1, std::min<int64_t>(nchunks, R / 65536));
    std::vector<int64_t> bounds(static_cast<size_t>(nchunks) + 1, 0);
    static const bool uniform = getenv("MS_EVENTS_UNIFORM") != nullptr;      // A/B switch (tools/expand_bench.py)
    if (uniform) {
        for (int64_t k = 1; k < nchunks; ++k) bounds[k] = (R * k / nchunks) & ~static_cast<int64_t>(7);
        bounds[nchunks] = R;
    } else {
        const int64_t even = R / nchunks;
        const int64_t last = nchunks > 1 ? std::max<int64_t>(std::min<int64_t>(even, 65536), even / 4) : R;
        const int64_t second = nchunks > 2 ? std::max<int64_t>(std::min<int64_t>(even, 65536), even / 2) : 0;
        const int64_t front = nchunks > 2 ? nchunks - 2 : (nchunks > 1 ? 1 : 0);
        const int64_t rest = R - last - second;
        for (int64_t k = 1; k <= front; ++k) bounds[k] = (rest * k / front) & ~static_cast<int64_t>(7);   // whole tiles
        if (nchunks > 2) bounds[nchunks - 1] = (rest + second) & ~static_cast<int64_t>(7);
        bounds[nchunks] = R;
    }
    // the copy stream must not overwrite the staging buffers before earlier work on the main stream is done
    MS_CUDA(h, cudaEventRecord(h->ev_copy[1], h->stream));
    MS_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_copy[1], 0));
    if (h->timing) cudaEventRecord(h->ev_stage[MS_STAGE_UPLOAD][0], h->copy_stream);
    // MS_TRACE_CHUNKS=1 (debug aid): when each chunk's copy, expansion and pile-up ended, relative to the start of the upload
    static const bool trace = getenv("MS_TRACE_CHUNKS") != nullptr;
    std::vector<cudaEvent_t> tr;
    if (trace) {
        tr.resize(static_cast<size_t>(3 * nchunks + 1));
        for (cudaEvent_t& e : tr) cudaEventCreate(&e);
        cudaEventRecord(tr[0], h->copy_stream);
    }
    for (int64_t k = 0; k < nchunks; ++k) {
        const int64_t r0 = bounds[k], r1 = bounds[k + 1];
        if (r1 <= r0) continue;
        const int64_t e0 = hdr[r0].ev_off, e1 = hdr[r1].ev_off;
        if (e1 < e0 || e1 > total_ev) MS_FAIL(h, MS_ERR_FORMAT, "event rows: header offsets are not ascending");
        const int64_t h0 = r0 + (k > 0 ? 1 : 0);     // entry r0 went up with the previous chunk (as its end marker)
        MS_CUDA(h, cudaMemcpyAsync(d_hdr + h0, hdr + h0, static_cast<size_t>(r1 - h0 + 1) * sizeof(ms_read_hdr), cudaMemcpyHostToDevice, h->copy_stream));
        if (e1 > e0) MS_CUDA(h, cudaMemcpyAsync(d_ev + e0, events + e0, static_cast<size_t>(e1 - e0), cudaMemcpyHostToDevice, h->copy_stream));
        cudaEvent_t ev = h->ev_chunk[k & 15];
        MS_CUDA(h, cudaEventRecord(ev, h->copy_stream));
        if (h->timing && k == nchunks - 1) { cudaEventRecord(h->ev_stage[MS_STAGE_UPLOAD][1], h->copy_stream); h->stage_seen[MS_STAGE_UPLOAD] = true; }
        MS_CUDA(h, cudaStreamWaitEvent(h->stream, ev, 0));
        uint32_t* dst = h->d_upload + static_cast<size_t>(r0) * (row_bytes / 4);
        if (trace) cudaEventRecord(tr[static_cast<size_t>(3 * k + 1)], h->copy_stream);
        int rc = expand_launch(h, d_hdr + r0, d_ev, r1 - r0, dst);
        if (rc != MS_OK) return rc;
        if (trace) cudaEventRecord(tr[static_cast<size_t>(3 * k + 2)], h->stream);
        rc = ms_pileup_dev(h, dst, r1 - r0);
        if (rc != MS_OK) return rc;
        if (trace) cudaEventRecord(tr[static_cast<size_t>(3 * k + 3)], h->stream);
    }
    if (trace) {
        cudaStreamSynchronize(h->stream);
        for (int64_t k = 0; k < nchunks; ++k) {
            float a = 0.f, b = 0.f, c = 0.f;
            cudaEventElapsedTime(&a, tr[0], tr[static_cast<size_t>(3 * k + 1)]);
            cudaEventElapsedTime(&b, tr[0], tr[static_cast<size_t>(3 * k + 2)]);
            cudaEventElapsedTime(&c, tr[0], tr[static_cast<size_t>(3 * k + 3)]);
            fprintf(stderr, "chunk %2lld: %7lld reads  copied %.3f  expanded %.3f  piled up %.3f ms\n", static_cast<long long>(k),
                    static_cast<long long>(bounds[k + 1] - bounds[k]), a, b, c);
        }
        for (cudaEvent_t e : tr) cudaEventDestroy(e);
    }
    if (keep_dev) *keep_dev = h->d_upload;
    return MS_OK;
}

}  // extern "C"
