// events.cu -- compact "event rows": the host->device form of an aligned read, and its expansion on the GPU.
//
// A PacBio CCS alignment is "the reference, except at a few columns" (CIGAR `=` runs with sparse X / D / I ops,
// /root/reference/doc/JULIET.md:49-58, plus the QV-filtered bases that become N, :256-259).  The planar rows K1 and K3
// read cost L/2 bytes per read on the PCIe link (1504 B at 3 kb), which is what bounds the end-to-end pass.  Here the
// host ships, per read, its span and a sorted list of 12-bit events (column delta, new 4-bit column value) against a
// base sequence both sides hold (~85 events = ~136 B per 3 kb read at CCS error rates), and expand_events_kernel
// rebuilds the packed reads in HBM (as tiles, rows.cuh), where the pile-up and the phasing kernels run unchanged.  SURVEY.md rows a2/a3
// (host CIGAR walk) and 8f-2 ("GPU-side CIGAR expansion is the next real speed-up").
//
// Format (include/minorseq_b200.h):  ms_read_hdr hdr[R+1] = {ev_off, begin, end}; the byte string of read r is
// events[hdr[r].ev_off .. hdr[r+1].ev_off), event k in its bits [12k, 12k+12): delta << 4 | nibble, the column is the
// previous event's column (the read's `begin` for the first) + delta, nibble = state | insertion-follows << 3 of that
// column.  Every spanned column without an event holds the base sequence's base, columns outside [begin, end) are "not
// spanned".  The encoder emits a filler event (a column's unchanged value) when two events are more than 255 columns
// apart.  hdr[R] is a sentinel: ev_off = total number of event bytes, begin | end << 16 = a 32-bit hash of the base
// sequence, so that rows encoded against another base are rejected instead of silently mis-expanded.
#include <algorithm>
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

constexpr int32_t kMaxDelta = 255;

// appends 12-bit events to a read's byte string
struct EventWriter {
    uint8_t* ev;
    int64_t cap, o = 0;   // o: bytes completely or partly written
    bool half = false;    // the low nibble of ev[o-1]... see put(): an odd number of events so far
    bool put(uint32_t e) {
        if (!half) {
            if (o + 2 > cap) return false;
            ev[o] = static_cast<uint8_t>(e & 0xffu);
            ev[o + 1] = static_cast<uint8_t>(e >> 8);          // low nibble; the next event fills the high one
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
    EventWriter w{ev, cap};
    int32_t prev = begin;
    for (int32_t b = begin >> 5; b <= (end - 1) >> 5; ++b) {
        const uint32_t* q = row + 4 * b;
        uint32_t diff = ((q[0] ^ bpl[2 * b]) | (q[1] ^ bpl[2 * b + 1]) | q[2] | q[3]) & span_mask(b, begin, end);
        while (diff) {
            const int32_t c = 32 * b + __builtin_ctz(diff);
            diff &= diff - 1;
            while (c - prev > kMaxDelta) {   // filler: restate an unchanged column
                prev += kMaxDelta;
                if (!w.put((static_cast<uint32_t>(kMaxDelta) << 4) | nibble_at(row, prev))) return -1;
            }
            if (!w.put((static_cast<uint32_t>(c - prev) << 4) | nibble_at(row, c))) return -1;
            prev = c;
        }
    }
    return w.o;
}

// ---------------------------------------------------------------- device side
__device__ __forceinline__ uint32_t ev_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

constexpr int kExpandMaxWarps = 16;

// One warp per read, eight warps per tile (rows.cuh).  The row is built in the warp's shared-memory slice: base planes
// masked by the span, then the events are applied 32 at a time -- a warp scan turns the deltas into columns, lanes whose
// events fall into the same 32-column block (consecutive lanes: the events are sorted) take turns, one read-modify-write
// per event.  No atomics.  COOP: the eight warps of a tile then write it together, 512 contiguous bytes per warp
// instruction (rows are kept an odd number of 16-byte words apart, so these reads are bank-conflict free); rows too long
// for eight of them to fit in shared memory leave with per-warp scattered 16-byte stores instead.
template <bool COOP>
__global__ void __launch_bounds__(kExpandMaxWarps * 32) expand_events_kernel(const ms_read_hdr* __restrict__ hdr,
                                                                           const uint8_t* __restrict__ events, int64_t R,
                                                                           int32_t nblk, int32_t rstride, const uint2* __restrict__ basepl,
                                                                           uint4* __restrict__ out) {
    extern __shared__ __align__(16) uint4 rows_sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    uint4* row = rows_sm + static_cast<size_t>(warp) * rstride;
    const int64_t Rpad = ((R + 7) >> 3) << 3;                // whole tiles (COOP: every warp of the CTA takes part in the barriers)
    const int64_t step = static_cast<int64_t>(gridDim.x) * wpc;
    const int64_t iters = (Rpad + step - 1) / step;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t r = (it * gridDim.x + blockIdx.x) * wpc + warp;
        if (r < R) {
            const ms_read_hdr hd = hdr[r];
            const uint32_t off1 = hdr[r + 1].ev_off;
            const int32_t begin = hd.begin, end = hd.end;
            for (int32_t b = lane; b < nblk; b += 32) {
                const int32_t lo = max(0, begin - 32 * b), hi = min(32, end - 32 * b);
                uint32_t m = 0u;
                if (hi > lo) m = (hi == 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
                const uint2 bp = basepl[b];
                row[b] = make_uint4((bp.x & m) | ~m, (bp.y & m) | ~m, ~m, 0u);
            }
            __syncwarp();
            const uint32_t n = (2u * (off1 - hd.ev_off)) / 3u;        // ceil(1.5 n) bytes hold n 12-bit events
            const uint8_t* ev = events + hd.ev_off;
            int32_t carry = begin;
            for (uint32_t i0 = 0; i0 < n; i0 += 32) {
                const uint32_t i = i0 + lane;
                const bool have = i < n;
                uint32_t e = 0u;
                if (have) {                                            // event i sits in bits [12 i, 12 i + 12) of the byte string
                    const uint32_t o = (3u * i) >> 1;
                    const uint32_t two = static_cast<uint32_t>(ev[o]) | (static_cast<uint32_t>(ev[o + 1]) << 8);
                    e = (i & 1u) ? two >> 4 : two & 0xfffu;
                }
                int32_t s = static_cast<int32_t>(e >> 4);
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int32_t t = __shfl_up_sync(0xffffffffu, s, d);
                    if (lane >= d) s += t;
                }
                const int32_t c = carry + s;
                carry = __shfl_sync(0xffffffffu, c, 31);
                const int32_t blk = c >> 5;
                const int bit = c & 31;
                const bool ok = have && blk < nblk;            // a column past the row can only come from a corrupt list
                uint32_t x = 0u;
                if (ok) {
                    const uint2 bp = basepl[blk];
                    const uint32_t basenib = (c >= begin && c < end) ? (((bp.x >> bit) & 1u) | (((bp.y >> bit) & 1u) << 1)) : 7u;
                    x = (e & 15u) ^ basenib;
                }
                // lanes of one 32-column block are consecutive (the events are sorted): a lane's position inside its block's run comes
                // from one ballot of the run heads; the lanes then update shared memory in rounds, round k = the k-th event of every
                // block, so no two lanes of a round touch the same block.  Runs are 1-3 lanes long at CCS error rates.  (MATCH.ANY and
                // REDUX on per-group masks both serialise over the ~25 distinct groups of a batch and were ~30x slower.)
                const int32_t key = ok ? blk : -1;
                const int32_t left = __shfl_up_sync(0xffffffffu, key, 1);
                const uint32_t heads = __ballot_sync(0xffffffffu, lane == 0 || key != left);
                const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
                const int rank = (ok && x) ? lane - start : -1;
                for (int k = 0; __any_sync(0xffffffffu, rank >= k); ++k) {
                    if (rank == k) {
                        uint4 v = row[blk];
                        v.x ^= (x & 1u) << bit; v.y ^= ((x >> 1) & 1u) << bit; v.z ^= ((x >> 2) & 1u) << bit; v.w ^= ((x >> 3) & 1u) << bit;
                        row[blk] = v;
                    }
                    __syncwarp();
                }
            }
        }
        if (COOP) {
            __syncthreads();
            // the CTA's tiles are contiguous in the output: slot i of them holds block (i >> 3) % nblk of read pos ^ (block & 7)
            const int64_t r0 = (it * gridDim.x + blockIdx.x) * wpc;          // first read of this CTA's tiles (a multiple of 8)
            const int nslots = (wpc >> 3) * nblk * 8;
            for (int i = threadIdx.x; i < nslots; i += blockDim.x) {
                const int tb = i >> 3;
                const int tl = tb / nblk, b = tb - tl * nblk;
                const int w = tl * 8 + ((i & 7) ^ (b & 7));
                if (r0 + tl * 8 < Rpad)
                    out[static_cast<size_t>(r0) * nblk + i] = r0 + w < R ? rows_sm[static_cast<size_t>(w) * rstride + b] : make_uint4(~0u, ~0u, ~0u, 0u);
            }
            __syncthreads();
        } else if (r < Rpad) {     // the padding of the last tile is "not spanned" here as well
            for (int32_t b = lane; b < nblk; b += 32) out[tile_slot(r, b, nblk)] = r < R ? row[b] : make_uint4(~0u, ~0u, ~0u, 0u);
            __syncwarp();
        }
    }
}

}  // namespace ms

static int expand_launch(ms_handle* h, const ms_read_hdr* d_hdr, const uint8_t* d_events, int64_t R, uint32_t* d_packed) {
    const int rstride = h->nblk | 1;                      // uint4 between the rows in shared memory: odd
    const int row_bytes = rstride * 16;
    const int smem_cap = std::min(h->max_smem, 100 << 10);
    const bool coop = 8 * row_bytes <= smem_cap;
    int wpc = coop ? (16 * row_bytes <= (64 << 10) ? 16 : 8) : std::max(1, std::min(ms::kExpandMaxWarps, smem_cap / row_bytes));
    if (!coop && row_bytes > smem_cap) MS_FAIL(h, MS_ERR_ARG, "reference too long for the event expansion's shared memory");
    const int64_t want = (R + wpc - 1) / wpc;
    const int ctas_per_sm = std::max(1, 64 / wpc);      // 64 resident warps per SM (40 registers), grid-stride over the reads
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(want, static_cast<int64_t>(h->num_sms) * ctas_per_sm)));
    MS_STAGE_BEGIN(h, MS_STAGE_EXPAND);
    if (coop)
        ms::expand_events_kernel<true><<<grid, wpc * 32, wpc * row_bytes, h->stream>>>(d_hdr, d_events, R, h->nblk, rstride, h->b_base.as<uint2>(),
                                                                                      reinterpret_cast<uint4*>(d_packed));
    else
        ms::expand_events_kernel<false><<<grid, wpc * 32, wpc * row_bytes, h->stream>>>(d_hdr, d_events, R, h->nblk, rstride, h->b_base.as<uint2>(),
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

int64_t ms_events_bound(int32_t L) { return L > 0 ? ((static_cast<int64_t>(L) + L / ms::kMaxDelta + 2) * 3 + 1) / 2 : 0; }

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
    for (int64_t k = 0; k < nchunks; ++k) {
        const int64_t r0 = bounds[k], r1 = bounds[k + 1];
        if (r1 <= r0) continue;
        const int64_t e0 = hdr[r0].ev_off, e1 = hdr[r1].ev_off;
        const int64_t h0 = r0 + (k > 0 ? 1 : 0);     // entry r0 went up with the previous chunk (as its end marker)
        MS_CUDA(h, cudaMemcpyAsync(d_hdr + h0, hdr + h0, static_cast<size_t>(r1 - h0 + 1) * sizeof(ms_read_hdr), cudaMemcpyHostToDevice, h->copy_stream));
        if (e1 > e0) MS_CUDA(h, cudaMemcpyAsync(d_ev + e0, events + e0, static_cast<size_t>(e1 - e0), cudaMemcpyHostToDevice, h->copy_stream));
        cudaEvent_t ev = h->ev_chunk[k & 15];
        MS_CUDA(h, cudaEventRecord(ev, h->copy_stream));
        MS_CUDA(h, cudaStreamWaitEvent(h->stream, ev, 0));
        uint32_t* dst = h->d_upload + static_cast<size_t>(r0) * (row_bytes / 4);
        int rc = expand_launch(h, d_hdr + r0, d_ev, r1 - r0, dst);
        if (rc != MS_OK) return rc;
        rc = ms_pileup_dev(h, dst, r1 - r0);
        if (rc != MS_OK) return rc;
    }
    if (keep_dev) *keep_dev = h->d_upload;
    return MS_OK;
}

}  // extern "C"
