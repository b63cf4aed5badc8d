// fast_inflate.hpp -- raw DEFLATE decoder for BGZF members (RFC 1951), written for the one thing the tools spend
// their wall clock on: inflating the BAM.  zlib's inflate runs at ~0.45 GB/s on CCS BAM blocks; this decoder keeps a
// 64-bit bit buffer refilled eight bytes at a time, resolves a literal or a length + its extra-bit count with ONE
// table lookup (11-bit primary table, sub-tables for longer codes), and copies matches eight bytes at a time.
// Every BGZF member carries a CRC-32 and its uncompressed size, so the caller verifies each block and falls back to
// zlib on any disagreement (bgzf_bam.hpp, inflate_file) -- a decoding bug can cost time, never correctness.
#pragma once
#include <cstdint>
#include <cstring>

namespace msinflate {

namespace detail {

constexpr int kLitBits = 11, kDistBits = 8;
constexpr int kLitTable = (1 << kLitBits) + 1024;     // primary + worst-case sub-tables (codes up to 15 bits)
constexpr int kDistTable = (1 << kDistBits) + 512;

// entry: bits 0-7 code length to consume (for a sub-table pointer: the primary bits), bits 8-11 extra-bit count,
// bits 12-15 kind, bits 16-31 literal / base length / base distance / sub-table offset
enum : uint32_t { kLiteral = 0u << 12, kBase = 1u << 12, kEnd = 2u << 12, kSub = 3u << 12, kBad = 4u << 12, kKindMask = 15u << 12 };

static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

inline uint32_t reverse_bits(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}

// symbol -> table entry payload (without the code length)
inline uint32_t litlen_payload(int sym) {
    if (sym < 256) return kLiteral | (static_cast<uint32_t>(sym) << 16);
    if (sym == 256) return kEnd;
    if (sym > 285) return kBad;
    return kBase | (static_cast<uint32_t>(kLenExtra[sym - 257]) << 8) | (static_cast<uint32_t>(kLenBase[sym - 257]) << 16);
}
inline uint32_t dist_payload(int sym) {
    if (sym > 29) return kBad;
    return kBase | (static_cast<uint32_t>(kDistExtra[sym]) << 8) | (static_cast<uint32_t>(kDistBase[sym]) << 16);
}

// canonical Huffman decode table from code lengths; returns false on an over-subscribed or (non-trivially) incomplete code
template <class Payload>
inline bool build_table(const uint8_t* lens, int nsym, int tbits, uint32_t* table, int table_cap, Payload payload) {
    int count[16] = {0};
    for (int s = 0; s < nsym; ++s) ++count[lens[s]];
    count[0] = 0;
    int left = 1, maxlen = 0, ncodes = 0;
    for (int l = 1; l <= 15; ++l) {
        left = (left << 1) - count[l];
        if (left < 0) return false;
        if (count[l]) { maxlen = l; ncodes += count[l]; }
    }
    const int psize = 1 << tbits;
    for (int i = 0; i < psize; ++i) table[i] = kBad | 1u;
    if (ncodes == 0) return true;                       // no codes at all (e.g. a block without distances)
    if (left > 0 && !(ncodes == 1 && count[1] == 1)) return false;    // incomplete: only the single-code case is legal
    uint32_t next[16];
    uint32_t code = 0;
    for (int l = 1; l <= 15; ++l) { code = (code + count[l - 1]) << 1; next[l] = code; }
    // pass 1: direct entries; remember, per primary index, the longest code that falls through it
    uint8_t sublen[1 << 11];
    memset(sublen, 0, static_cast<size_t>(psize));
    for (int s = 0; s < nsym; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t c = reverse_bits(next[l]++, l);
        if (l <= tbits) {
            const uint32_t e = payload(s) | static_cast<uint32_t>(l);
            for (uint32_t i = c; i < static_cast<uint32_t>(psize); i += 1u << l) table[i] = e;
        } else {
            const uint32_t p = c & (psize - 1);
            if (l - tbits > sublen[p]) sublen[p] = static_cast<uint8_t>(l - tbits);
        }
    }
    if (maxlen <= tbits) return true;
    // pass 2: sub-table offsets
    int used = psize;
    for (int p = 0; p < psize; ++p) {
        if (!sublen[p]) continue;
        if (used + (1 << sublen[p]) > table_cap) return false;
        table[p] = kSub | (static_cast<uint32_t>(sublen[p]) << 8) | static_cast<uint32_t>(tbits) | (static_cast<uint32_t>(used) << 16);
        for (int i = 0; i < (1 << sublen[p]); ++i) table[used + i] = kBad | 1u;
        used += 1 << sublen[p];
    }
    // pass 3: long codes into their sub-tables (codes are regenerated in the same order)
    code = 0;
    for (int l = 1; l <= 15; ++l) { code = (code + count[l - 1]) << 1; next[l] = code; }
    for (int s = 0; s < nsym; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t c = reverse_bits(next[l]++, l);
        if (l <= tbits) continue;
        const uint32_t p = c & (psize - 1);
        const uint32_t off = table[p] >> 16, sb = (table[p] >> 8) & 15u;
        const uint32_t e = payload(s) | static_cast<uint32_t>(l - tbits);
        for (uint32_t i = c >> tbits; i < (1u << sb); i += 1u << (l - tbits)) table[off + i] = e;
    }
    return true;
}

struct Bits {
    const uint8_t* in;
    const uint8_t* end;
    uint64_t buf = 0;
    int cnt = 0;
    // at least 56 valid bits afterwards (fewer only at the very end of the input, padded with zeros)
    inline void refill() {
        if (end - in >= 8) {
            uint64_t w;
            memcpy(&w, in, 8);
            buf |= w << cnt;
            in += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56 && in < end) { buf |= static_cast<uint64_t>(*in++) << cnt; cnt += 8; }
        }
    }
    inline uint32_t peek(int n) const { return static_cast<uint32_t>(buf) & ((1u << n) - 1u); }
    inline void drop(int n) { buf >>= n; cnt -= n; }
    inline uint32_t take(int n) { const uint32_t v = peek(n); drop(n); return v; }
};

}  // namespace detail

// Inflates exactly out_len bytes from a raw deflate stream.  Returns false on malformed input, on a stream that ends
// early or produces more than out_len bytes (the caller then uses zlib).
inline bool fast_inflate(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
    using namespace detail;
    Bits b{in, in + in_len};
    uint8_t* o = out;
    uint8_t* const oend = out + out_len;
    uint32_t lit[kLitTable], dst[kDistTable];
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    for (;;) {
        b.refill();
        if (b.cnt < 3) return false;
        const uint32_t final_block = b.take(1), type = b.take(2);
        if (type == 0) {                                     // stored
            b.drop(b.cnt & 7);
            b.refill();
            if (b.cnt < 32) return false;
            const uint32_t len = b.take(16), nlen = b.take(16);
            if ((len ^ nlen) != 0xffffu) return false;
            // give whole bytes in the bit buffer back to the byte stream
            const uint8_t* p = b.in - (b.cnt >> 3);
            b.buf = 0; b.cnt = 0;
            if (static_cast<size_t>(b.end - p) < len || static_cast<size_t>(oend - o) < len) return false;
            memcpy(o, p, len);
            o += len;
            b.in = p + len;
        } else if (type == 1 || type == 2) {
            uint8_t lens[288 + 32];
            int nlit, ndist;
            if (type == 1) {
                nlit = 288; ndist = 32;
                for (int i = 0; i < 144; ++i) lens[i] = 8;
                for (int i = 144; i < 256; ++i) lens[i] = 9;
                for (int i = 256; i < 280; ++i) lens[i] = 7;
                for (int i = 280; i < 288; ++i) lens[i] = 8;
                for (int i = 0; i < 32; ++i) lens[288 + i] = 5;
            } else {
                if (b.cnt < 14) return false;
                nlit = static_cast<int>(b.take(5)) + 257;
                ndist = static_cast<int>(b.take(5)) + 1;
                const int ncl = static_cast<int>(b.take(4)) + 4;
                if (nlit > 286 || ndist > 30) return false;
                uint8_t cl[19] = {0};
                for (int i = 0; i < ncl; ++i) { b.refill(); if (b.cnt < 3) return false; cl[order[i]] = static_cast<uint8_t>(b.take(3)); }
                uint32_t clt[(1 << 7) + 8];
                if (!build_table(cl, 19, 7, clt, (1 << 7) + 8, [](int s) { return kLiteral | (static_cast<uint32_t>(s) << 16); })) return false;
                int n = 0;
                while (n < nlit + ndist) {
                    b.refill();
                    const uint32_t e = clt[b.peek(7)];
                    if ((e & kKindMask) != kLiteral || static_cast<int>(e & 255u) > b.cnt) return false;
                    b.drop(static_cast<int>(e & 255u));
                    const int sym = static_cast<int>(e >> 16);
                    if (sym < 16) lens[n++] = static_cast<uint8_t>(sym);
                    else {
                        int rep;
                        uint8_t v = 0;
                        if (sym == 16) { if (n == 0) return false; v = lens[n - 1]; rep = 3 + static_cast<int>(b.take(2)); }
                        else if (sym == 17) rep = 3 + static_cast<int>(b.take(3));
                        else rep = 11 + static_cast<int>(b.take(7));
                        if (b.cnt < 0 || n + rep > nlit + ndist) return false;
                        while (rep--) lens[n++] = v;
                    }
                }
                if (lens[256] == 0) return false;            // no end-of-block code
                // distance lengths follow the literal/length ones: move them to a fixed place
                memmove(lens + 288, lens + nlit, static_cast<size_t>(ndist));
            }
            if (!build_table(lens, nlit, kLitBits, lit, kLitTable, litlen_payload)) return false;
            if (!build_table(lens + 288, ndist, kDistBits, dst, kDistTable, dist_payload)) return false;
            for (;;) {
                b.refill();
                uint32_t e = lit[b.peek(kLitBits)];
                if ((e & kKindMask) == kSub) {
                    b.drop(kLitBits);
                    e = lit[(e >> 16) + b.peek(static_cast<int>((e >> 8) & 15u))];
                }
                const int cl = static_cast<int>(e & 255u);
                if (cl > b.cnt) return false;                // ran out of input
                b.drop(cl);
                const uint32_t kind = e & kKindMask;
                if (kind == kLiteral) {
                    if (o >= oend) return false;
                    *o++ = static_cast<uint8_t>(e >> 16);
                    // a second and third literal without another refill: 56 bits cover three 15-bit codes
                    uint32_t e2 = lit[b.peek(kLitBits)];
                    if ((e2 & (kKindMask | 0u)) == kLiteral && o < oend && static_cast<int>(e2 & 255u) <= b.cnt) {
                        b.drop(static_cast<int>(e2 & 255u));
                        *o++ = static_cast<uint8_t>(e2 >> 16);
                        e2 = lit[b.peek(kLitBits)];
                        if ((e2 & kKindMask) == kLiteral && o < oend && static_cast<int>(e2 & 255u) <= b.cnt) {
                            b.drop(static_cast<int>(e2 & 255u));
                            *o++ = static_cast<uint8_t>(e2 >> 16);
                        }
                    }
                    continue;
                }
                if (kind == kEnd) break;
                if (kind != kBase) return false;
                const int xl = static_cast<int>((e >> 8) & 15u);
                uint32_t len = (e >> 16) + b.take(xl);
                b.refill();
                uint32_t d = dst[b.peek(kDistBits)];
                if ((d & kKindMask) == kSub) {
                    b.drop(kDistBits);
                    d = dst[(d >> 16) + b.peek(static_cast<int>((d >> 8) & 15u))];
                }
                if ((d & kKindMask) != kBase || static_cast<int>(d & 255u) > b.cnt) return false;
                b.drop(static_cast<int>(d & 255u));
                const int xd = static_cast<int>((d >> 8) & 15u);
                const uint32_t dist = (d >> 16) + b.take(xd);
                if (b.cnt < 0) return false;
                if (dist > static_cast<size_t>(o - out) || len > static_cast<size_t>(oend - o)) return false;
                const uint8_t* src = o - dist;
                if (dist >= 8 && static_cast<size_t>(oend - o) >= len + 8) {
                    uint8_t* const stop = o + len;
                    do { memcpy(o, src, 8); o += 8; src += 8; } while (o < stop);
                    o = stop;
                } else if (dist == 1) {                      // run of one byte (BAM: absent qualities, homopolymers)
                    memset(o, *src, len);
                    o += len;
                } else {
                    while (len--) *o++ = *src++;
                }
            }
        } else {
            return false;
        }
        if (final_block) break;
    }
    return o == oend;
}

}  // namespace msinflate
