// format.cu -- host-side packed-format helpers of the C ABI (no GPU work).
//
// ms_expand_cigar replaces the per-record walk juliet and fuse do over an aligned
// BAM record (/root/reference/doc/JULIET.md:49-58: PacBio-compliant BAM, CIGAR 'M'
// forbidden, QV-filtered bases become 'N' :256-259; /root/reference/doc/FUSE.md:13-15).
#include <algorithm>
#include <cstring>
#include <vector>
#include "handle.h"

namespace {

// 8 state bytes (one per column, bits 0..3) -> 8 bits of plane p, via the multiply-gather trick
inline uint32_t gather8(uint64_t x, int p) {
    return static_cast<uint32_t>((((x >> p) & 0x0101010101010101ULL) * 0x0102040810204080ULL) >> 56);
}

// L state bytes -> ceil(L/32) blocks of four planes; columns past L are state 7
void pack_row(const uint8_t* st, int32_t L, uint32_t* row) {
    const int32_t nblk = (L + 31) / 32;
    for (int32_t b = 0; b < nblk; ++b) {
        uint8_t tmp[32];
        const uint8_t* src = st + static_cast<size_t>(b) * 32;
        if (b * 32 + 32 > L) {
            memset(tmp, 7, 32);
            memcpy(tmp, src, static_cast<size_t>(L - b * 32));
            src = tmp;
        }
        uint32_t pl[4] = {0, 0, 0, 0};
        for (int g = 0; g < 4; ++g) {
            uint64_t x;
            memcpy(&x, src + 8 * g, 8);
            for (int p = 0; p < 4; ++p) pl[p] |= gather8(x, p) << (8 * g);
        }
        row[4 * b] = pl[0]; row[4 * b + 1] = pl[1]; row[4 * b + 2] = pl[2]; row[4 * b + 3] = pl[3];
    }
}

}  // namespace

extern "C" {

int32_t ms_row_words(int32_t L) { return L > 0 ? 4 * ((L + 31) / 32) : 0; }

// "Reads that are not primary or supplementary alignments, get ignored" (doc/JULIET.md:58):
// drop unmapped (0x4) and secondary (0x100) records; supplementary (0x800) stays.
int ms_read_admitted(uint32_t bam_flag) { return (bam_flag & (0x4u | 0x100u)) == 0 ? 1 : 0; }

int ms_pack_states(const uint8_t* states, int64_t R, int32_t L, uint32_t* packed) {
    if (!states || !packed || R < 0 || L <= 0) return MS_ERR_ARG;
    const int32_t nblk = (L + 31) / 32;
    for (int64_t r = 0; r < R; ++r) {
        const uint8_t* s = states + static_cast<size_t>(r) * L;
        for (int32_t c = 0; c < L; ++c)
            if ((s[c] & 7u) == 6u || s[c] > 15u) return MS_ERR_FORMAT;
        pack_row(s, L, packed + static_cast<size_t>(r) * 4 * nblk);
    }
    return MS_OK;
}

int ms_unpack_states(const uint32_t* packed, int64_t R, int32_t L, uint8_t* states) {
    if (!states || !packed || R < 0 || L <= 0) return MS_ERR_ARG;
    const int32_t nblk = (L + 31) / 32;
    for (int64_t r = 0; r < R; ++r) {
        const uint32_t* row = packed + static_cast<size_t>(r) * 4 * nblk;
        uint8_t* s = states + static_cast<size_t>(r) * L;
        for (int32_t c = 0; c < L; ++c) {
            const uint32_t* w = row + 4 * (c >> 5);
            const int sh = c & 31;
            s[c] = static_cast<uint8_t>(((w[0] >> sh) & 1u) | (((w[1] >> sh) & 1u) << 1) | (((w[2] >> sh) & 1u) << 2) |
                                        (((w[3] >> sh) & 1u) << 3));
        }
    }
    return MS_OK;
}

int ms_expand_cigar(const uint32_t* cigar, int32_t ncigar, int32_t pos, const char* seq, const uint8_t* qv_mask,
                    int32_t lseq, int32_t L, uint32_t* row, int32_t* ins_col, int64_t* ins_off, int32_t* ins_len,
                    int64_t ins_cap, int64_t* nins, char* ins_pool, int64_t pool_cap, int64_t* pool_used) {
    if (!cigar || !seq || !row || L <= 0 || ncigar < 0 || lseq < 0) return MS_ERR_ARG;
    static const struct Lut { uint8_t v[256]; Lut() { memset(v, MS_N, 256); v['A'] = v['a'] = MS_A; v['C'] = v['c'] = MS_C; v['G'] = v['g'] = MS_G; v['T'] = v['t'] = MS_T; } } lut;
    thread_local std::vector<uint8_t> st_tls;   // one state byte per column (bit 3 = insertion follows), packed at the end
    st_tls.assign(static_cast<size_t>(L), 7);
    uint8_t* const st = st_tls.data();         // (a thread_local in a shared object costs a call per access)
    int64_t rc = pos;  // reference column
    int32_t qi = 0;    // query index
    int64_t ni = nins ? *nins : 0, pu = pool_used ? *pool_used : 0;
    for (int32_t k = 0; k < ncigar; ++k) {
        const uint32_t op = cigar[k] & 15u;
        const int32_t len = static_cast<int32_t>(cigar[k] >> 4);
        switch (op) {
        case 0:  // M
            return MS_ERR_FORMAT;  // "cigar M is forbidden" doc/JULIET.md:53
        case 7:    // =
        case 8: {  // X
            if (qi + len > lseq) return MS_ERR_FORMAT;
            const int64_t c0 = std::max<int64_t>(rc, 0), c1 = std::min<int64_t>(rc + len, L);
            const char* sq = seq + (qi - rc);          // query base of column c is sq[c]
            if (qv_mask) {                                // branch-free select: low-QV bases are not predictable
                const uint8_t* mk = qv_mask + (qi - rc);
                for (int64_t c = c0; c < c1; ++c) {
                    const uint8_t b = lut.v[static_cast<uint8_t>(sq[c])];
                    const uint8_t m = static_cast<uint8_t>(-static_cast<int8_t>(mk[c] != 0));   // 0x00 or 0xff
                    st[c] = static_cast<uint8_t>((b & ~m) | (MS_N & m));
                }
            } else {
                for (int64_t c = c0; c < c1; ++c) st[c] = lut.v[static_cast<uint8_t>(sq[c])];
            }
            rc += len; qi += len;
            break;
        }
        case 2: {  // D
            const int64_t c0 = std::max<int64_t>(rc, 0), c1 = std::min<int64_t>(rc + len, L);
            if (c1 > c0) memset(st + c0, MS_DEL, static_cast<size_t>(c1 - c0));
            rc += len;
            break;
        }
        case 3:  // N (reference skip): columns stay "not spanned"
            rc += len;
            break;
        case 1: {  // I: attaches to the previous reference column
            if (qi + len > lseq) return MS_ERR_FORMAT;
            const int64_t c = rc - 1;
            if (c >= 0 && c < L && c >= pos && (st[c] & 7u) != 7u) {
                st[c] |= 8;
                if (ins_col && ins_off && ins_len && ins_pool) {
                    if (ni >= ins_cap || pu + len > pool_cap) return MS_ERR_CAPACITY;
                    ins_col[ni] = static_cast<int32_t>(c); ins_off[ni] = pu; ins_len[ni] = len;
                    memcpy(ins_pool + pu, seq + qi, static_cast<size_t>(len));
                    pu += len; ++ni;
                }
            }
            qi += len;
            break;
        }
        case 4:  // S
            qi += len;
            break;
        case 5:  // H
        case 6:  // P
            break;
        default:
            return MS_ERR_FORMAT;
        }
    }
    pack_row(st, L, row);
    if (nins) *nins = ni;
    if (pool_used) *pool_used = pu;
    return MS_OK;
}

}  // extern "C"
