// format.cu -- host-side packed-format helpers of the C ABI (no GPU work).
//
// ms_expand_cigar replaces the per-record walk juliet and fuse do over an aligned
// BAM record (/root/reference/doc/JULIET.md:49-58: PacBio-compliant BAM, CIGAR 'M'
// forbidden, QV-filtered bases become 'N' :256-259; /root/reference/doc/FUSE.md:13-15).
#include <cstring>
#include "handle.h"

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
        uint32_t* row = packed + static_cast<size_t>(r) * 4 * nblk;
        for (int32_t b = 0; b < nblk; ++b) {
            uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0;
            for (int32_t j = 0; j < 32; ++j) {
                const int32_t c = b * 32 + j;
                const uint32_t v = c < L ? s[c] : 7u;
                if ((v & 7u) == 6u) return MS_ERR_FORMAT;
                p0 |= (v & 1u) << j;
                p1 |= ((v >> 1) & 1u) << j;
                p2 |= ((v >> 2) & 1u) << j;
                p3 |= ((v >> 3) & 1u) << j;
            }
            row[4 * b] = p0; row[4 * b + 1] = p1; row[4 * b + 2] = p2; row[4 * b + 3] = p3;
        }
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

static inline void set_col(uint32_t* row, int32_t c, uint32_t st) {
    uint32_t* w = row + 4 * (c >> 5);
    const uint32_t bit = 1u << (c & 31);
    // row starts as state 7 everywhere: clear the planes that are 0 in st
    if (!(st & 1u)) w[0] &= ~bit;
    if (!(st & 2u)) w[1] &= ~bit;
    if (!(st & 4u)) w[2] &= ~bit;
}

int ms_expand_cigar(const uint32_t* cigar, int32_t ncigar, int32_t pos, const char* seq, const uint8_t* qv_mask,
                    int32_t lseq, int32_t L, uint32_t* row, int32_t* ins_col, int64_t* ins_off, int32_t* ins_len,
                    int64_t ins_cap, int64_t* nins, char* ins_pool, int64_t pool_cap, int64_t* pool_used) {
    if (!cigar || !seq || !row || L <= 0 || ncigar < 0 || lseq < 0) return MS_ERR_ARG;
    const int32_t nblk = (L + 31) / 32;
    for (int32_t b = 0; b < nblk; ++b) {
        row[4 * b] = row[4 * b + 1] = row[4 * b + 2] = 0xffffffffu;  // not spanned
        row[4 * b + 3] = 0;
    }
    int32_t rc = pos;  // reference column
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
            for (int32_t i = 0; i < len; ++i, ++rc, ++qi) {
                if (rc < 0 || rc >= L) continue;
                uint32_t st;
                switch (seq[qi]) {
                case 'A': case 'a': st = MS_A; break;
                case 'C': case 'c': st = MS_C; break;
                case 'G': case 'g': st = MS_G; break;
                case 'T': case 't': st = MS_T; break;
                default: st = MS_N; break;
                }
                if (qv_mask && qv_mask[qi]) st = MS_N;
                set_col(row, rc, st);
            }
            break;
        }
        case 2:  // D
            for (int32_t i = 0; i < len; ++i, ++rc)
                if (rc >= 0 && rc < L) set_col(row, rc, MS_DEL);
            break;
        case 3:  // N (reference skip): columns stay "not spanned"
            rc += len;
            break;
        case 1: {  // I: attaches to the previous reference column
            if (qi + len > lseq) return MS_ERR_FORMAT;
            const int32_t c = rc - 1;
            if (c >= 0 && c < L && c >= pos) {
                row[4 * (c >> 5) + 3] |= 1u << (c & 31);
                if (ins_col && ins_off && ins_len && ins_pool) {
                    if (ni >= ins_cap || pu + len > pool_cap) return MS_ERR_CAPACITY;
                    ins_col[ni] = c; ins_off[ni] = pu; ins_len[ni] = len;
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
    if (nins) *nins = ni;
    if (pool_used) *pool_used = pu;
    return MS_OK;
}

}  // extern "C"
