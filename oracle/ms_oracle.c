/*
 * ms_oracle.c -- CPU RESTATEMENT (test oracle) of juliet's pileup / codon test /
 * phasing and fuse's consensus.  See ms_oracle.h: test infrastructure only,
 * PARITY UNPINNED (the reference ships documentation only).
 *
 * Every function cites the reference text it restates (path:line under
 * /root/reference) and, where the docs are silent, the SURVEY.md App. B
 * "restatement choice" it implements.  Written to be obviously correct, not
 * fast: one byte per column, plain loops, long-double log-gamma for Fisher.
 */
#define _GNU_SOURCE
#include "ms_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- format bridge -------------------------------------------------------
 * The product stores a read as ceil(L/32) blocks of four u32 bit-planes
 * (include/minorseq_b200.h): plane p of block b is word 4*b+p, bit j is column
 * 32*b+j; planes 0..2 are the state bits, plane 3 the insertion flag.        */
void mso_unpack_planar_mt(const uint32_t *packed, int64_t R, int32_t L, uint8_t *states, int nthreads)
{
    int32_t nblk = (L + 31) / 32;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 1 ? nthreads : 1) schedule(static)
#endif
    for (int64_t r = 0; r < R; ++r) {
        const uint32_t *row = packed + (size_t)r * 4 * nblk;
        uint8_t *out = states + (size_t)r * L;
        for (int32_t j = 0; j < L; ++j) {
            const uint32_t *w = row + 4 * (j >> 5);
            int sh = j & 31;
            out[j] = (uint8_t)(((w[0] >> sh) & 1) | (((w[1] >> sh) & 1) << 1) |
                               (((w[2] >> sh) & 1) << 2) | (((w[3] >> sh) & 1) << 3));
        }
    }
}

void mso_unpack_planar(const uint32_t *packed, int64_t R, int32_t L, uint8_t *states)
{
    mso_unpack_planar_mt(packed, R, L, states, 1);
}

/* ---- a4 + a5: per-column and per-codon histograms --------------------------
 * doc/JULIET.md:99-100 ("counts of the multiple-sequence alignment"), screenshot
 * juliet_hiv-context.png: columns A C G T - N; every row sums to the number of
 * reads spanning the column (SURVEY C-1).  A QV-filtered base is 'N' and "does
 * not count towards the coverage" (doc/JULIET.md:256-259): codon coverage is the
 * number of reads whose three states are all A/C/G/T.         */
static void pileup_range(const uint8_t *states, int64_t r0, int64_t r1, int32_t L,
                         const uint8_t *start_mask, uint32_t *col, uint32_t *codon)
{
    for (int64_t r = r0; r < r1; ++r) {
        const uint8_t *s = states + (size_t)r * L;
        for (int32_t j = 0; j < L; ++j) {
            int st = s[j] & 7;
            if (st <= MSO_N) {
                col[(size_t)j * 8 + st]++;
                col[(size_t)j * 8 + MSO_COL_COV]++;
                if (s[j] & 8) col[(size_t)j * 8 + MSO_COL_INS]++;
            }
        }
        if (!codon) continue;
        for (int32_t j = 0; j + 2 < L; ++j) {
            if (start_mask && !start_mask[j]) continue;
            int a = s[j] & 7, b = s[j + 1] & 7, c = s[j + 2] & 7;
            if (a > 3 || b > 3 || c > 3) continue;
            codon[(size_t)j * 64 + 16 * a + 4 * b + c]++;
        }
    }
}

void mso_pileup(const uint8_t *states, int64_t R, int32_t L, const uint8_t *start_mask,
                uint32_t *col, uint32_t *codon, int nthreads)
{
    memset(col, 0, (size_t)L * 8 * sizeof(uint32_t));
    if (codon) memset(codon, 0, (size_t)L * 64 * sizeof(uint32_t));
#ifdef _OPENMP
    if (nthreads > 1) {
#pragma omp parallel num_threads(nthreads)
        {
            int t = omp_get_thread_num(), nt = omp_get_num_threads();
            uint32_t *c1 = calloc((size_t)L * 8, sizeof(uint32_t));
            uint32_t *c2 = codon ? calloc((size_t)L * 64, sizeof(uint32_t)) : NULL;
            pileup_range(states, R * t / nt, R * (t + 1) / nt, L, start_mask, c1, c2);
#pragma omp critical
            {
                for (size_t i = 0; i < (size_t)L * 8; ++i) col[i] += c1[i];
                if (codon) for (size_t i = 0; i < (size_t)L * 64; ++i) codon[i] += c2[i];
            }
            free(c1); free(c2);
        }
        return;
    }
#endif
    (void)nthreads;
    pileup_range(states, 0, R, L, start_mask, col, codon);
}

/* ---- a8: Fisher's exact test ---------------------------------------------
 * doc/JULIET.md:42 "Bonferroni-corrected Fisher's Exact test".  Sidedness and
 * table are restatement choice U2: one-sided "greater" on [[a,b],[c,d]].
 * P(X >= a), X ~ Hypergeometric(white = a+c, black = b+d, draws = a+b).
 * Long-double log-gamma keeps the relative error near 1e-11 at n = 1e6.      */
static long double lchoose_ld(long double n, long double k)
{
    return lgammal(n + 1.0L) - lgammal(k + 1.0L) - lgammal(n - k + 1.0L);
}

double mso_fisher_greater(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    long double white = (long double)a + c, black = (long double)b + d, draws = (long double)a + b;
    long double xmax = white < draws ? white : draws;
    long double x = a;
    long double lp = lchoose_ld(white, x) + lchoose_ld(black, draws - x) - lchoose_ld(white + black, draws);
    long double term = expl(lp), sum = term;
    while (x < xmax && term > 0.0L) {
        /* pmf(x+1)/pmf(x) */
        term *= (white - x) * (draws - x) / ((x + 1.0L) * (black - draws + x + 1.0L));
        sum += term;
        x += 1.0L;
    }
    if (sum > 1.0L) sum = 1.0L;
    return (double)sum;
}

/* ---- a7: error-model expectation -------------------------------------------
 * doc/JULIET.md:40-41 "number of expected mutations at a given position"; the
 * constants are not documented (doc/JULIET.md:221-225 only says chemistry-
 * dependent with a permissive fallback).  Restatement choice U1: independent
 * per-base errors, P = prod(match if equal else substitution/3),
 * match = 1 - substitution - deletion.                                         */
double mso_codon_error_prob(int ref_codon, int codon, double sub_rate, double del_rate)
{
    double match = 1.0 - sub_rate - del_rate, mis = sub_rate / 3.0;
    int nm = 0;
    for (int i = 0; i < 3; ++i)
        if (((ref_codon >> (2 * i)) & 3) != ((codon >> (2 * i)) & 3)) nm++;
    /* fixed evaluation order so every implementation can reproduce the double exactly */
    double p = 1.0;
    for (int i = 0; i < 3 - nm; ++i) p *= match;
    for (int i = 0; i < nm; ++i) p *= mis;
    return p;
}

static int base_code(char ch)
{
    switch (ch) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
    }
}

/* ---- a6-a9: per-gene, per-codon minor-variant test ---------------------------
 * doc/JULIET.md:38-42 (model), :133-136 (reference sequence vs major codon;
 * 1-based [begin,end) in alignment space), :261-264 (each gene separately,
 * overlaps allowed), :270-271 (--region), :342-354 (--min-perc/--max-perc).
 * Restatement choices: U2 table [[k,n-k],[e,n-e]] with e = ceil(n*P) clipped
 * to n; U3 Bonferroni factor = codon positions tested in this gene; U6 lowest
 * code wins majority ties.  Synonymous codons are reported (screenshot
 * juliet_hiv-own.png); all 64 codons translate (stop = X, F19).    */
static int cmp_variant(const void *pa, const void *pb)
{
    const mso_variant *a = pa, *b = pb;
    if (a->gene != b->gene) return a->gene < b->gene ? -1 : 1;
    if (a->col != b->col) return a->col < b->col ? -1 : 1;
    return (a->codon > b->codon) - (a->codon < b->codon);
}

int64_t mso_call(const uint32_t *codon, int32_t L, const mso_gene *genes, int32_t ngenes,
                 const char *refseq, const mso_call_params *prm, mso_variant *out, int64_t cap)
{
    int64_t nout = 0;
    int32_t lo = 0, hi = L;
    if (prm->region_end > prm->region_begin) {
        lo = prm->region_begin - 1; hi = prm->region_end - 1;
        if (lo < 0) lo = 0;
        if (hi > L) hi = L;
    }
    for (int32_t g = 0; g < ngenes; ++g) {
        int32_t gb = genes[g].begin - 1, ge = genes[g].end - 1; /* 0-based [gb,ge) */
        if (ge > L) ge = L;
        /* count tested positions first: Bonferroni factor (U3) */
        uint32_t ntests = 0;
        for (int32_t s = gb; s + 3 <= ge; s += 3)
            if (s >= lo && s + 3 <= hi && s >= 0) ntests++;
        for (int32_t s = gb, ci = 0; s + 3 <= ge; s += 3, ++ci) {
            if (!(s >= lo && s + 3 <= hi && s >= 0)) continue;
            const uint32_t *h = codon + (size_t)s * 64;
            uint64_t n = 0;
            for (int c = 0; c < 64; ++c) n += h[c];
            if (n == 0) continue;
            int ref = -1;
            if (refseq) {
                int b0 = base_code(refseq[s]), b1 = base_code(refseq[s + 1]), b2 = base_code(refseq[s + 2]);
                if (b0 >= 0 && b1 >= 0 && b2 >= 0) ref = 16 * b0 + 4 * b1 + b2;
            }
            if (ref < 0) { /* doc/JULIET.md:133-134 "otherwise it will be tested against the major codon" */
                ref = 0;
                for (int c = 1; c < 64; ++c) if (h[c] > h[ref]) ref = c;
            }
            for (int c = 0; c < 64; ++c) {
                uint32_t k = h[c];
                if (c == ref || k == 0) continue;
                double P = mso_codon_error_prob(ref, c, prm->substitution_rate, prm->deletion_rate);
                double ex = ceil((double)n * P);
                uint32_t e = ex >= (double)n ? (uint32_t)n : (uint32_t)ex;
                double p = mso_fisher_greater(k, (uint32_t)(n - k), e, (uint32_t)(n - e));
                if (!(p * (double)ntests < prm->alpha)) continue;
                double perc = 100.0 * (double)k / (double)n;
                if (prm->min_perc >= 0 && !(perc > prm->min_perc)) continue;
                if (prm->max_perc >= 0 && !(perc < prm->max_perc)) continue;
                if (nout < cap) {
                    mso_variant *v = &out[nout];
                    v->gene = g; v->codon_index = ci; v->col = s; v->ref_codon = ref; v->codon = c;
                    v->count = k; v->coverage = (uint32_t)n; v->expected = e; v->ntests = ntests;
                    v->pvalue = p;
                }
                nout++;
            }
        }
    }
    qsort(out, (size_t)(nout < cap ? nout : cap), sizeof(mso_variant), cmp_variant);
    return nout;
}

/* ---- a11: per-read variant presence and damage flags -------------------------
 * doc/JULIET.md:194-203 (which variants co-occur on a read), :278-288 (a read
 * with a deletion in any identified variant codon cannot be assigned), :372-381
 * and screenshot juliet_haplotype-tooltip.png (marginals: gaps, heteroduplexes
 * = 'N' codons, partial = read does not span every variant codon). */
void mso_phase_bits(const uint8_t *states, int64_t R, int32_t L,
                    const int32_t *var_col, const int32_t *var_codon, int32_t V,
                    uint32_t *bits, uint8_t *flags)
{
    mso_phase_bits_mt(states, R, L, var_col, var_codon, V, bits, flags, 1);
}

void mso_phase_bits_mt(const uint8_t *states, int64_t R, int32_t L,
                       const int32_t *var_col, const int32_t *var_codon, int32_t V,
                       uint32_t *bits, uint8_t *flags, int nthreads)
{
    int32_t nw = (V + 31) / 32;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 1 ? nthreads : 1) schedule(static)
#endif
    for (int64_t r = 0; r < R; ++r) {
        const uint8_t *s = states + (size_t)r * L;
        uint32_t *bw = bits + (size_t)r * nw;
        uint8_t f = 0;
        for (int32_t w = 0; w < nw; ++w) bw[w] = 0;
        for (int32_t v = 0; v < V; ++v) {
            int32_t j = var_col[v];
            int st[3];
            for (int i = 0; i < 3; ++i) st[i] = (j + i < L) ? (s[j + i] & 7) : MSO_UNCOV;
            int clean = 1;
            for (int i = 0; i < 3; ++i) {
                if (st[i] == MSO_DEL) { f |= MSO_FLAG_GAP; clean = 0; }
                else if (st[i] == MSO_N) { f |= MSO_FLAG_HET; clean = 0; }
                else if (st[i] > 3) { f |= MSO_FLAG_PARTIAL; clean = 0; }
            }
            if (clean && 16 * st[0] + 4 * st[1] + st[2] == var_codon[v]) bw[v >> 5] |= 1u << (v & 31);
        }
        flags[r] = f;
    }
}

/* ---- a12: haplotype grouping --------------------------------------------------
 * doc/JULIET.md:198-211, :253-254 (>= 10 reads to report), :372-381 (reported /
 * insufficient / unsuitable partition the reads).  Order: descending count
 * (screenshot juliet_hiv-phasing.png 92.5, 1.2, 1.2, 1 ...), ties by ascending
 * pattern words, word 0 first (restatement choice U12).                        */
typedef struct { const uint32_t *bits; int32_t nw; } sort_ctx;
static int cmp_read_pattern(const void *pa, const void *pb, void *vctx)
{
    const sort_ctx *ctx = vctx;
    const uint32_t *a = ctx->bits + (size_t)(*(const int64_t *)pa) * ctx->nw;
    const uint32_t *b = ctx->bits + (size_t)(*(const int64_t *)pb) * ctx->nw;
    for (int32_t w = 0; w < ctx->nw; ++w)
        if (a[w] != b[w]) return a[w] < b[w] ? -1 : 1;
    return 0;
}
typedef struct { int64_t first, count; } grp;
static int cmp_grp(const void *pa, const void *pb, void *vctx)
{
    const grp *a = pa, *b = pb;
    if (a->count != b->count) return a->count > b->count ? -1 : 1;
    return cmp_read_pattern(&a->first, &b->first, vctx);
}

int64_t mso_phase_group(const uint32_t *bits, const uint8_t *flags, int64_t R, int32_t V,
                        int32_t min_reads, int32_t *hap_id, uint32_t *patterns, uint64_t *counts,
                        int64_t cap, int64_t *nreported, mso_phase_counters *ctr)
{
    int32_t nw = (V + 31) / 32;
    sort_ctx ctx = { bits, nw };
    mso_phase_counters c = { 0, 0, 0, 0, 0, 0 };
    int64_t *idx = malloc(sizeof(int64_t) * (size_t)(R > 0 ? R : 1));
    int64_t n = 0;
    for (int64_t r = 0; r < R; ++r) {
        hap_id[r] = -1;
        if (flags[r]) {
            c.damaged++;
            if (flags[r] & MSO_FLAG_GAP) c.gaps++;
            if (flags[r] & MSO_FLAG_HET) c.heteroduplex++;
            if (flags[r] & MSO_FLAG_PARTIAL) c.partial++;
        } else idx[n++] = r;
    }
    qsort_r(idx, (size_t)n, sizeof(int64_t), cmp_read_pattern, &ctx);
    grp *g = malloc(sizeof(grp) * (size_t)(n > 0 ? n : 1));
    int64_t H = 0;
    for (int64_t i = 0; i < n;) {
        int64_t j = i + 1;
        while (j < n && cmp_read_pattern(&idx[i], &idx[j], &ctx) == 0) ++j;
        g[H].first = idx[i]; g[H].count = j - i; ++H;
        i = j;
    }
    qsort_r(g, (size_t)H, sizeof(grp), cmp_grp, &ctx);
    /* rank lookup: walk the sorted reads again */
    int64_t nrep = 0;
    for (int64_t h = 0; h < H; ++h) {
        if (g[h].count >= min_reads) { nrep++; c.reported += (uint64_t)g[h].count; }
        else c.insufficient += (uint64_t)g[h].count;
        if (h < cap) {
            memcpy(patterns + (size_t)h * nw, bits + (size_t)g[h].first * nw, sizeof(uint32_t) * (size_t)nw);
            counts[h] = (uint64_t)g[h].count;
        }
    }
    for (int64_t i = 0; i < n; ++i) {
        /* linear probe over groups is O(n*H); fine for an oracle, but use bsearch-free two-pointer:
           groups are re-ordered, so map by comparing against each group's first read lazily */
        hap_id[idx[i]] = -2;
    }
    for (int64_t h = 0; h < H; ++h) {
        /* locate the group's run in idx via binary search on the pattern */
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            int64_t mid = (lo + hi) / 2;
            if (cmp_read_pattern(&idx[mid], &g[h].first, &ctx) < 0) lo = mid + 1; else hi = mid;
        }
        for (int64_t i = lo; i < lo + g[h].count; ++i) hap_id[idx[i]] = (int32_t)h;
    }
    free(idx); free(g);
    if (nreported) *nreported = nrep;
    if (ctr) *ctr = c;
    return H;
}

void mso_haplotype_name(int64_t i, char *buf)
{
    /* doc/JULIET.md:198 "[A-Z]{1}[a-z]?" */
    if (i < 26) { buf[0] = (char)('A' + i); buf[1] = 0; return; }
    i -= 26;
    buf[0] = (char)('A' + (i / 26) % 26); buf[1] = (char)('a' + i % 26); buf[2] = 0;
}

/* ---- a13: co-occurrence (north_star addition; doc/JULIET.md:201 "variants that co-occur") */
void mso_cooccurrence(const uint32_t *bits, int64_t R, int32_t V, int32_t *C)
{
    int32_t nw = (V + 31) / 32;
    memset(C, 0, sizeof(int32_t) * (size_t)V * V);
    int32_t *set = malloc(sizeof(int32_t) * (size_t)(V > 0 ? V : 1));
    for (int64_t r = 0; r < R; ++r) {
        const uint32_t *bw = bits + (size_t)r * nw;
        int32_t ns = 0;
        for (int32_t v = 0; v < V; ++v) if ((bw[v >> 5] >> (v & 31)) & 1) set[ns++] = v;
        for (int32_t i = 0; i < ns; ++i)
            for (int32_t j = 0; j < ns; ++j) C[(size_t)set[i] * V + set[j]]++;
    }
    free(set);
}

/* ---- a14 + a15: fuse consensus --------------------------------------------------
 * doc/FUSE.md:17-20: "creation of a high-quality consensus sequence", "includes
 * in-frame insertions with a certain distance to each other", "Major deletions
 * are being removed".  Restatement choices: U6 majority over A,C,G,T,- with the
 * lowest code winning ties and 'N' not voting; U8 columns with fewer than
 * min_coverage voting reads emit nothing; U7 an inserted string after column j
 * is accepted iff its length is a multiple of 3, its support exceeds
 * ins_fraction * voting reads at j, and it is >= ins_distance columns past the
 * previously accepted insertion (greedy, left to right).                       */
typedef struct { int32_t col; int32_t len; const char *s; } ins_ev;
static int cmp_ins(const void *pa, const void *pb)
{
    const ins_ev *a = pa, *b = pb;
    if (a->col != b->col) return a->col < b->col ? -1 : 1;
    if (a->len != b->len) return a->len < b->len ? -1 : 1;
    int c = memcmp(a->s, b->s, (size_t)a->len);
    return c < 0 ? -1 : (c > 0);
}

int64_t mso_fuse(const uint32_t *col, int32_t L, const int32_t *ins_col, const int64_t *ins_off,
                 const int32_t *ins_len, int64_t nins, const char *ins_pool,
                 const mso_fuse_params *prm, char *seq, int64_t cap)
{
    static const char base[4] = { 'A', 'C', 'G', 'T' };
    ins_ev *ev = malloc(sizeof(ins_ev) * (size_t)(nins > 0 ? nins : 1));
    for (int64_t i = 0; i < nins; ++i) {
        ev[i].col = ins_col[i]; ev[i].len = ins_len[i]; ev[i].s = ins_pool + ins_off[i];
    }
    qsort(ev, (size_t)nins, sizeof(ins_ev), cmp_ins);
    /* accepted insertion per column: index into ev or -1 */
    int64_t *acc = malloc(sizeof(int64_t) * (size_t)L);
    for (int32_t j = 0; j < L; ++j) acc[j] = -1;
    int32_t last = -1;
    for (int64_t i = 0; i < nins;) {
        int64_t j = i + 1;
        while (j < nins && cmp_ins(&ev[i], &ev[j]) == 0) ++j;
        int32_t c = ev[i].col;
        if (c >= 0 && c < L && acc[c] < 0 && ev[i].len > 0 && ev[i].len % 3 == 0) {
            const uint32_t *h = col + (size_t)c * 8;
            uint64_t votes = (uint64_t)h[0] + h[1] + h[2] + h[3] + h[4];
            if ((double)(j - i) > prm->ins_fraction * (double)votes &&
                (last < 0 || c - last >= prm->ins_distance)) {
                acc[c] = i; last = c;
            }
        }
        i = j;
    }
    int64_t n = 0;
    for (int32_t j = 0; j < L; ++j) {
        const uint32_t *h = col + (size_t)j * 8;
        uint64_t votes = (uint64_t)h[0] + h[1] + h[2] + h[3] + h[4];
        if (votes >= (uint64_t)prm->min_coverage && votes > 0) {
            int best = 0;
            for (int s = 1; s < 5; ++s) if (h[s] > h[best]) best = s;
            if (best < 4) { if (n < cap) seq[n] = base[best]; n++; }
        }
        if (acc[j] >= 0)
            for (int32_t k = 0; k < ev[acc[j]].len; ++k) { if (n < cap) seq[n] = ev[acc[j]].s[k]; n++; }
    }
    free(ev); free(acc);
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * cleric restatement (doc/CLERIC.md:19-23,41-44), choices U13 in the header.
 * ------------------------------------------------------------------------------------------------ */
enum { NW_MATCH = 2, NW_MISMATCH = -3, NW_GAP = -4 };

int64_t mso_nw_align(const char *a, int32_t la, const char *b, int32_t lb, char *ops, int64_t cap, int64_t *score)
{
    if (la < 0 || lb < 0) return -1;
    const size_t W = (size_t)lb + 1;
    int32_t *prev = (int32_t *)malloc(W * sizeof(int32_t)), *cur = (int32_t *)malloc(W * sizeof(int32_t));
    uint8_t *dir = (uint8_t *)malloc(((size_t)la + 1) * W);      /* 0 diagonal, 1 up (consume A), 2 left (consume B) */
    if (!prev || !cur || !dir) { free(prev); free(cur); free(dir); return -1; }
    for (int32_t j = 0; j <= lb; ++j) { prev[j] = j * NW_GAP; dir[j] = 2; }
    for (int32_t i = 1; i <= la; ++i) {
        cur[0] = i * NW_GAP;
        dir[(size_t)i * W] = 1;
        for (int32_t j = 1; j <= lb; ++j) {
            const int32_t d = prev[j - 1] + (a[i - 1] == b[j - 1] ? NW_MATCH : NW_MISMATCH);
            const int32_t u = prev[j] + NW_GAP, l = cur[j - 1] + NW_GAP;
            int32_t best = d; uint8_t k = 0;
            if (u > best) { best = u; k = 1; }
            if (l > best) { best = l; k = 2; }
            cur[j] = best;
            dir[(size_t)i * W + j] = k;
        }
        int32_t *t = prev; prev = cur; cur = t;
    }
    if (score) *score = prev[lb];
    /* trace back, then reverse */
    int64_t n = 0;
    int32_t i = la, j = lb;
    int64_t rc = 0;
    while (i > 0 || j > 0) {
        const uint8_t k = dir[(size_t)i * W + j];
        if (n >= cap) { rc = -1; break; }
        if (k == 0) { ops[n++] = 'M'; --i; --j; }
        else if (k == 1) { ops[n++] = 'D'; --i; }
        else { ops[n++] = 'I'; --j; }
    }
    free(prev); free(cur); free(dir);
    if (rc < 0) return -1;
    for (int64_t x = 0, y = n - 1; x < y; ++x, --y) { const char t = ops[x]; ops[x] = ops[y]; ops[y] = t; }
    return n;
}

/* append one op of length len, merging with the previous one */
static int push_op(char *op, int32_t *len, int64_t cap, int64_t *n, char c, int32_t l)
{
    if (l <= 0) return 0;
    if (*n > 0 && op[*n - 1] == c) { len[*n - 1] += l; return 0; }
    if (*n >= cap) return -1;
    op[*n] = c; len[*n] = l; ++*n;
    return 0;
}

int64_t mso_project_read(const char *ops, int64_t nops, const char *b, int32_t lb,
                         int32_t pos, const char *cig_op, const int32_t *cig_len, int32_t ncig,
                         const char *seq, int32_t lseq,
                         char *new_op, int32_t *new_len, int64_t cap, int32_t *new_pos)
{
    /* per-op coordinates of the path: ai/bj = A and B columns consumed BEFORE op x */
    int32_t la = 0;
    for (int64_t x = 0; x < nops; ++x) la += ops[x] != 'I';
    /* first path op that consumes A column `pos` */
    if (pos < 0 || pos > la) return -1;
    /* middle part of the CIGAR: between the leading and trailing clips */
    int32_t c0 = 0, c1 = ncig;
    while (c0 < c1 && (cig_op[c0] == 'S' || cig_op[c0] == 'H')) ++c0;
    while (c1 > c0 && (cig_op[c1 - 1] == 'S' || cig_op[c1 - 1] == 'H')) --c1;
    /* raw projected ops (1 column each, merged on the fly) in a scratch list */
    const int64_t scap = (int64_t)lseq + nops + 8;
    char *rop = (char *)malloc((size_t)scap);
    int32_t *rlen = (int32_t *)malloc((size_t)scap * sizeof(int32_t));
    int32_t *rb = (int32_t *)malloc((size_t)scap * sizeof(int32_t));     /* B column of the op's first base (ref-consuming ops) */
    if (!rop || !rlen || !rb) { free(rop); free(rlen); free(rb); return -1; }
    int64_t rn = 0;
    int rc = 0;
#define RAW(c, l, bj) do { if ((l) > 0) { if (rn > 0 && rop[rn - 1] == (c)) rlen[rn - 1] += (l); \
        else if (rn >= scap) rc = -1; else { rop[rn] = (c); rlen[rn] = (l); rb[rn] = (bj); ++rn; } } } while (0)
    int64_t x = 0;
    int32_t ai = 0, bj = 0, q = 0;
    /* skip the path up to A column pos (B columns before it do not belong to the read) */
    while (x < nops && !(ops[x] != 'I' && ai == pos)) { if (ops[x] != 'I') ++ai; if (ops[x] != 'D') ++bj; ++x; }
    for (int32_t c = 0; c < c0; ++c) if (cig_op[c] == 'S') q += cig_len[c];
    int started = 0;
    for (int32_t c = c0; c < c1 && rc == 0; ++c) {
        const char o = cig_op[c];
        const int32_t l = cig_len[c];
        if (o == 'I') { RAW('I', l, -1); q += l; continue; }
        if (o != '=' && o != 'X' && o != 'D') { rc = -1; break; }
        for (int32_t t = 0; t < l && rc == 0; ++t) {
            /* B-only columns in front of the next A column lie inside the read once it has started */
            while (x < nops && ops[x] == 'I') { if (started) RAW('D', 1, bj); ++bj; ++x; }
            if (x >= nops) { rc = -1; break; }              /* the read runs past the end of A */
            if (ops[x] == 'M') {
                if (o == 'D') RAW('D', 1, bj);
                else {
                    if (q >= lseq || bj >= lb) { rc = -1; break; }
                    RAW(seq[q] == b[bj] ? '=' : 'X', 1, bj);
                    ++q;
                }
                ++bj;
            } else {                                        /* 'D': this A column has no partner on B */
                if (o != 'D') { RAW('I', 1, -1); ++q; }
            }
            ++ai; ++x;
            started = 1;
        }
    }
#undef RAW
    if (rc == 0) {
        for (int32_t c = c1; c < ncig; ++c) if (cig_op[c] == 'S') q += cig_len[c];
        if (q != lseq) rc = -1;                             /* CIGAR and sequence disagree */
    }
    int64_t n = 0;
    if (rc == 0) {
        /* trim: leading / trailing D dropped, leading / trailing I become clips */
        int64_t lo = 0, hi = rn;
        int32_t lead_clip = 0, tail_clip = 0;
        while (lo < hi && (rop[lo] == 'D' || rop[lo] == 'I')) { if (rop[lo] == 'I') lead_clip += rlen[lo]; ++lo; }
        while (hi > lo && (rop[hi - 1] == 'D' || rop[hi - 1] == 'I')) { if (rop[hi - 1] == 'I') tail_clip += rlen[hi - 1]; --hi; }
        if (lo == hi) { n = 0; }
        else {
            *new_pos = rb[lo];
            /* leading clips: hard clips first, then soft */
            for (int32_t c = 0; c < c0 && rc == 0; ++c) if (cig_op[c] == 'H') rc = push_op(new_op, new_len, cap, &n, 'H', cig_len[c]);
            int32_t s0 = lead_clip;
            for (int32_t c = 0; c < c0; ++c) if (cig_op[c] == 'S') s0 += cig_len[c];
            if (rc == 0) rc = push_op(new_op, new_len, cap, &n, 'S', s0);
            for (int64_t k = lo; k < hi && rc == 0; ++k) rc = push_op(new_op, new_len, cap, &n, rop[k], rlen[k]);
            int32_t s1 = tail_clip;
            for (int32_t c = c1; c < ncig; ++c) if (cig_op[c] == 'S') s1 += cig_len[c];
            if (rc == 0) rc = push_op(new_op, new_len, cap, &n, 'S', s1);
            for (int32_t c = c1; c < ncig && rc == 0; ++c) if (cig_op[c] == 'H') rc = push_op(new_op, new_len, cap, &n, 'H', cig_len[c]);
            if (rc != 0) { free(rop); free(rlen); free(rb); return -2; }
        }
    }
    free(rop); free(rlen); free(rb);
    return rc == 0 ? n : -1;
}

/* ---- workload generator: C twin of minorseq_b200/synth.py (see ms_oracle.h) ---------------------------------- */
static uint64_t synth_mix64(uint64_t x)
{
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}

void mso_synth_states(const mso_synth_params *p, const uint8_t *strain_base, const uint32_t *thr_del,
                      const uint32_t *strain_cum, int64_t read0, int64_t R, uint8_t *states, int nthreads)
{
    const uint64_t K1 = 0x9E3779B97F4A7C15ULL, K2 = 0xD1B54A32D192ED03ULL;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 1 ? nthreads : 1) schedule(static)
#endif
    for (int64_t i = 0; i < R; ++i) {
        const uint64_t r = (uint64_t)(read0 + i);
        const uint64_t y = synth_mix64(p->seed + r * K1);
        const uint32_t us = (uint32_t)y;
        int32_t strain = p->nstrains - 1;
        for (int32_t s = 0; s < p->nstrains; ++s)
            if (us < strain_cum[s]) { strain = s; break; }
        int32_t begin = 0, end = p->L;
        if (((y >> 32) & 0xffffu) < p->thr_trunc16) {
            const uint32_t z = (uint32_t)(y >> 48);
            const int32_t amount = (int32_t)(((uint64_t)(z >> 1) * (uint64_t)(p->L / 2)) >> 15);
            if (z & 1u) begin = amount; else end = p->L - amount;
        }
        const uint8_t *sb = strain_base + (size_t)strain * p->L;
        uint8_t *out = states + (size_t)i * p->L;
        for (int32_t c = 0; c < p->L; ++c) {
            uint32_t st = 7, ins = 0;
            if (c >= begin && c < end) {
                const uint64_t x = synth_mix64(p->seed + r * K1 + (uint64_t)(c + 1) * K2);
                const uint64_t u = x & 0xffffffffULL;
                const uint64_t tN = p->thr_N, tD = tN + thr_del[c], tS = tD + p->thr_sub;
                const uint32_t base = sb[c];
                if (u < tN) st = 5;
                else if (u < tD) st = 4;
                else if (u < tS) st = (base + 1u + (uint32_t)((x >> 32) % 3ULL)) & 3u;
                else st = base;
                ins = ((x >> 44) & 0xfffffULL) < p->thr_ins20 ? 1u : 0u;
            }
            out[c] = (uint8_t)(st == 7 ? 7 : (st | (ins << 3)));
        }
    }
}
