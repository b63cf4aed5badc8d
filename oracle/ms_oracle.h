/*
 * ms_oracle.h -- CPU RESTATEMENT of minorseq's juliet/fuse data-parallel core.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may build,
 * link or call anything in oracle/.  The product (minorseq_b200/) never does.
 *
 * PARITY UNPINNED: /root/reference ships documentation only (7 markdown files,
 * 21 screenshots; no source, tests or fixtures), so this file restates the
 * algorithm from doc/JULIET.md, doc/FUSE.md and doc/CLERIC.md.  Behaviour the docs do not pin
 * (U13 for cleric) is a named "restatement choice" below.  The only
 * pins are the screenshot-derived known answers of SURVEY.md App. C, checked in
 * tests/test_oracle_doc_kats.py.
 *
 * Data model (one byte per reference column, "alignment space",
 * doc/JULIET.md:134-136):
 *   bits 0..2  state: 0=A 1=C 2=G 3=T 4='-' (deletion) 5='N' (QV-filtered base,
 *              doc/JULIET.md:256-259) 7=column not spanned by the read
 *   bit  3     an insertion follows this column in this read
 */
#ifndef MS_ORACLE_H
#define MS_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { MSO_A = 0, MSO_C = 1, MSO_G = 2, MSO_T = 3, MSO_DEL = 4, MSO_N = 5, MSO_UNCOV = 7 };
enum { MSO_COL_INS = 6, MSO_COL_COV = 7 };            /* col[j][6]=insertion flags, col[j][7]=A+C+G+T+-+N */
enum { MSO_FLAG_GAP = 1, MSO_FLAG_HET = 2, MSO_FLAG_PARTIAL = 4 };

/* restatement choices: U1 error model, U3 alpha, F21 threshold, U7/U8 fuse */
#define MSO_DEFAULT_SUBSTITUTION_RATE 5e-4
#define MSO_DEFAULT_DELETION_RATE     3e-3
#define MSO_DEFAULT_ALPHA             0.01
#define MSO_DEFAULT_MIN_HAP_READS     10
#define MSO_DEFAULT_FUSE_MIN_COVERAGE 50
#define MSO_DEFAULT_FUSE_INS_FRACTION 0.5
#define MSO_DEFAULT_FUSE_INS_DISTANCE 20

typedef struct {
    int32_t begin, end;      /* 1-based, begin inclusive, end exclusive (doc/JULIET.md:134-136) */
} mso_gene;

typedef struct {
    double substitution_rate, deletion_rate, alpha;
    double min_perc, max_perc;     /* doc/JULIET.md:342-354; <0 = off */
    int32_t region_begin, region_end; /* 1-based [begin,end); 0,0 = off (doc/JULIET.md:270-271) */
} mso_call_params;

typedef struct {
    int32_t gene;        /* index into genes[] */
    int32_t codon_index; /* 0-based codon number inside the gene (AA position - 1) */
    int32_t col;         /* 0-based start column */
    int32_t ref_codon;   /* 0..63, 16*b0+4*b1+b2 */
    int32_t codon;       /* 0..63 */
    uint32_t count;      /* reads carrying this codon */
    uint32_t coverage;   /* reads with a clean (all A/C/G/T) codon here */
    uint32_t expected;   /* ceil(coverage * P(ref->codon)) */
    uint32_t ntests;     /* Bonferroni factor */
    double pvalue;       /* uncorrected one-sided Fisher p */
} mso_variant;

typedef struct {
    uint64_t reported, insufficient, damaged, gaps, heteroduplex, partial; /* doc/JULIET.md:372-381 */
} mso_phase_counters;

/* unpack the product's planar 4-bit rows (include/minorseq_b200.h) into one byte per column */
void mso_unpack_planar(const uint32_t *packed, int64_t R, int32_t L, uint8_t *states);
void mso_unpack_planar_mt(const uint32_t *packed, int64_t R, int32_t L, uint8_t *states, int nthreads);

/* a4 + a5 of SURVEY section 8a.  start_mask[j]!=0 marks columns where a codon of some gene begins. */
void mso_pileup(const uint8_t *states, int64_t R, int32_t L, const uint8_t *start_mask,
                uint32_t *col /* L*8 */, uint32_t *codon /* L*64 */, int nthreads);

/* a8: one-sided (greater) Fisher exact p for [[a,b],[c,d]] */
double mso_fisher_greater(uint32_t a, uint32_t b, uint32_t c, uint32_t d);
/* a7: P(ref codon -> codon) under the restated error model (U1) */
double mso_codon_error_prob(int ref_codon, int codon, double sub_rate, double del_rate);

/* a6-a9: returns number of called variants (sorted by gene, col, codon); writes up to cap */
int64_t mso_call(const uint32_t *codon /* L*64 */, int32_t L,
                 const mso_gene *genes, int32_t ngenes,
                 const char *refseq /* NULL or >= L chars */,
                 const mso_call_params *prm, mso_variant *out, int64_t cap);

/* a11: per read bit-vector (V bits in ceil(V/32) words, variant v -> word v/32 bit v%32) + flags */
void mso_phase_bits(const uint8_t *states, int64_t R, int32_t L,
                    const int32_t *var_col, const int32_t *var_codon, int32_t V,
                    uint32_t *bits /* R*ceil(V/32) */, uint8_t *flags /* R */);
void mso_phase_bits_mt(const uint8_t *states, int64_t R, int32_t L,
                       const int32_t *var_col, const int32_t *var_codon, int32_t V,
                       uint32_t *bits, uint8_t *flags, int nthreads);

/* a12: groups undamaged reads by identical bit-vector.  hap_id[r] = rank of the read's
 * haplotype in the output order (count desc, then ascending words) or -1 damaged.
 * Returns the number of distinct patterns H; patterns/counts hold up to cap entries.
 * Entries [0,nreported) have count >= min_reads. */
int64_t mso_phase_group(const uint32_t *bits, const uint8_t *flags, int64_t R, int32_t V,
                        int32_t min_reads, int32_t *hap_id,
                        uint32_t *patterns /* cap*ceil(V/32) */, uint64_t *counts /* cap */,
                        int64_t cap, int64_t *nreported, mso_phase_counters *ctr);

/* haplotype name for rank i (doc/JULIET.md:198): A..Z, Aa..Az, Ba.. ; buf >= 3 bytes */
void mso_haplotype_name(int64_t i, char *buf);

/* a13: C[v][w] = sum_r bit[r][v]&bit[r][w] */
void mso_cooccurrence(const uint32_t *bits, int64_t R, int32_t V, int32_t *C /* V*V */);

/* a14+a15: fuse consensus.  Insertion events: ins_read/ins_col/ins_off/ins_len index into ins_pool.
 * Returns consensus length written to seq (cap bytes, no NUL). */
typedef struct {
    int32_t min_coverage; double ins_fraction; int32_t ins_distance;
} mso_fuse_params;
int64_t mso_fuse(const uint32_t *col /* L*8 */, int32_t L,
                 const int32_t *ins_col, const int64_t *ins_off, const int32_t *ins_len,
                 int64_t nins, const char *ins_pool,
                 const mso_fuse_params *prm, char *seq, int64_t cap);

/* ---- cleric: "converting a given alignment to a different reference ... by aligning the original
 * and target reference sequences.  A transitive alignment is used to generate the new alignment"; "The alignment
 * step runs a Needleman-Wunsch; with NxM runtime" (/root/reference/doc/CLERIC.md:19-23,41-44).  Nothing in the
 * reference pins the scoring or the projection rules: restatement choice U13 --
 *   global alignment, match +2, mismatch -3, linear gap -4; cell = max(diagonal, up (consume A), left (consume B)),
 *   ties resolved in that order; the path is read back from (N,M) to (0,0).
 * ops (forward order): 'M' A and B advance, 'D' only A advances (base of A absent from B), 'I' only B advances.
 * Returns the number of ops written (<= la+lb), or -1 if cap is too small; *score = alignment score. */
int64_t mso_nw_align(const char *a, int32_t la, const char *b, int32_t lb, char *ops, int64_t cap, int64_t *score);

/* One read, aligned to A at 0-based `pos` with CIGAR ops `cig_op` (chars of "=XIDSH"; 'M', 'N', 'P' are refused, doc
 * CLERIC.md:14-15) and lengths `cig_len`, re-expressed against B through the A/B path `ops`:
 *   A column paired with a B column: read base -> '=' / 'X' against B, read deletion -> 'D'
 *   A column without a partner:      read base -> 'I' (extra relative to B), read deletion -> nothing
 *   B column without a partner strictly inside the read's span -> 'D'
 *   read insertions stay 'I'; clips stay; leading/trailing 'I' become 'S', leading/trailing 'D' are dropped;
 *   neighbouring equal ops are merged.
 * Writes up to cap new ops; returns their number, 0 if nothing of the read lands on B (the record becomes unmapped),
 * -1 on an unsupported CIGAR op or inconsistent input, -2 if cap is too small.  *new_pos = 0-based start on B. */
int64_t mso_project_read(const char *ops, int64_t nops, const char *b, int32_t lb,
                         int32_t pos, const char *cig_op, const int32_t *cig_len, int32_t ncig,
                         const char *seq, int32_t lseq,
                         char *new_op, int32_t *new_len, int64_t cap, int32_t *new_pos);

/* ---- workload generator (not a restatement of anything in the reference): C twin of minorseq_b200/synth.py and
 * csrc/synth.cu -- the same pure function of (seed, read, column) -- so that the CPU arm of bench.py can produce its own
 * reads at full size without the GPU.  tests/test_oracle_properties.py checks it against the numpy generator.        */
typedef struct {
    uint64_t seed;
    int32_t L, nstrains;
    uint32_t thr_N, thr_sub, thr_ins20, thr_trunc16;
} mso_synth_params;
void mso_synth_states(const mso_synth_params *p, const uint8_t *strain_base, const uint32_t *thr_del,
                      const uint32_t *strain_cum, int64_t read0, int64_t R, uint8_t *states, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
