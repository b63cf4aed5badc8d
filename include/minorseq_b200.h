/*
 * minorseq_b200.h -- C ABI of the B200-native juliet/fuse hot path.
 *
 * The reference (PacificBiosciences/minorseq) documents no library, plugin or
 * FFI interface: its only boundary is the process interface
 *     juliet [--config|-c X] [--region b-e] [--mode-phasing] [--min-perc p]
 *            [--max-perc p] [--drm-only] in.bam out.{json,html}
 *                                        (/root/reference/doc/JULIET.md:61-66,
 *                                         :121,:195,:270-271,:342-354,:370)
 *     fuse in.bam out.fasta              (/root/reference/doc/FUSE.md:22-32)
 * so this header IS the drop-in boundary a maintainer would bind: the juliet /
 * fuse mains (minorseq_b200/host/), bench.py and the tests all sit on top of
 * it.  Each entry point names the documented juliet/fuse behaviour it replaces.
 *
 * Conventions: plain C types only; every function returns MS_OK (0) or a
 * negative ms_status; ms_last_error(h) holds the text; no exceptions cross the
 * ABI; one handle per GPU (per rank); calls on one handle are serialised by the
 * caller; all device work is enqueued on the handle's stream.  There is NO CPU
 * fallback: ms_create fails when no sm_100 device is present.
 *
 * Packed read format ("planar 4-bit", L/2 bytes per read rounded up to 16 B):
 *   a read is ceil(L/32) blocks; block b is four consecutive u32 bit-planes
 *   P0,P1,P2,P3 (words 4b..4b+3); bit j of each plane is reference column
 *   32b+j.  (P0,P1,P2) = state bits 0,1,2:
 *       0=A 1=C 2=G 3=T 4='-' deletion 5='N' QV-filtered base 7=not spanned
 *   P3 = an insertion follows this column in this read.  Columns >= L in the
 *   last block are state 7.
 * Two arrangements of the same 16-byte blocks:
 *   plain rows (HOST side: ms_expand_cigar, ms_pack_states, ms_pileup_host, ms_encode_rows): rows are
 *     contiguous, read r starts at word r * ms_row_words(L);
 *   tiles (DEVICE side: everything that takes a device pointer to packed reads -- ms_pileup_dev, ms_phase_dev,
 *     ms_juliet_pass_dev, ms_synth_dev, ms_expand_events_dev): reads are grouped in tiles of 8 consecutive
 *     reads stored block-major; the 16-byte block b of read 8t+i is the uint4 at index
 *     (t * nblk + b) * 8 + (i ^ (b & 7)),  nblk = ceil(L/32).  A buffer holds ceil(R/8) whole tiles
 *     (ms_tiled_words); the slots of reads >= R in the last tile are never interpreted, and a pointer into the
 *     middle of a buffer must start at a tile (a multiple of 8 reads).  One 128-byte line thus serves one block
 *     of 8 reads (the phasing gather) and a column segment of a tile is contiguous (the pile-up's bulk copies).
 *     ms_pileup_host / the event expansion produce tiles themselves; ms_tile_rows(_dev) converts plain rows.
 */
#ifndef MINORSEQ_B200_H
#define MINORSEQ_B200_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    MS_OK = 0,
    MS_ERR_ARG = -1,        /* bad argument / call order */
    MS_ERR_CUDA = -2,       /* CUDA runtime error, text in ms_last_error */
    MS_ERR_NODEVICE = -3,   /* no usable sm_100 GPU: there is no CPU fallback */
    MS_ERR_CAPACITY = -4,   /* caller buffer too small; required size reported */
    MS_ERR_FORMAT = -5      /* malformed input (e.g. CIGAR 'M', doc/JULIET.md:53) */
} ms_status;

enum { MS_A = 0, MS_C = 1, MS_G = 2, MS_T = 3, MS_DEL = 4, MS_N = 5, MS_UNCOV = 7 };
enum { MS_COL_INS = 6, MS_COL_COV = 7 };     /* col[j][6] insertion flags, col[j][7] = A+C+G+T+-+N */
enum { MS_FLAG_GAP = 1, MS_FLAG_HET = 2, MS_FLAG_PARTIAL = 4 };

typedef struct ms_handle ms_handle;

/* ---- format helpers (host only, no GPU work) -------------------------------- */
/* u32 words per packed read = 4*ceil(L/32). */
int32_t ms_row_words(int32_t L);
/* Read admission filter (doc/JULIET.md:58 "Reads that are not primary or supplementary
 * alignments, get ignored"): 1 = use the record, 0 = skip (BAM FLAG 0x4 unmapped, 0x100 secondary). */
int ms_read_admitted(uint32_t bam_flag);
/* one byte per column (bits0-2 state, bit3 insertion-follows) -> plain planar rows (host arrangement). */
int ms_pack_states(const uint8_t *states, int64_t R, int32_t L, uint32_t *packed);
int ms_unpack_states(const uint32_t *packed, int64_t R, int32_t L, uint8_t *states);
/* plain rows <-> device tiles (see the format note above).  ms_tiled_words = u32 words of a tile buffer for R reads. */
int64_t ms_tiled_words(int32_t L, int64_t R);
int ms_tile_rows(const uint32_t *rows, int64_t R, int32_t L, uint32_t *tiled /* ms_tiled_words(L,R) */);
int ms_untile_rows(const uint32_t *tiled, int64_t R, int32_t L, uint32_t *rows);
/* Replaces juliet/fuse's per-record CIGAR walk (doc/JULIET.md:49-58, doc/FUSE.md:13-15):
 * expands one aligned record into one packed row.  cigar = BAM-encoded ops
 * (len<<4|op); op 'M'(0) is rejected with MS_ERR_FORMAT ("cigar M is forbidden").
 * seq = one base per byte (ACGTN), qv_mask[i]!=0 turns base i into 'N'
 * (doc/JULIET.md:256-259), may be NULL.  Inserted strings are appended to the
 * caller's event arrays (ins_col/ins_off/ins_len/ins_pool) when non-NULL.       */
int ms_expand_cigar(const uint32_t *cigar, int32_t ncigar, int32_t pos,
                    const char *seq, const uint8_t *qv_mask, int32_t lseq,
                    int32_t L, uint32_t *row /* ms_row_words(L) */,
                    int32_t *ins_col, int64_t *ins_off, int32_t *ins_len,
                    int64_t ins_cap, int64_t *nins, char *ins_pool, int64_t pool_cap,
                    int64_t *pool_used);

/* ---- lifecycle --------------------------------------------------------------- */
int ms_create(int device, ms_handle **out);
void ms_destroy(ms_handle *h);
const char *ms_last_error(const ms_handle *h);   /* h == NULL: error of the last failed ms_create */
/* cudaStream_t to enqueue on (e.g. torch's current stream); NULL = the handle's own stream. */
int ms_set_stream(ms_handle *h, void *cuda_stream);
int ms_synchronize(ms_handle *h);
/* page-locked host memory for the packed rows handed to ms_pileup_host (NULL on failure) */
void *ms_alloc_pinned(size_t bytes);
void ms_free_pinned(void *p);
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t ms_launch_count(const ms_handle *h);
/* on: record CUDA events around each K1 launch on the handle's stream; ms_pileup_kernel_ms
 * then returns the device time of the most recent K1 kernel and the reads it processed
 * (bench.py's roofline.achieved).                                                     */
int ms_set_timing(ms_handle *h, int on);
/* CUDA-event stopwatch on the handle's stream: start records an event, stop records a second one,
 * waits for it and returns the device time between them in milliseconds.              */
int ms_timer_start(ms_handle *h);
int ms_timer_stop(ms_handle *h, double *ms);
int ms_pileup_kernel_ms(ms_handle *h, double *ms, int64_t *reads);
/* Same for the other measured kernels (most recent launch while timing was on; MS_ERR_ARG if none):
 * phase_bits_kernel, the co-occurrence kernel (popcount-AND or tcgen05), expand_events_kernel, and with a communicator
 * attached the two exchanges: the count all-reduce, and the haplotype-list all-gather + merge kernels.  Two spans of
 * the host-buffer pass (ms_juliet_pass_events_host) as well: MS_STAGE_UPLOAD = first to last host->device copy on the
 * copy stream, MS_STAGE_PASS = start of the pass to the arrival of its last result on the main stream.             */
enum { MS_STAGE_PHASE_BITS = 1, MS_STAGE_COOCCURRENCE = 2, MS_STAGE_EXPAND = 3, MS_STAGE_ALLREDUCE = 4, MS_STAGE_HAPMERGE = 5,
       MS_STAGE_UPLOAD = 6, MS_STAGE_PASS = 7 };
int ms_stage_kernel_ms(ms_handle *h, int stage, double *ms);

/* ---- K1: pileup (juliet "MSA counts", doc/JULIET.md:96-100; fuse doc/FUSE.md:17-20) -- */
/* Reference length and the columns where a codon of some configured gene starts
 * (bit j of word j/32; NULL = count no codons, i.e. fuse).  Zeroes the counts.
 * With a start mask the insertion flags are not tallied (col[j][6] stays 0: juliet
 * ignores insertions, doc/JULIET.md:26-27) unless ms_set_count_insertions(h, 1).  */
int ms_set_layout(ms_handle *h, int32_t L, const uint32_t *start_mask);
int ms_set_count_insertions(ms_handle *h, int on);
int ms_reset_counts(ms_handle *h);
/* Accumulate R device-resident packed reads (tiles) into the handle's count tensor.     */
int ms_pileup_dev(ms_handle *h, const uint32_t *d_packed, int64_t R);
/* device-resident plain rows -> tiles (d_tiled: ms_tiled_words(L,R) words; needs ms_set_layout) */
int ms_tile_rows_dev(ms_handle *h, const uint32_t *d_rows, int64_t R, uint32_t *d_tiled);
/* Same from host memory (plain rows, pinned recommended): chunked H2D copies, permuted into
 * tiles on the GPU and overlapped with the kernel.  If keep_dev != NULL it receives a device
 * pointer to the uploaded reads (tiles), owned by the handle and valid until the next
 * ms_pileup_host / ms_destroy, so ms_phase_dev can run without a second upload.          */
int ms_pileup_host(ms_handle *h, const uint32_t *h_packed, int64_t R, const uint32_t **keep_dev);
/* The count tensor [ col: L*8 u32 | codon: L*64 u32 ] in device memory, for the
 * caller's cross-GPU sum (one NCCL all-reduce; integer sums are order independent). */
int ms_counts_device(ms_handle *h, uint32_t **d_counts, int64_t *nwords);
/* Copy counts to host: col[L][8] = A,C,G,T,-,N,ins,coverage ; codon[L][64] indexed by
 * start column, codon = 16*b0+4*b1+b2, only reads whose codon is all A/C/G/T.    */
int ms_get_counts(ms_handle *h, uint32_t *col, uint32_t *codon);
/* selects the K1 implementation: 0 = bit-sliced carry-save kernel (default),
 * 1 = shared-memory-atomic histogram kernel (kept only as the ncu A/B baseline). */
int ms_set_pileup_variant(ms_handle *h, int variant);

/* ---- event rows: the compact host->device form of an aligned read ------------------------------
 * What juliet/fuse's per-record CIGAR walk (doc/JULIET.md:49-58, doc/FUSE.md:13-15) hands to the GPU when the
 * caller wants the PCIe link to carry events instead of whole rows: a CCS alignment is "the base sequence, except
 * at a few columns" (X / D / I ops and QV-filtered bases, :256-259).  Per read: its span [begin,end) and a byte string
 *   [nN: u16 little-endian] [N list: nN bytes] [rest list: 12-bit entries, entry k in bits [12k, 12k+12) of the rest]
 * (no bytes at all when the read equals the base on its whole span).  Both lists walk the columns from `begin`: an
 * entry's column = the previous entry's column + its delta.  N list, for the QV-filtered bases (two thirds of the
 * events of CCS data): one byte per entry, 0..254 = this column is 'N', 255 = move on by 255 columns.  Rest list:
 * entry = delta << 4 | nibble, delta 0..254 = this column holds nibble = state | insertion-follows << 3, delta 255 =
 * move on by 255 columns; n entries take ceil(1.5 n) bytes.  Spanned columns without an entry hold the base
 * sequence's base, columns outside the span are "not spanned"; no column appears twice.  hdr has R+1 entries: the
 * byte string of read r is events[hdr[r].ev_off .. hdr[r+1].ev_off); hdr[R] is the sentinel written by
 * ms_events_seal (total byte count + a hash of the base sequence; rows encoded against another base are rejected
 * with MS_ERR_FORMAT).  Reference length <= 65535.  The GPU rebuilds the packed reads (expand_events_kernel) and
 * K1/K3 run on them unchanged.  ~112 B per 3 kb read at CCS error rates (85 events, 60 of them 'N'), against
 * 1504 B for the plain row.                                                                                     */
typedef struct { uint32_t ev_off; uint16_t begin, end; } ms_read_hdr;
/* worst-case number of event BYTES of one read (every column in the rest list + skips + the count) */
int64_t ms_events_bound(int32_t L);
/* base: L bytes, one base (0..3) per reference column: the configured referenceSequence, or any sequence close to
 * the reads (the encoding is lossless for every base; a close one makes it short).                                */
int ms_encode_states(const uint8_t *states, int64_t R, int32_t L, const uint8_t *base,
                     ms_read_hdr *hdr /* R+1 */, uint8_t *events, int64_t cap /* bytes */, int64_t *nbytes);
int ms_encode_rows(const uint32_t *packed, int64_t R, int32_t L, const uint8_t *base,
                   ms_read_hdr *hdr /* R+1 */, uint8_t *events, int64_t cap /* bytes */, int64_t *nbytes);
/* Incremental form for a BAM loop (one ms_expand_cigar row at a time, e.g. per host thread): base_planes from
 * ms_base_planes (2*ceil(L/32) words); appends the row's event bytes at events[*nbytes...] and advances *nbytes;
 * fills *hdr (ev_off = the old *nbytes).  The caller seals the finished array with ms_events_seal.                */
int ms_base_planes(const uint8_t *base, int32_t L, uint32_t *planes);
int ms_encode_row(const uint32_t *row, int32_t L, const uint32_t *base_planes, ms_read_hdr *hdr,
                  uint8_t *events, int64_t cap /* bytes */, int64_t *nbytes);
int ms_events_seal(ms_read_hdr *hdr, int64_t R, int64_t nbytes, const uint8_t *base, int32_t L);
/* The handle's copy of the base sequence (after ms_set_layout, which forgets it). */
int ms_set_base(ms_handle *h, const uint8_t *base);
/* Device-resident event rows -> R packed reads at d_packed (tiles, ms_tiled_words(L,R) words).  */
int ms_expand_events_dev(ms_handle *h, const ms_read_hdr *d_hdr, const uint8_t *d_events, int64_t R,
                         uint32_t *d_packed);
/* ms_pileup_host for event rows: chunked H2D of the events, expansion and pile-up behind each chunk.
 * keep_dev as in ms_pileup_host (the expanded rows, for ms_phase_dev).                               */
int ms_pileup_events_host(ms_handle *h, const ms_read_hdr *hdr, const uint8_t *events, int64_t R,
                          const uint32_t **keep_dev);

/* ---- cross-GPU exchange (one process per GPU) ---------------------------------
 * Rank 0 creates the 128-byte NCCL unique id, the caller hands it to every rank over its own
 * control channel, every rank attaches its handle.  world == 1 is a no-op.              */
int ms_comm_unique_id(char id[128]);
int ms_comm_init(ms_handle *h, const char id[128], int rank, int world);
int ms_comm_size(const ms_handle *h);
/* The path's one data-path collective: in-place integer sum of the count tensor over all
 * ranks (ncclAllReduce, uint32, sum), enqueued on the handle's stream between K1 and K2.
 * With a communicator attached, ms_phase_groups also all-gathers the ranks' compact
 * (pattern, count) lists and returns the concatenation and the summed marginals.        */
int ms_allreduce_counts(ms_handle *h);

/* ---- K2: per-codon minor-variant test (doc/JULIET.md:38-42) --------------------- */
typedef struct { int32_t begin, end; } ms_gene;   /* 1-based [begin,end), doc/JULIET.md:134-136 */
typedef struct {
    double substitution_rate, deletion_rate;   /* error model (restatement choice U1) */
    double alpha;                              /* call iff p * ntests < alpha */
    double min_perc, max_perc;                 /* --min-perc / --max-perc, < 0 = off */
    int32_t region_begin, region_end;          /* --region, 1-based [b,e); 0,0 = off */
} ms_call_params;
typedef struct {
    int32_t gene, codon_index, col, ref_codon, codon;
    uint32_t count, coverage, expected, ntests;
    double pvalue;                             /* uncorrected one-sided Fisher p (fp64) */
} ms_variant;
void ms_call_params_default(ms_call_params *p);
/* Runs on the handle's (all-reduced) counts.  refseq NULL or shorter than L =>
 * tested against the major codon (doc/JULIET.md:133-134).  Variants come back
 * sorted by (gene, col, codon).  *n receives the number found (may exceed cap). */
int ms_call(ms_handle *h, const ms_gene *genes, int32_t ngenes, const char *refseq,
            const ms_call_params *prm, ms_variant *out, int64_t cap, int64_t *n);

/* ---- K3: read-level phasing (--mode-phasing, doc/JULIET.md:192-211,:372-381) ----- */
typedef struct { uint64_t reported, insufficient, damaged, gaps, heteroduplex, partial; } ms_phase_counters;
/* Declare the pooled variant list (start column, codon) and the number of reads
 * this handle will phase; allocates the bit matrix and the grouping table.       */
int ms_phase_begin(ms_handle *h, const int32_t *var_col, const int32_t *var_codon, int32_t V,
                   int64_t max_reads);
/* Bit-vectors + damage flags for R more device-resident reads (tiles), appended.          */
int ms_phase_dev(ms_handle *h, const uint32_t *d_packed, int64_t R);
/* Distinct patterns of this handle's undamaged reads with their counts (in no particular
 * order; with a communicator attached, the concatenation over all ranks), and the damage
 * marginals.  *H may exceed cap: call again with a larger buffer (the pass is cached).  */
int ms_phase_groups(ms_handle *h, uint32_t *patterns /* cap*ceil(V/32) */, uint64_t *counts,
                    int64_t cap, int64_t *H, ms_phase_counters *ctr);
/* Host: (pattern,count) lists from all ranks -> equal patterns merged, then juliet's
 * haplotype order (count desc, then ascending words) in place; entries [0,*nreported)
 * have count >= min_reads; the unreported rest follows in the same order (left unsorted
 * when it exceeds 65536 entries); fills reported/insufficient of ctr.  Pure host logic. */
int ms_haplotype_order(uint32_t *patterns, uint64_t *counts, int64_t H, int32_t V,
                       int32_t min_reads, int64_t *Hmerged, int64_t *nreported,
                       ms_phase_counters *ctr);
void ms_haplotype_name(int64_t rank, char buf[3]);      /* [A-Z][a-z]? doc/JULIET.md:198 */
/* hap_id (host, one per phased read, in upload order): rank in ordered_patterns,
 * or -1 if damaged.                                                               */
int ms_phase_assign(ms_handle *h, const uint32_t *ordered_patterns, int64_t H, int32_t *hap_id);
/* Grouping, merge over ranks (with a communicator: one all-gather of the ranks' compact
 * lists, merged on the device), juliet's haplotype order (count desc, then ascending
 * words) and the per-read haplotype id in ONE call, without the full pattern list ever
 * leaving the GPU: patterns/counts receive the first min(cap,*H) haplotypes of the order,
 * *H the number of distinct patterns, *nreported how many have >= min_reads reads (they
 * come first), ctr all six read categories of doc/JULIET.md:372-381.  hap_id (host, one per
 * read phased on this handle, may be NULL): position in that order, -1 if damaged.
 * Same results as ms_phase_groups + ms_haplotype_order + ms_phase_assign.               */
int ms_phase_haplotypes(ms_handle *h, int32_t min_reads, uint32_t *patterns, uint64_t *counts,
                        int64_t cap, int64_t *H, int64_t *nreported, ms_phase_counters *ctr,
                        int32_t *hap_id);
/* Device pointers to the per-read results (R*ceil(V/32) words, R bytes).          */
int ms_phase_device(ms_handle *h, uint32_t **d_bits, uint8_t **d_flags, int64_t *R);
/* C[v][w] = #reads carrying both (V*V int32, device buffer owned by the handle).
 * Popcount-AND over the transposed bit matrix, or -- when the contraction is large (V >= 256 and
 * >= 32768 reads) -- u8 x u8 -> s32 tcgen05 MMA with the accumulator in tensor memory; same integers. */
int ms_cooccurrence(ms_handle *h, int32_t **d_C);
/* 0 = choose by size (default), 1 = popcount-AND kernel, 2 = tensor-core kernel (A/B measurements, tests) */
int ms_set_cooccurrence_variant(ms_handle *h, int32_t variant);

/* ---- the whole juliet pass in one call ------------------------------------------------------
 * reset -> pileup -> (all-reduce when a communicator is attached) -> codon test -> phasing, i.e. everything
 * juliet does between BAM decode and report writing.  The caller owns the result buffers; on
 * MS_ERR_CAPACITY the n* fields hold the sizes needed.  patterns/counts receive the first
 * min(patterns_cap, npatterns) haplotypes of juliet's order, [0,nreported) being the reported ones (rows
 * ceil(nkeys/32) words apart; the buffer must hold patterns_cap * ceil(keys_cap/32) words); npatterns is the
 * number of distinct patterns, and only nreported > patterns_cap counts as too small.  key_col/key_codon is
 * the pooled variant list the bit-vectors refer to.  hap_id (host, R entries, may be NULL) as in
 * ms_phase_haplotypes.  Afterwards ms_get_counts / ms_cooccurrence can be called as usual.            */
typedef struct {
    ms_variant *variants;  int64_t variants_cap, nvariants;
    int32_t *key_col, *key_codon; int32_t keys_cap, nkeys;
    uint32_t *patterns; uint64_t *counts; int64_t patterns_cap, npatterns, nreported;
    ms_phase_counters counters;
    int32_t *hap_id;
} ms_juliet_result;
int ms_juliet_pass_dev(ms_handle *h, const uint32_t *d_packed, int64_t R, const ms_gene *genes, int32_t ngenes,
                       const char *refseq, const ms_call_params *prm, int32_t phase, int32_t min_hap_reads,
                       ms_juliet_result *out);
int ms_juliet_pass_host(ms_handle *h, const uint32_t *h_packed, int64_t R, const ms_gene *genes, int32_t ngenes,
                        const char *refseq, const ms_call_params *prm, int32_t phase, int32_t min_hap_reads,
                        ms_juliet_result *out);

/* The same pass from event rows in host memory (ms_pileup_events_host in place of ms_pileup_host): what the
 * bench.py's e2e calls -- the PCIe link carries ~112 B instead of 1504 B per 3 kb read.                       */
int ms_juliet_pass_events_host(ms_handle *h, const ms_read_hdr *hdr, const uint8_t *events, int64_t R,
                               const ms_gene *genes, int32_t ngenes, const char *refseq, const ms_call_params *prm,
                               int32_t phase, int32_t min_hap_reads, ms_juliet_result *out);

/* ---- K4: fuse consensus (doc/FUSE.md:17-24) --------------------------------------- */
typedef struct { int32_t min_coverage; double ins_fraction; int32_t ins_distance; } ms_fuse_params;
void ms_fuse_params_default(ms_fuse_params *p);
/* Consensus of the handle's (all-reduced) column counts plus the in-frame
 * insertion rule over the caller's insertion events.  *len may exceed cap.        */
int ms_fuse(ms_handle *h, const ms_fuse_params *prm,
            const int32_t *ins_col, const int64_t *ins_off, const int32_t *ins_len,
            int64_t nins, const char *ins_pool, int64_t pool_len,
            char *seq, int64_t cap, int64_t *len);

/* ---- cleric's alignment step (doc/CLERIC.md:19-23,41-44) ------------------------
 * Global Needleman-Wunsch of the original reference a against the target reference b on the GPU
 * (tiled wavefront, N x M cells; match +2, mismatch -3, linear gap -4, ties diagonal > consume-a >
 * consume-b -- restatement choice U13).  ops receives the path in forward order: 'M' both advance,
 * 'D' only a advances (a base of a that b lacks), 'I' only b advances; cap must be >= la + lb.
 * The per-read CIGAR projection through this path is host code (minorseq_b200/host/cleric.hpp).   */
int ms_align_refs(ms_handle *h, const char *a, int32_t la, const char *b, int32_t lb,
                  char *ops, int64_t cap, int64_t *nops, int64_t *score);

/* ---- synthetic amplicon generator (bench/tests) ------------------------- */
typedef struct {
    uint64_t seed;
    int32_t L, nstrains;
    uint32_t thr_N, thr_sub, thr_ins20, thr_trunc16;  /* integer thresholds, see synth.py */
} ms_synth_params;
/* strain_base: nstrains*L bytes (0..3); thr_del: L u32; strain_cum: nstrains u32 cumulative
 * mixture thresholds.  Writes R packed reads (tiles, ms_tiled_words(L,R) words) for reads [read0, read0+R) on the device.  */
int ms_synth_dev(ms_handle *h, const ms_synth_params *p, const uint8_t *strain_base,
                 const uint32_t *thr_del, const uint32_t *strain_cum,
                 int64_t read0, int64_t R, uint32_t *d_packed);

#ifdef __cplusplus
}
#endif
#endif
