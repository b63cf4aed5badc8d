"""Screenshot-derived known answers -- the only pins the documentation-only
reference offers for the oracle's *definitions* (what coverage, %, categories mean)."""
import numpy as np

A, Cc, G, T, DEL, N, UNC = 0, 1, 2, 3, 4, 5, 7


def _context_reads():
    """2998 reads x 9 columns reproducing juliet_hiv-context.png (rows -3..+5 around K65, AAA)."""
    R = 2998
    s = np.zeros((R, 9), dtype=np.uint8)            # all 'A'
    s[:, 2] = G                                     # rel -1 is G (2952 G)
    # rel -3: 2947 A, 51 N
    s[:51, 0] = N
    # rel -2: 2923 A, 2 G, 73 N
    s[100:173, 1] = N
    s[200:202, 1] = G
    # rel -1: 4 A, 2952 G, 42 N
    s[300:342, 2] = N
    s[400:404, 2] = A
    # rel 0: 2606 A, 339 '-', 53 N   -> 392 dirty reads [1000,1392)
    s[1000:1339, 3] = DEL
    s[1339:1392, 3] = N
    # rel 1: 2905 A, 29 G, 64 N; 47 of the N overlap the rel-0 dirty reads
    s[1345:1392, 4] = N
    s[1392:1409, 4] = N
    # rel 2: 2938 A, 60 N, disjoint from the above
    s[1500:1560, 5] = N
    # K65R AAA->AGA: 28 clean reads + 1 read that is dirty at rel 0
    s[2000:2028, 4] = G
    s[1000, 4] = G
    # rel 3, 4, 5
    s[2100:2160, 6] = N
    s[2200:2256, 7] = N
    s[2300:2547, 8] = N
    return s


def test_context_screenshot_invariants(oracle):
    s = _context_reads()
    start = np.zeros(9, dtype=np.uint8)
    start[[0, 3, 6]] = 1
    col, codon = oracle.pileup(s, start)
    # C-1: every row sums to the number of reads spanning it
    assert (col[:, :6].sum(axis=1) == 2998).all()
    assert (col[:, 7] == 2998).all()
    rows = {0: (2947, 0, 0, 0, 0, 51), 1: (2923, 0, 2, 0, 0, 73), 2: (4, 0, 2952, 0, 0, 42),
            3: (2606, 0, 0, 0, 339, 53), 4: (2905, 0, 29, 0, 0, 64), 5: (2938, 0, 0, 0, 0, 60),
            6: (2938, 0, 0, 0, 0, 60), 7: (2942, 0, 0, 0, 0, 56), 8: (2751, 0, 0, 0, 0, 247)}
    for j, want in rows.items():
        assert tuple(int(x) for x in col[j, :6]) == want
    # C-2: coverage = reads with a clean codon <= min ACGT count of the three columns
    cov = int(codon[3].sum())
    assert cov == 2529
    assert cov <= min(int(col[j, :4].sum()) for j in (3, 4, 5))
    # C-3: frequency = codon count / coverage, two significant digits -> "1.1"
    aga = 16 * A + 4 * G + A
    assert int(codon[3, aga]) == 28
    assert f"{100.0 * codon[3, aga] / cov:.2g}" == "1.1"


def test_k65r_is_called(oracle):
    s = _context_reads()
    start = np.zeros(9, dtype=np.uint8)
    start[[0, 3, 6]] = 1
    _, codon = oracle.pileup(s, start)
    v = oracle.call(codon, [(1, 10)], refseq="AAGAAAAAA")
    hits = [x for x in v if x.col == 3]
    assert len(hits) == 1 and hits[0].codon == 8 and hits[0].coverage == 2529 and hits[0].count == 28
    assert hits[0].expected == 1 and hits[0].ntests == 3
    assert abs(hits[0].pvalue / 5.2e-8 - 1) < 0.02


def test_category_partition_and_percentages(oracle):
    """C-4 / C-5 / C-7: reported + insufficient + damaged = all reads; marginals overlap;
    haplotype % = count / reported reads; wild type is itself a haplotype."""
    R = 5695
    V = 5
    bits = np.zeros((R, 1), dtype=np.uint32)
    flags = np.zeros(R, dtype=np.uint8)
    flags[:3894] = 2                    # heteroduplex
    flags[:786] |= 1                    # gaps overlap
    flags[3800:3876] |= 4               # partial overlap
    flags[3709:3800] = 1                # gap only: keep 3894 damaged in total
    und = np.arange(3894, R)
    bits[und[:1500], 0] = 0             # wild type
    bits[und[1500:1700], 0] = 1
    bits[und[1700:1735], 0] = 6
    for i, r in enumerate(und[1735:]):  # 66 reads in patterns of < 10 reads
        bits[r, 0] = 8 + i % 11        # 11 patterns x 6 reads
    g = oracle.phase_group(bits, flags, V)
    c = g["counters"]
    assert c["damaged"] == 3894 and c["reported"] == 1735 and c["insufficient"] == 66
    assert c["reported"] + c["insufficient"] + c["damaged"] == R
    assert c["gaps"] + c["heteroduplex"] + c["partial"] >= c["damaged"]
    assert g["nreported"] == 3
    assert list(g["counts"][:3]) == [1500, 200, 35]
    assert g["patterns"][0, 0] == 0          # wild type first: most reads, "plain dark gray" column A
    perc = [100.0 * int(x) / c["reported"] for x in g["counts"][:3]]
    assert abs(sum(perc) - 100.0) < 1e-9
    assert (g["hap_id"][:3894] == -1).all() and (g["hap_id"][und[:1500]] == 0).all()


def test_haplotype_names(oracle):
    assert [oracle.hap_name(i) for i in (0, 1, 25, 26, 27, 51, 52)] == ["A", "B", "Z", "Aa", "Ab", "Az", "Ba"]


def test_major_variants_fragment_minors(oracle):
    """C-6: a haplotype is the exact pattern over ALL called variants."""
    R = 400
    rng = np.random.default_rng(3)
    bits = np.zeros((R, 1), dtype=np.uint32)
    bits[:12, 0] = 1                      # a 12-read minor on variant 0 only
    flags = np.zeros(R, dtype=np.uint8)
    g1 = oracle.phase_group(bits, flags, 1)
    assert g1["nreported"] == 2
    # add 5 "major" variants each present in ~half of the reads: the minor splinters below 10 reads
    extra = (rng.integers(0, 32, size=R).astype(np.uint32) << 1)
    g2 = oracle.phase_group(bits | extra[:, None], flags, 6)
    minor_groups = [int(c) for p, c in zip(g2["patterns"][:, 0], g2["counts"]) if p & 1]
    assert sum(minor_groups) == 12 and max(minor_groups) < 10
