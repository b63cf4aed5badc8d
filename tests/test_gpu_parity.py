"""GPU parity: every kernel behind the C ABI against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): counts, consensus, haplotypes bit-exact; p-values to a
relative tolerance of 1e-9.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from minorseq_b200 import Fuse, Handle, Juliet, _lib, device_rows, host_rows, synth_device  # noqa: E402
from minorseq_b200._lib import SynthParams  # noqa: E402
from minorseq_b200.synth import (SynthConfig, make_tables, pack_states, read_strains, start_mask_words,  # noqa: E402
                                 synth_states)

PVAL_RTOL = 1e-9


def to_dev(a):
    """plain host rows [R, row_words] -> device tiles (csrc/rows.cuh); anything else goes up as it is"""
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint32 and a.ndim == 2:
        return device_rows(a)
    return torch.from_numpy(a.view(np.int32)).cuda()


def mask_bytes(words, L):
    return np.array([(int(words[j >> 5]) >> (j & 31)) & 1 for j in range(L)], dtype=np.uint8)


def gpu_synth(hd, t, read0, R):
    return synth_device(hd, t, read0, R)


@pytest.fixture(scope="module")
def hd():
    h = Handle(0)
    yield h
    h.close()


@pytest.fixture(scope="module")
def c1():
    """BASELINE configs[0]: 5k CCS reads x 3 kb, one gene in frame 0, 4 strains."""
    cfg = SynthConfig(L=3000, seed=20240001)
    t = make_tables(cfg)
    st = synth_states(t, 0, 5000)
    return t, st, pack_states(st)


def check_pileup(oracle, hd, st, L, genes, variant=0, count_ins=False):
    packed = pack_states(st) if st.shape[0] else np.zeros((0, 4 * ((L + 31) // 32)), dtype=np.uint32)
    j = Juliet(L, genes, handle=hd)
    j.set_count_insertions(count_ins)
    _lib.check(j.lib.ms_set_pileup_variant(hd.h, variant), hd.h)
    d = to_dev(packed) if st.shape[0] else None
    j.pileup_device(d.data_ptr() if d is not None else 0, st.shape[0])
    col, codon = j.get_counts()
    _lib.check(j.lib.ms_set_pileup_variant(hd.h, 0), hd.h)
    ocol, ocodon = oracle.pileup(st, mask_bytes(j.start_mask, L)) if st.shape[0] else (np.zeros((L, 8), np.uint32), np.zeros((L, 64), np.uint32))
    if not count_ins and variant == 0:
        ocol = ocol.copy()
        ocol[:, 6] = 0          # juliet mode does not tally insertion flags (doc/JULIET.md:26-27)
    j.set_count_insertions(False)
    assert np.array_equal(col, ocol), f"column counts differ at {np.argwhere(col != ocol)[:5]}"
    assert np.array_equal(codon, ocodon), f"codon counts differ at {np.argwhere(codon != ocodon)[:5]}"
    return j, d, col, codon


def test_pileup_c1(oracle, hd, c1):
    t, st, _ = c1
    check_pileup(oracle, hd, st, 3000, [(1, 3001)])


def test_pileup_c1_with_insertion_tally(oracle, hd, c1):
    t, st, _ = c1
    check_pileup(oracle, hd, st, 3000, [(1, 3001), (2, 3001)], count_ins=True)


def test_pileup_c1_atomic_variant(oracle, hd, c1):
    t, st, _ = c1
    check_pileup(oracle, hd, st[:1500], 3000, [(1, 3001)], variant=1)


@pytest.mark.parametrize("R", [0, 1, 7, 8, 9, 23, 24, 25, 71, 72, 73, 1000])
def test_pileup_ragged_read_counts(oracle, hd, c1, R):
    _, st, _ = c1
    check_pileup(oracle, hd, st[:R], 3000, [(1, 3001)])


@pytest.mark.parametrize("L", [3, 5, 32, 33, 64, 95, 96, 97, 1000, 1024, 1025, 4096, 4100, 6144, 7000, 9719, 11936, 12001, 20000])
def test_pileup_reference_lengths(oracle, hd, L):
    cfg = SynthConfig(L=L, seed=77 + L, trunc=0.2, variants_per_minor=(1, 2) if L >= 30 else (0, 0),
                      minor_fracs=(0.1, 0.05) if L >= 30 else ())
    t = make_tables(cfg)
    st = synth_states(t, 0, 777 if L < 5000 else 300)
    # three overlapping genes, one per reading frame (doc/JULIET.md:261-264)
    genes = [(1, L + 1), (2, L + 1), (3, L + 1)] if L >= 9 else [(1, L + 1)]
    check_pileup(oracle, hd, st, L, genes)


def test_pileup_all_states_and_uncovered(oracle, hd):
    rng = np.random.default_rng(5)
    L, R = 200, 500
    st = rng.choice(np.array([0, 1, 2, 3, 4, 5, 7], dtype=np.uint8), size=(R, L)).astype(np.uint8)
    st |= (rng.integers(0, 2, size=(R, L), dtype=np.uint8) << 3)
    st = np.where((st & 7) == 7, np.uint8(7), st).astype(np.uint8)
    st[:20] = 7                                  # reads that span nothing
    check_pileup(oracle, hd, st, L, [(1, L + 1), (3, L + 1)])


def test_pileup_reference_differs_from_sample(oracle, hd):
    """The pivot is a performance device only: a sample whose majority is far from any fixed
    guess (here: random per-read bases, no majority at all) must still count exactly."""
    rng = np.random.default_rng(9)
    L, R = 300, 2000
    st = rng.integers(0, 4, size=(R, L), dtype=np.uint8)
    check_pileup(oracle, hd, st, L, [(1, L + 1), (2, L + 1), (3, L + 1)])


def test_pileup_dense_high_frequency_variants(oracle, hd):
    """Phasing-stress style data: a variant codon in half of the reads at every third codon.  The pivot sample sees
    > 2 % non-pivot bases and K1 runs its DENSE instantiation (second-codon counters in shared memory); counts
    must not care.  Also with ragged read counts and across two accumulated batches."""
    cfg = SynthConfig(L=960, seed=56, dense_sites=150, dense_strains=16, n_rate=1e-3, dele=1e-3, trunc=0.02)
    t = make_tables(cfg)
    st = synth_states(t, 0, 6001)
    check_pileup(oracle, hd, st, 960, [(1, 961)])
    check_pileup(oracle, hd, st[:777], 960, [(1, 961)], count_ins=True)
    j = Juliet(960, [(1, 961)], handle=hd)
    j.reset()
    d = to_dev(pack_states(st))
    j.pileup_device(d.data_ptr(), 4000)
    j.pileup_device(d.data_ptr() + 4000 * j.row_words * 4, 2001)
    col, codon = j.get_counts()
    ocol, ocodon = oracle.pileup(st, mask_bytes(j.start_mask, 960))
    ocol = ocol.copy()
    ocol[:, 6] = 0
    assert np.array_equal(col, ocol) and np.array_equal(codon, ocodon)


def test_pileup_dense_variants_on_segmented_rows(oracle, hd):
    """BASELINE configs[4] shape at reduced read count: L = 6144 rows are cut into single-warp column segments AND
    the data make K1 pick its DENSE instantiation -- the one combination of the two template switches the other tests
    do not reach.  Also a 3-frame layout on the same rows (logged rare path + segments)."""
    cfg = SynthConfig(L=6144, seed=20240005, dense_sites=2048, dense_strains=64, n_rate=2e-5, dele=2e-5, trunc=0.0)
    t = make_tables(cfg)
    st = synth_states(t, 0, 1501)
    check_pileup(oracle, hd, st, 6144, [(1, 6145)])
    check_pileup(oracle, hd, st[:300], 6144, [(1, 6145), (2, 6145), (3, 6145)])


def test_pileup_accumulates_batches_and_host_path(oracle, hd, c1):
    t, st, packed = c1
    L = 3000
    j = Juliet(L, [(1, 3001)], handle=hd)
    d = to_dev(packed)
    row = packed.shape[1]
    j.pileup_device(d.data_ptr(), 2000)
    j.pileup_device(d.data_ptr() + 2000 * row * 4, 3000)
    col, codon = j.get_counts()
    ocol, ocodon = oracle.pileup(st, mask_bytes(j.start_mask, L))
    ocol[:, 6] = 0
    assert np.array_equal(col, ocol) and np.array_equal(codon, ocodon)
    j.reset()
    pinned = torch.from_numpy(packed.view(np.int32)).pin_memory()
    j.pileup_host(pinned.numpy().view(np.uint32))
    col2, codon2 = j.get_counts()
    assert np.array_equal(col2, ocol) and np.array_equal(codon2, ocodon)


def test_synth_gpu_equals_numpy(hd):
    for cfg in (SynthConfig(L=3000, seed=20240003), SynthConfig(L=97, seed=5, trunc=0.5),
                SynthConfig(L=700, seed=6, dense_sites=200, dense_strains=16)):
        t = make_tables(cfg)
        g = host_rows(gpu_synth(hd, t, 1000, 600), 600, t.cfg.L)
        assert np.array_equal(g, pack_states(synth_states(t, 1000, 600)))


def test_pileup_counter_overflow_flush(oracle, hd):
    """More than 2040 reads per row-group forces the mid-kernel flush of the 11-plane counters."""
    cfg = SynthConfig(L=96, seed=31, variants_per_minor=(1, 1), minor_fracs=(0.05,))
    t = make_tables(cfg)
    R = 148 * 12 * 2040 + 12345
    d = gpu_synth(hd, t, 0, R)
    j = Juliet(96, [(1, 97), (2, 97)], handle=hd)
    j.pileup_device(d.data_ptr(), R)
    col, codon = j.get_counts()
    st = oracle.unpack(host_rows(d, R, 96), 96)
    ocol, ocodon = oracle.pileup(st, mask_bytes(j.start_mask, 96), nthreads=8)
    ocol[:, 6] = 0
    assert np.array_equal(col, ocol) and np.array_equal(codon, ocodon)


def variants_equal(gv, ov):
    assert len(gv) == len(ov)
    for a, b in zip(gv, ov):
        for f in ("gene", "codon_index", "col", "ref_codon", "codon", "count", "coverage", "expected", "ntests"):
            assert getattr(a, f) == getattr(b, f), (f, a.col, a.codon)
        if b.pvalue < 1e-290:
            assert a.pvalue < 1e-280
        else:
            assert abs(a.pvalue - b.pvalue) <= PVAL_RTOL * b.pvalue, (a.col, a.codon, a.pvalue, b.pvalue)


def test_call_c1(oracle, hd, c1):
    t, st, packed = c1
    genes = [(1, 3001)]
    j, d, col, codon = check_pileup(oracle, hd, st, 3000, genes)
    gv = j.call()
    ov = oracle.call(codon, genes)
    variants_equal(gv, ov)
    # planted 10 % and 5 % minors must be found (1 % of 5000 reads is below the doc's 2500x minimum... still usually found)
    found = {(v.col, v.codon) for v in gv}
    for strain, colv, cod in t.truth:
        if strain in (1, 2):
            assert (colv, cod) in found
    # with a reference sequence, region and the percentage filters
    j2 = Juliet(3000, genes, refseq=t.refseq, region=(301, 2402), min_perc=2.0, max_perc=50.0, handle=hd)
    j2.pileup_device(d.data_ptr(), st.shape[0])
    _, codon2 = j2.get_counts()
    variants_equal(j2.call(), oracle.call(codon2, genes, refseq=t.refseq, region=(301, 2402), min_perc=2.0, max_perc=50.0))


def test_call_multi_gene_frames_and_big_counts(oracle, hd):
    """Overlapping genes in three frames, plus synthetic histograms with coverage up to 1e6
    pushed straight into the count tensor (exercises the saddle-point Fisher at scale)."""
    L = 300
    genes = [(1, 151), (100, 301), (2, 200), (3, 299)]
    rng = np.random.default_rng(12)
    codon = np.zeros((L, 64), dtype=np.uint32)
    mask = mask_bytes(start_mask_words(L, genes), L)
    for s in np.nonzero(mask)[0]:
        n = int(10 ** rng.uniform(1, 6))
        major = int(rng.integers(0, 64))
        codon[s, major] = n
        for _ in range(int(rng.integers(0, 5))):
            codon[s, int(rng.integers(0, 64))] += int(10 ** rng.uniform(0, np.log10(n)))
    j = Juliet(L, genes, handle=hd)
    ct = j.counts_tensor()
    ct[L * 8:] = torch.from_numpy(codon.reshape(-1).view(np.int32)).cuda()
    torch.cuda.synchronize()
    variants_equal(j.call(), oracle.call(codon, genes))
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, size=L))
    j.refseq = ref
    variants_equal(j.call(), oracle.call(codon, genes, refseq=ref))


def phase_equal(oracle, hd, st, L, var_col, var_codon, packed_dev=None, host_merge=False, cap=4096):
    class V:  # minimal variant record for Juliet.phase_device
        def __init__(self, c, k):
            self.col, self.codon = c, k
    j = Juliet(L, [(1, L + 1)], mode_phasing=True, handle=hd)
    d = packed_dev if packed_dev is not None else to_dev(pack_states(st))
    hap, keys = j.phase_device([V(c, k) for c, k in zip(var_col, var_codon)], d.data_ptr(), st.shape[0], host_merge=host_merge, cap=cap)
    kc = [k[0] for k in keys]
    kd = [k[1] for k in keys]
    obits, oflags = oracle.phase_bits(st, kc, kd)
    g = oracle.phase_group(obits, oflags, len(keys))
    pb, pf, pn = C.c_void_p(), C.c_void_p(), C.c_int64()
    _lib.check(j.lib.ms_phase_device(hd.h, C.byref(pb), C.byref(pf), C.byref(pn)), hd.h)
    nw = max(1, (len(keys) + 31) // 32)
    from minorseq_b200.api import _as_tensor
    if st.shape[0]:
        gbits = _as_tensor(pb.value, (st.shape[0] * nw,), torch.int32, 0).cpu().numpy().view(np.uint32).reshape(-1, nw)
        gflags = _as_tensor(pf.value, (st.shape[0],), torch.uint8, 0).cpu().numpy()
        assert np.array_equal(gflags, oflags)
        assert np.array_equal(gbits, obits)
    k = len(hap.counts)       # the device-ordered path hands back the first min(cap, H) haplotypes of the order
    assert hap.nreported == g["nreported"] and k >= hap.nreported
    assert (k == g["H"]) if host_merge else (hap.ndistinct == g["H"] and k == min(g["H"], max(cap, hap.nreported)))
    assert np.array_equal(hap.counts, g["counts"][:k]) and np.array_equal(hap.patterns, g["patterns"][:k])
    assert hap.counters == {k: int(v) for k, v in g["counters"].items()}
    assert np.array_equal(hap.hap_id, g["hap_id"])
    assert hap.names == [oracle.hap_name(i) for i in range(g["nreported"])]
    return j, hap, obits


def test_phase_c3_style(oracle, hd):
    """juliet --mode-phasing on a 4-strain mix (BASELINE configs[2] shape, reduced read count)."""
    cfg = SynthConfig(L=3000, seed=20240003, n_rate=2e-3)
    t = make_tables(cfg)
    st = synth_states(t, 0, 20000)
    genes = [(1, 3001)]
    j, d, col, codon = check_pileup(oracle, hd, st, 3000, genes)
    gv = j.call()
    variants_equal(gv, oracle.call(codon, genes))
    assert len(gv) >= 8
    _, hap, _ = phase_equal(oracle, hd, st, 3000, [v.col for v in gv], [v.codon for v in gv], d)
    assert hap.nreported >= 4 and hap.patterns[0].sum() == 0       # wild type is haplotype A
    c = hap.counters
    assert c["reported"] + c["insufficient"] + c["damaged"] == st.shape[0]


@pytest.mark.parametrize("V", [0, 1, 31, 32, 33, 100])
def test_phase_variant_counts(oracle, hd, V):
    cfg = SynthConfig(L=600, seed=100 + V, n_rate=1e-3, dele=1e-3, trunc=0.05)
    t = make_tables(cfg)
    st = synth_states(t, 0, 3000)
    rng = np.random.default_rng(V)
    cols = np.sort(rng.choice(np.arange(0, 598), size=V, replace=False)) if V else np.array([], dtype=int)
    cods = [16 * int(t.strain_base[0, c]) + 4 * int(t.strain_base[0, c + 1]) + int(t.strain_base[0, c + 2]) for c in cols]
    phase_equal(oracle, hd, st, 600, list(cols), cods)


def test_phase_dense_and_cooccurrence(oracle, hd):
    """Phasing stress layout (BASELINE configs[4] shape, reduced): many shared sites, dense bits."""
    cfg = SynthConfig(L=960, seed=55, dense_sites=150, dense_strains=16, n_rate=1e-4, dele=1e-4, trunc=0.0)
    t = make_tables(cfg)
    st = synth_states(t, 0, 4000)
    sites = sorted({(c, k) for (_, c, k) in t.truth})
    j, hap, obits = phase_equal(oracle, hd, st, 960, [s[0] for s in sites], [s[1] for s in sites])
    Cg = j.cooccurrence().cpu().numpy()
    assert np.array_equal(Cg, oracle.cooccurrence(obits, len(sites)))
    assert (np.diag(Cg) > 0).all()


def _many_pattern_states(R, L, V, ndup, seed):
    """Reads that are either one of `ndup` recurring site patterns or a random one: thousands of distinct
    single-read haplotypes next to a few reported ones, as in BASELINE configs[4]."""
    rng = np.random.default_rng(seed)
    ref = rng.integers(0, 4, size=L, dtype=np.uint8)
    cols = np.sort(rng.choice(np.arange(0, L - 2, 3), size=V, replace=False))
    alt = (ref[cols] + 1 + rng.integers(0, 3, size=V, dtype=np.uint8)) % 4       # variant codon differs in its first base
    cods = [16 * int(alt[i]) + 4 * int(ref[c + 1]) + int(ref[c + 2]) for i, c in enumerate(cols)]
    st = np.tile(ref, (R, 1))
    carry = rng.random((R, V)) < 0.5
    proto = rng.random((ndup, V)) < 0.5
    which = rng.integers(0, ndup, size=R)
    recurring = rng.random(R) < 0.3
    carry[recurring] = proto[which[recurring]]
    st[:, cols] = np.where(carry, alt[None, :], ref[cols][None, :])
    st[rng.random(R) < 0.05, cols[0]] = 4          # some gapped reads
    st[rng.random(R) < 0.05, cols[-1] + 1] = 5     # some heteroduplex reads
    return st, list(cols), cods


@pytest.mark.parametrize("R,V,cap", [(12000, 40, 4096), (12000, 70, 20000), (3000, 70, 100)])
def test_phase_many_distinct_patterns(oracle, hd, R, V, cap):
    """More distinct patterns than the counting rank takes (4096): the bitonic path; and fewer: the counting path
    with two-word patterns.  Either way the full oracle order (count desc, pattern asc) and every read's id match."""
    st, cols, cods = _many_pattern_states(R, 900, V, 25, seed=R + V)
    _, hap, _ = phase_equal(oracle, hd, st, 900, cols, cods, cap=cap)
    assert hap.nreported >= 20 and hap.ndistinct > (4096 if R > 5000 else 1000)


@pytest.mark.parametrize("seed", range(8))
def test_phase_order_randomised(oracle, hd, seed):
    """Random read counts, variant counts and recurrence levels: few variants give heavy ties in the counts (order decided
    by the pattern words), many give long bit-vectors; every case must reproduce the oracle's full order and read ids."""
    rng = np.random.default_rng(1000 + seed)
    R = int(rng.integers(50, 6000))
    V = int(rng.choice([1, 2, 3, 5, 9, 31, 32, 33, 64, 65, 90]))
    ndup = int(rng.integers(1, 40))
    L = 3 * V + 30 + int(rng.integers(0, 50))
    st, cols, cods = _many_pattern_states(R, L, V, ndup, seed=seed)
    phase_equal(oracle, hd, st, L, cols, cods, cap=int(rng.choice([1, 50, 4096, 10000])))


def test_phase_host_merge_protocol(oracle, hd):
    """ms_phase_groups + ms_haplotype_order + ms_phase_assign (the three-call protocol the torch.distributed
    fallback uses) gives the same answer as the device-ordered call."""
    st, cols, cods = _many_pattern_states(3000, 900, 40, 25, seed=7)
    phase_equal(oracle, hd, st, 900, cols, cods, host_merge=True)
    phase_equal(oracle, hd, st, 900, cols, cods)


@pytest.mark.parametrize("R,V", [(300, 40), (5000, 300), (33000, 257)])
def test_cooccurrence_tensor_core_equals_popcount_and_oracle(oracle, hd, R, V):
    """a13 on tcgen05 (u8 x u8 -> s32 UMMA, cooc_tc.cu): same integers as the popcount-AND kernel and the oracle, for
    read counts that are not a multiple of the 128-read stage and variant counts that are not a multiple of the tile."""
    lib = _lib.load()
    st, cols, cods = _many_pattern_states(R, 3 * V + 60, V, 25, seed=R + V)
    j, hap, obits = phase_equal(oracle, hd, st, 3 * V + 60, cols, cods)
    want = oracle.cooccurrence(obits, V)
    try:
        for variant in (1, 2, 0):
            _lib.check(lib.ms_set_cooccurrence_variant(hd.h, variant), hd.h)
            assert np.array_equal(j.cooccurrence().cpu().numpy(), want), variant
    finally:
        _lib.check(lib.ms_set_cooccurrence_variant(hd.h, 0), hd.h)
    assert lib.ms_set_cooccurrence_variant(hd.h, 3) == -1


def test_phase_haplotypes_edge_cases(oracle, hd):
    """ms_phase_haplotypes: no reads, every read damaged, cap = 0 / no read ids, fewer slots than reported haplotypes."""
    lib = _lib.load()
    st, cols, cods = _many_pattern_states(2000, 900, 40, 25, seed=3)
    phase_equal(oracle, hd, st[:0], 900, cols, cods)                       # no reads at all
    dead = st.copy()
    dead[:, cols[3]] = 4
    _, hap, _ = phase_equal(oracle, hd, dead, 900, cols, cods)             # every read has a gap in a variant codon
    assert hap.ndistinct == 0 and hap.counters["damaged"] == 2000 and (hap.hap_id == -1).all()
    _, hap, _ = phase_equal(oracle, hd, st, 900, cols, cods, cap=3)        # the API grows cap to the reported count
    assert len(hap.counts) == hap.nreported > 3
    # raw call: cap 0, no buffers, no ids -> only the tallies
    vc, vk = np.array(cols, dtype=np.int32), np.array(cods, dtype=np.int32)
    d = to_dev(pack_states(st))
    _lib.check(lib.ms_set_layout(hd.h, 900, None), hd.h)
    _lib.check(lib.ms_phase_begin(hd.h, vc.ctypes.data_as(C.c_void_p), vk.ctypes.data_as(C.c_void_p), 40, 2000), hd.h)
    _lib.check(lib.ms_phase_dev(hd.h, C.c_void_p(d.data_ptr()), 2000), hd.h)
    H, nrep, ctr = C.c_int64(), C.c_int64(), _lib.PhaseCounters()
    _lib.check(lib.ms_phase_haplotypes(hd.h, 10, None, None, 0, C.byref(H), C.byref(nrep), C.byref(ctr), None), hd.h)
    ob, of = oracle.phase_bits(st, cols, cods)
    g = oracle.phase_group(ob, of, 40)
    assert (H.value, nrep.value) == (g["H"], g["nreported"])
    assert {k: getattr(ctr, k) for k, _ in _lib.PhaseCounters._fields_} == {k: int(v) for k, v in g["counters"].items()}
    # a different threshold on the same grouping (cached table): juliet's 10 is a parameter here
    _lib.check(lib.ms_phase_haplotypes(hd.h, 1, None, None, 0, C.byref(H), C.byref(nrep), C.byref(ctr), None), hd.h)
    assert nrep.value == g["H"] and ctr.insufficient == 0 and ctr.reported == 2000 - ctr.damaged
    assert lib.ms_phase_haplotypes(hd.h, -1, None, None, 0, C.byref(H), C.byref(nrep), C.byref(ctr), None) == -1
    assert lib.ms_phase_haplotypes(hd.h, 10, None, None, 5, C.byref(H), C.byref(nrep), C.byref(ctr), None) == -1


def test_fuse_consensus(oracle, hd):
    cfg = SynthConfig(L=1000, seed=91, dele=3e-3)
    t = make_tables(cfg)
    st = synth_states(t, 0, 3000)
    st[:, 100:104] = np.where(np.arange(3000)[:, None] % 10 < 7, np.uint8(4), st[:, 100:104])   # a major deletion
    st[:30, 500:520] = 7
    # insertion events: an in-frame majority insertion, a frame-shifting one, a minority one, one too close
    ev = []
    for r in range(3000):
        if r % 10 < 8: ev.append((r, 200, "GGA"))
        if r % 10 < 9: ev.append((r, 300, "GG"))
        if r % 10 < 2: ev.append((r, 400, "TTTAAA"))
        if r % 10 < 7: ev.append((r, 210, "CCC"))
        if r % 10 < 6: ev.append((r, 700, "ACGTAC"))
        if r % 10 >= 6: ev.append((r, 700, "ACGTAG"))
    for (r, c, s) in ev:
        st[r, c] |= 8
    pool = "".join(s for (_, _, s) in ev).encode()
    ic = [c for (_, c, _) in ev]
    il = [len(s) for (_, _, s) in ev]
    io = np.concatenate([[0], np.cumsum(il)[:-1]])
    f = Fuse(1000, handle=hd)
    d = to_dev(pack_states(st))
    f.pileup_device(d.data_ptr(), 3000)
    ocol, _ = oracle.pileup(st, None, codons=False)
    want = oracle.fuse(ocol, ic, io, il, pool)
    got = f.consensus(ic, io, il, pool)
    assert got == want
    assert len(got) == 1000 - 4 + 3 + 6 and "GGA" in got
    assert f.consensus() == oracle.fuse(ocol)
    f.params.min_coverage = 2990
    assert f.consensus() == oracle.fuse(ocol, min_coverage=2990)


def test_full_size_properties(hd):
    """1M x 3 kb (the metric's size): size-independent invariants instead of an oracle run."""
    cfg = SynthConfig(L=3000, seed=20240002)
    t = make_tables(cfg)
    R = 1_000_000
    d = gpu_synth(hd, t, 0, R)
    j = Juliet(3000, [(1, 3001)], handle=hd)
    j.pileup_device(d.data_ptr(), R)
    col, codon = j.get_counts()
    # column-sum invariant (SURVEY C-1): A+C+G+T+-+N == reads spanning the column
    _, begin, end = read_strains(t, 0, R)
    span = np.zeros(3001, dtype=np.int64)
    np.add.at(span, begin, 1)
    np.add.at(span, end, -1)
    span = np.cumsum(span)[:3000]
    assert np.array_equal(col[:, :6].sum(axis=1), span) and np.array_equal(col[:, 7], span)
    # coverage <= min ACGT count of the codon's columns (C-2)
    acgt = col[:, :4].sum(axis=1)
    cov = codon.sum(axis=1)
    for s in range(0, 2998, 3):
        assert cov[s] <= min(acgt[s], acgt[s + 1], acgt[s + 2])
        assert cov[s + 1] == 0 and cov[s + 2] == 0
    # marginalising the codon histogram over two positions reproduces ... the base counts among clean codons:
    # first-base marginal can never exceed the column's base count
    first = codon.reshape(3000, 4, 16).sum(axis=2)
    assert (first[::3] <= col[::3, :4]).all()
    # shard-and-sum == one shot (what the multi-GPU all-reduce relies on)
    j.reset()
    row = j.row_words * 4
    j.pileup_device(d.data_ptr(), 400_000)
    j.pileup_device(d.data_ptr() + 400_000 * row, 600_000)
    col2, codon2 = j.get_counts()
    assert np.array_equal(col, col2) and np.array_equal(codon, codon2)
    # the planted minors are called at their mixture frequencies
    found = {(v.col, v.codon): v.count / v.coverage for v in j.call()}
    fr = dict(zip(range(1, 4), cfg.minor_fracs))
    for strain, c, k in t.truth:
        assert abs(found[(c, k)] / fr[strain] - 1) < 0.1


# ---- cleric's alignment step: GPU Needleman-Wunsch against the restatement
def _mutate(rng, a, indel=0.02, sub=0.04):
    b = []
    for ch in a:
        u = rng.random()
        if u < indel:
            continue
        if u < 2 * indel:
            b.append(str(rng.choice(list("ACGT"))))
        b.append(ch if rng.random() > sub else str(rng.choice(list("ACGT"))))
    return "".join(b)


def gpu_nw(hd, a, b):
    lib = _lib.load()
    cap = len(a) + len(b) + 1
    ops = C.create_string_buffer(cap)
    n, score = C.c_int64(), C.c_int64()
    _lib.check(lib.ms_align_refs(hd.h, a.encode(), len(a), b.encode(), len(b), ops, cap, C.byref(n), C.byref(score)), hd.h)
    return ops.raw[: n.value].decode(), score.value


@pytest.mark.parametrize("la", [0, 1, 5, 127, 128, 129, 255, 257, 700, 3000])
def test_nw_align_equals_oracle(oracle, hd, la):
    """Same path (so: same tie-breaking in every cell) and same score as the CPU restatement, for lengths around the
    128-cell tile edges, unrelated sequences, and references that differ by indels and substitutions."""
    rng = np.random.default_rng(la)
    a = "".join(rng.choice(list("ACGT"), size=la))
    for b in (_mutate(rng, a), a, "".join(rng.choice(list("ACGT"), size=max(0, la - 3))), a[: la // 2], "ACGT" * 40):
        assert gpu_nw(hd, a, b) == oracle.nw_align(a, b)
        assert gpu_nw(hd, b, a) == oracle.nw_align(b, a)
    lib = _lib.load()
    n = C.c_int64()
    assert lib.ms_align_refs(hd.h, b"ACGT", 4, b"ACG", 3, C.create_string_buffer(3), 3, C.byref(n), None) == -4 and n.value == 7


@pytest.mark.parametrize("nminor,frames,nvar,copies,planned", [(3, 1, 5, 1, True), (2, 3, 4, 1, True), (3, 1, 5, 2, True), (7, 3, 5, 1, False),
                                                               (9, 1, 5, 1, False), (3, 1, 5, 5, False)],
                         ids=["few-1frame", "few-3frames", "same-gene-twice", "over32keys-3frames", "over32keys-1frame", "over64calls"])
def test_pass_device_planner_equals_host_planned_phasing(oracle, hd, nminor, frames, nvar, copies, planned):
    """The single-call pass builds the phasing plan on the device (phase_plan_kernel: pooled keys, touched blocks, block / layer /
    variant stream) and falls back to the host-built plan beyond 32 keys: same variants, keys, haplotypes, read ids as the staged
    three-call path, and the oracle's bit-vectors.  Overlapping frames give calls that share columns (several layers per block)
    and the same codon called in more than one gene (duplicates to pool)."""
    L, R = 1500, 30_000
    cfg = SynthConfig(L=L, seed=4100 + nminor, minor_fracs=tuple([0.08] * nminor), variants_per_minor=(nvar, nvar), n_rate=2e-3)
    t = make_tables(cfg)
    st = synth_states(t, 0, R)
    genes = [(f + 1, L + 1) for f in range(frames)] * copies
    d = to_dev(pack_states(st))
    j = Juliet(L, genes, refseq=t.refseq, mode_phasing=True, min_perc=1.0, handle=hd)
    a = j.run_device(d.data_ptr(), R, want_hap_id=True)              # planner (or its fallback)
    b = j._run_staged(d.data_ptr(), R, True)                         # ms_call + host-built plan + ms_phase_haplotypes
    key = lambda v: (v.gene, v.col, v.codon, v.count, v.coverage, v.pvalue)
    assert [key(v) for v in a.variants] == [key(v) for v in b.variants]
    assert a.keys == b.keys == sorted({(v.col, v.codon) for v in a.variants})
    assert (len(a.keys) <= 32 and len(a.variants) <= 64) == planned, (len(a.keys), len(a.variants))   # device planner, or its fallback
    ha, hb = a.haplotypes, b.haplotypes
    assert ha.ndistinct == hb.ndistinct and ha.nreported == hb.nreported and ha.counters == hb.counters
    k = min(len(ha.counts), len(hb.counts))
    assert np.array_equal(ha.counts[:k], hb.counts[:k]) and np.array_equal(ha.patterns[:k], hb.patterns[:k])
    assert np.array_equal(ha.hap_id, hb.hap_id)
    obits, oflags = oracle.phase_bits(st, [c for c, _ in a.keys], [k2 for _, k2 in a.keys], nthreads=8)
    g = oracle.phase_group(obits, oflags, len(a.keys))
    assert np.array_equal(ha.hap_id, g["hap_id"]) and ha.counters == {kk: int(v) for kk, v in g["counters"].items()}


def test_phase_bits_unsorted_and_shared_site_variants(oracle, hd):
    """Variant lists the packed (bit-compress) form of phase_bits_kernel's stream cannot take: shuffled order, several codons
    at one site (F18: one position may carry several variant codons), overlapping frames, a duplicate entry -- next to the
    sorted dense list it is made for.  Bits and flags word for word against the oracle."""
    L, R = 960, 5000
    cfg = SynthConfig(L=L, seed=77, dense_sites=120, dense_strains=8, n_rate=5e-3, dele=2e-3, trunc=0.05)
    t = make_tables(cfg)
    st = synth_states(t, 0, R)
    d = to_dev(pack_states(st))
    truth = sorted({(c, k) for (_, c, k) in t.truth})
    rng = np.random.default_rng(5)
    extra = [(c, (k + 1) % 64) for (c, k) in truth[::3]] + [(c + 1, 7) for (c, _) in truth[::5] if c + 4 < L] + [truth[0]]
    for name, keys in (("sorted-dense", truth), ("shuffled", [truth[i] for i in rng.permutation(len(truth))]),
                       ("shared-sites", sorted(truth + extra)), ("shuffled-shared", [(truth + extra)[i] for i in rng.permutation(len(truth + extra))])):
        j = Juliet(L, [(1, L + 1)], mode_phasing=True, handle=hd)

        class V:
            def __init__(self, c, k):
                self.col, self.codon = c, k
        cols = np.array([c for c, _ in keys], dtype=np.int32)
        cods = np.array([k for _, k in keys], dtype=np.int32)
        _lib.check(j.lib.ms_phase_begin(hd.h, cols.ctypes.data_as(C.c_void_p), cods.ctypes.data_as(C.c_void_p), len(keys), R), hd.h)
        _lib.check(j.lib.ms_phase_dev(hd.h, C.c_void_p(d.data_ptr()), R), hd.h)
        from minorseq_b200.api import _as_tensor
        nw = (len(keys) + 31) // 32
        pb, pf, pn = C.c_void_p(), C.c_void_p(), C.c_int64()
        _lib.check(j.lib.ms_phase_device(hd.h, C.byref(pb), C.byref(pf), C.byref(pn)), hd.h)
        _lib.check(j.lib.ms_synchronize(hd.h), hd.h)
        gbits = _as_tensor(pb.value, (R * nw,), torch.int32, 0).cpu().numpy().view(np.uint32).reshape(R, nw)
        gflags = _as_tensor(pf.value, (R,), torch.uint8, 0).cpu().numpy()
        obits, oflags = oracle.phase_bits(st, [int(c) for c in cols], [int(k) for k in cods])
        assert np.array_equal(gflags, oflags), name
        assert np.array_equal(gbits, obits), name


def test_pileup_mid_kernel_flush_with_shared_memory_planes(hd):
    """More than 8191 reads per row-group: the instantiation with the two shared-memory counter planes (HI) has to flush in the
    middle of the kernel.  Checked against the same reads piled up in launches small enough to need neither the planes nor a
    flush (that path is compared with the oracle in the tests above)."""
    cfg = SynthConfig(L=96, seed=32, variants_per_minor=(1, 1), minor_fracs=(0.05,))
    t = make_tables(cfg)
    R = 148 * 12 * 8191 + 40_003
    d = gpu_synth(hd, t, 0, R)
    j = Juliet(96, [(1, 97), (2, 97)], handle=hd)
    j.pileup_device(d.data_ptr(), R)
    col1, codon1 = j.get_counts()
    j.reset()
    step = 148 * 12 * 2000            # < 2048 reads per row-group: plain instantiation, no flush
    row_bytes = j.row_words * 4
    for r0 in range(0, R, step):
        j.pileup_device(d.data_ptr() + r0 * row_bytes, min(step, R - r0))
    col2, codon2 = j.get_counts()
    assert int(col1[:, 7].max()) > 14_000_000 * 0.9
    assert np.array_equal(col1, col2) and np.array_equal(codon1, codon2)
