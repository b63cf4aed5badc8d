"""Event rows (csrc/events.cu): the compact host->device form of an aligned read.

CPU part: the host encoders against an independent numpy decoder (minorseq_b200.api.decode_events) on generator reads,
random states, skips (> 254 unchanged columns between two entries of a list), unspanned reads, reference skips inside a read.
GPU part: expand_events_kernel rebuilds exactly the rows ms_pack_states / ms_expand_cigar would have produced (as device tiles), and
the pass from event rows gives the pass from rows (counts, variants, haplotypes, read ids) and the oracle's.
"""
import ctypes as C

import numpy as np
import pytest

from minorseq_b200 import _lib, decode_events, encode_rows, encode_states, host_rows, tile_rows
from minorseq_b200.api import HDR_DTYPE
from minorseq_b200.synth import SynthConfig, make_tables, pack_states, synth_states


def _random_states(rng, R, L, p_event):
    base = rng.integers(0, 4, size=L, dtype=np.uint8)
    st = np.tile(base, (R, 1))
    ev = rng.random((R, L)) < p_event
    st[ev] = rng.choice(np.array([0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 11, 12, 13, 15], dtype=np.uint8), size=int(ev.sum()))
    for r in range(R):       # ragged spans, some reads empty
        b, e = sorted(rng.integers(0, L + 1, size=2))
        if r % 7 == 0:
            b = e
        st[r, :b] = 7
        st[r, e:] = 7
    return base, st


@pytest.mark.parametrize("L", [3, 31, 32, 33, 100, 3000, 9719])
def test_encode_decode_roundtrip_random(L):
    rng = np.random.default_rng(L)
    base, st = _random_states(rng, 60, L, 0.03)
    hdr, ev = encode_states(st, base)
    assert hdr.dtype == HDR_DTYPE and len(hdr) == 61 and int(hdr["ev_off"][-1]) == len(ev)
    assert np.array_equal(decode_events(hdr, ev, L, base), st)
    hdr2, ev2 = encode_rows(pack_states(st), L, base)
    assert np.array_equal(hdr2, hdr) and np.array_equal(ev2, ev)


def test_encode_generator_reads_are_compact():
    """CCS-like reads against the major strain: ~85 events (N 2 %, deletions, substitutions, insertion flags) per 3 kb read,
    i.e. ~112 B (one byte per QV-filtered base, 12 bits per other event, a 2-byte count and an 8-byte header) instead of the
    1504-byte planar row."""
    t = make_tables(SynthConfig(L=3000, seed=20240003))
    st = synth_states(t, 0, 500)
    hdr, ev = encode_states(st, t.refseq)
    assert np.array_equal(decode_events(hdr, ev, 3000, t.refseq), st)
    per_read = (len(ev) + 8 * len(hdr)) / 500
    assert 90 < per_read < 130, per_read
    # lossless for ANY base: a wrong base only makes the list longer
    rng = np.random.default_rng(1)
    other = rng.integers(0, 4, size=3000, dtype=np.uint8)
    hdr2, ev2 = encode_states(st[:50], other)
    assert np.array_equal(decode_events(hdr2, ev2, 3000, other), st[:50]) and len(ev2) > 50 * 2000 * 3 // 2


def test_encode_skips_and_edges():
    L = 20000
    base = np.zeros(L, dtype=np.uint8)
    st = np.tile(base, (8, 1))
    st[0, 19999] = 2                      # one event 19999 columns after begin: 78 skips (19999 = 78 * 255 + 109)
    st[1, 254] = 1                        # exactly the largest delta: no skip
    st[2, 255] = 1                        # one more: one skip, then delta 0
    st[3, :] = 7                          # spans nothing
    st[4, :100] = 7; st[4, 150:] = 7; st[4, 120:125] = 7     # reference skip (CIGAR N) inside the read
    st[5, 0] = 15; st[5, 1] = 8           # insertion flags on an unspanned-looking and on an unchanged column
    st[6, 600] = 5                        # one N 600 columns after begin: 2 skips + 1 entry in the byte list, no rest list
    st[7, 0] = 5; st[7, 1] = 13; st[7, 2] = 5; st[7, 700] = 5; st[7, 701] = 1   # both lists, N with insertion flag goes to the rest list
    hdr, ev = encode_states(st, base)
    nbytes = np.diff(hdr["ev_off"].astype(np.int64))
    off = hdr["ev_off"].astype(np.int64)
    nN = np.array([int(ev[o]) | (int(ev[o + 1]) << 8) if n >= 2 else 0 for o, n in zip(off[:-1], nbytes)])
    nrest = np.where(nbytes >= 2, (2 * (nbytes - 2 - nN)) // 3, 0)      # ceil(1.5 n) bytes hold n 12-bit entries
    assert list(nN) == [0, 0, 0, 0, 0, 0, 3, 5] and list(nrest) == [79, 1, 2, 0, 5, 2, 0, 4]
    assert list(nbytes) == [121, 4, 5, 0, 10, 5, 5, 13]
    assert (int(hdr["begin"][3]), int(hdr["end"][3])) == (0, 0)
    assert (int(hdr["begin"][4]), int(hdr["end"][4])) == (100, 150)
    assert np.array_equal(decode_events(hdr, ev, L, base), st)
    # a read that equals the base on its span costs no bytes at all
    st2 = np.tile(base, (2, 1)); st2[1, :50] = 7
    hdr2, ev2 = encode_states(st2, base)
    assert len(ev2) == 0 and (int(hdr2["begin"][1]), int(hdr2["end"][1])) == (50, L)
    assert np.array_equal(decode_events(hdr2, ev2, L, base), st2)


def test_encode_all_n_read():
    """The longest N list: every column of a 65,535-column read is 'N' (the u16 count still holds)."""
    L = 65535
    base = np.zeros(L, dtype=np.uint8)
    st = np.full((1, L), 5, dtype=np.uint8)
    hdr, ev = encode_states(st, base)
    assert len(ev) == 2 + L and (int(ev[0]) | (int(ev[1]) << 8)) == L
    assert np.array_equal(decode_events(hdr, ev, L, base), st)


@pytest.mark.parametrize("L", [1, 2, 254, 255, 256, 511, 3000, 65535])
def test_events_bound_holds_for_worst_rows(mslib, L):
    """ms_events_bound(L) covers the longest strings: every column in the rest list, every column an N, one event per 255 columns
    (a skip before each), alternating lists."""
    base = np.zeros(L, dtype=np.uint8)
    rows = []
    rows.append(np.full(L, 1, dtype=np.uint8))                       # every column a substitution
    rows.append(np.full(L, 13, dtype=np.uint8))                      # ... with insertion flags
    rows.append(np.full(L, 5, dtype=np.uint8))                       # every column N
    alt = np.full(L, 5, dtype=np.uint8); alt[::2] = 2; rows.append(alt)
    sparse = base.copy(); sparse[254::255] = 3; rows.append(sparse)  # one event after every skip-sized gap
    sparse2 = base.copy(); sparse2[255::256] = 5; rows.append(sparse2)
    st = np.stack(rows)
    hdr, ev = encode_states(st, base)
    nbytes = np.diff(hdr["ev_off"].astype(np.int64))
    assert nbytes.max() <= mslib.ms_events_bound(L), (nbytes, mslib.ms_events_bound(L))
    assert np.array_equal(decode_events(hdr, ev, L, base), st)


def test_encode_decode_roundtrip_hypothesis():
    """Arbitrary short rows over every legal nibble (6 and 14 are reserved), arbitrary base: encode -> decode is the identity and
    rows and states encoders agree."""
    from hypothesis import given, settings, strategies as hs
    legal = [0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 11, 12, 13, 15]

    @settings(max_examples=150, deadline=None)
    @given(hs.integers(1, 700).flatmap(lambda L: hs.tuples(
        hs.lists(hs.sampled_from(legal), min_size=L, max_size=L),
        hs.lists(hs.integers(0, 3), min_size=L, max_size=L),
        hs.integers(0, L), hs.integers(0, L))))
    def check(args):
        nib, base, a, b = args
        L = len(nib)
        st = np.array([nib], dtype=np.uint8)
        lo, hi = min(a, b), max(a, b)
        st[0, :lo] = 7
        st[0, hi:] = 7
        base = np.array(base, dtype=np.uint8)
        hdr, ev = encode_states(st, base)
        assert np.array_equal(decode_events(hdr, ev, L, base), st)
        hdr2, ev2 = encode_rows(pack_states(st), L, base)
        assert np.array_equal(hdr2, hdr) and np.array_equal(ev2, ev)

    check()


def test_encode_errors(mslib):
    base = np.zeros(100, dtype=np.uint8)
    st = np.full((3, 100), 1, dtype=np.uint8)
    hdr = np.zeros(4, dtype=HDR_DTYPE)
    ev = np.zeros(10, dtype=np.uint8)
    n = C.c_int64()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert mslib.ms_encode_states(p(st), 3, 100, p(base), p(hdr), p(ev), 10, C.byref(n)) == -4      # MS_ERR_CAPACITY
    assert mslib.ms_encode_states(p(st), 3, 70000, p(base), p(hdr), p(ev), 10, C.byref(n)) == -1    # L > 65535
    st[1, 5] = 6
    big = np.zeros(600, dtype=np.uint8)
    assert mslib.ms_encode_states(p(st), 3, 100, p(base), p(hdr), p(big), 600, C.byref(n)) == -5    # reserved state
    assert mslib.ms_events_bound(3000) >= 2 + (3000 + 3000 // 255 + 1) * 3 // 2 + 1 and mslib.ms_events_bound(3000) < 4600


def test_encode_row_incremental_matches_batch(mslib):
    """The per-record form the BAM loop uses (ms_base_planes + ms_encode_row + ms_events_seal) == the batch encoder."""
    rng = np.random.default_rng(3)
    base, st = _random_states(rng, 40, 777, 0.05)
    rows = pack_states(st)
    want_hdr, want_ev = encode_states(st, base)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    planes = np.zeros(2 * ((777 + 31) // 32), dtype=np.uint32)
    assert mslib.ms_base_planes(p(base), 777, p(planes)) == 0
    hdr = np.zeros(41, dtype=HDR_DTYPE)
    ev = np.zeros(len(want_ev) + 8, dtype=np.uint8)
    n = C.c_int64(0)
    for r in range(40):
        assert mslib.ms_encode_row(p(rows[r]), 777, p(planes), C.c_void_p(hdr.ctypes.data + 8 * r), p(ev), len(ev), C.byref(n)) == 0
    assert mslib.ms_events_seal(p(hdr), 40, n.value, p(base), 777) == 0
    assert np.array_equal(hdr, want_hdr) and np.array_equal(ev[: n.value], want_ev)


# ------------------------------------------------------------------------------------------------ GPU
gpu = pytest.mark.gpu


def _dev(a, torch):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8)).cuda()


@gpu
@pytest.mark.parametrize("L,R", [(3, 5), (32, 40), (33, 100), (97, 1000), (3000, 3000), (6144, 300), (9719, 700), (20000, 200), (40000, 61),
                                 (65535, 20)])   # the last two: rows too long for eight of them in shared memory (per-warp stores)
def test_expand_events_equals_packed_rows(L, R):
    import torch
    from minorseq_b200 import Handle, Juliet
    rng = np.random.default_rng(L + R)
    base, st = _random_states(rng, R, L, 0.03 if L < 10000 else 0.002)
    hdr, ev = encode_states(st, base)
    hd = Handle(0)
    try:
        j = Juliet(L, [(1, L + 1)], handle=hd)
        lib = j.lib
        dh, de = _dev(hdr, torch), _dev(ev if len(ev) else np.zeros(1, np.uint8), torch)
        out = torch.zeros(int(lib.ms_tiled_words(L, R)), dtype=torch.int32, device="cuda")
        assert lib.ms_expand_events_dev(hd.h, C.c_void_p(dh.data_ptr()), C.c_void_p(de.data_ptr()), R, C.c_void_p(out.data_ptr())) == -1  # no base yet
        j.set_base(base)
        _lib.check(lib.ms_expand_events_dev(hd.h, C.c_void_p(dh.data_ptr()), C.c_void_p(de.data_ptr()), R, C.c_void_p(out.data_ptr())), hd.h)
        _lib.check(lib.ms_synchronize(hd.h), hd.h)
        assert np.array_equal(host_rows(out, R, L), pack_states(st))
        # the padding of the last tile is "not spanned", so whole tile buffers can be compared as well
        assert np.array_equal(out.cpu().numpy().view(np.uint32), tile_rows(pack_states(st)))
    finally:
        hd.close()


@gpu
def test_expand_corrupt_event_rows_stay_in_bounds():
    """Garbage event bytes and headers (counts past the string, columns past the row, offsets running backwards or past the array):
    the kernel completes, writes only its own tiles, and the untouched reads of the batch still expand exactly."""
    import torch
    from minorseq_b200 import Handle, Juliet
    L, R = 3000, 4096
    rng = np.random.default_rng(99)
    base, st = _random_states(rng, R, L, 0.03)
    hdr, ev = encode_states(st, base)
    hdr, ev = hdr.copy(), ev.copy()
    bad = rng.choice(R, size=600, replace=False)
    off = hdr["ev_off"].astype(np.int64)
    for r in bad[:400]:                                   # random bytes inside the read's own string
        ev[off[r]: off[r + 1]] = rng.integers(0, 256, size=int(off[r + 1] - off[r]), dtype=np.uint8)
    good = np.ones(R, dtype=bool)
    good[bad] = False
    for r in bad[400:500]:                                # offsets running backwards / past the array (neighbours see them as well)
        hdr["ev_off"][r] = rng.choice([0, 0xFFFFFFFF, len(ev) + 12345, int(off[r]) + 100000])
        good[max(0, r - 1)] = False
    for r in bad[500:]:                                   # spans past the reference
        hdr["begin"][r], hdr["end"][r] = rng.integers(0, 65536, size=2)
    hd = Handle(0)
    try:
        j = Juliet(L, [(1, L + 1)], handle=hd)
        j.set_base(base)
        lib = j.lib
        dh, de = _dev(hdr, torch), _dev(ev, torch)
        nwords = int(lib.ms_tiled_words(L, R))
        out = torch.full((nwords + 4096,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")      # guard words behind the tiles
        _lib.check(lib.ms_expand_events_dev(hd.h, C.c_void_p(dh.data_ptr()), C.c_void_p(de.data_ptr()), R, C.c_void_p(out.data_ptr())), hd.h)
        _lib.check(lib.ms_synchronize(hd.h), hd.h)
        assert bool((out[nwords:] == 0x5A5A5A5A).all())
        rows = host_rows(out[:nwords], R, L)
        assert np.array_equal(rows[good], pack_states(st)[good])
    finally:
        hd.close()


@gpu
def test_pass_from_event_rows_equals_pass_from_rows_and_oracle(oracle):
    """juliet --mode-phasing through ms_juliet_pass_events_host: same counts, variants, haplotypes and read ids as the pass from
    planar rows, and the oracle's counts.  Small chunk size so that the chunked upload pipeline has several chunks."""
    import torch
    from minorseq_b200 import Handle, Juliet
    cfg = SynthConfig(L=3000, seed=20240003, n_rate=5e-3)
    t = make_tables(cfg)
    R = 150_001
    st = synth_states(t, 0, R)
    packed = pack_states(st)
    hdr, ev = encode_rows(packed, 3000, t.refseq)
    hd = Handle(0)
    try:
        genes = [(1, 3001)]
        j = Juliet(3000, genes, refseq=t.refseq, mode_phasing=True, min_perc=0.5, handle=hd)
        a = j.run_host(packed, want_hap_id=True)
        col_a, codon_a = j.get_counts()
        with pytest.raises(_lib.MsError):
            j.run_events_host(hdr, ev)          # no base yet
        j.set_base(t.refseq)
        b = j.run_events_host(hdr, ev, want_hap_id=True)
        col_b, codon_b = j.get_counts()
        assert np.array_equal(col_a, col_b) and np.array_equal(codon_a, codon_b)
        mask = np.array([(int(j.start_mask[i >> 5]) >> (i & 31)) & 1 for i in range(3000)], dtype=np.uint8)
        ocol, ocodon = oracle.pileup(st, mask, nthreads=8)
        ocol[:, 6] = 0
        assert np.array_equal(col_b, ocol) and np.array_equal(codon_b, ocodon)
        assert [(v.col, v.codon, v.count, v.coverage, v.pvalue) for v in a.variants] == [(v.col, v.codon, v.count, v.coverage, v.pvalue) for v in b.variants]
        assert len(b.variants) >= 8 and a.keys == b.keys
        assert np.array_equal(a.haplotypes.patterns, b.haplotypes.patterns) and np.array_equal(a.haplotypes.counts, b.haplotypes.counts)
        assert a.haplotypes.counters == b.haplotypes.counters and np.array_equal(a.haplotypes.hap_id, b.haplotypes.hap_id)
        # rows encoded against another base are rejected, not mis-expanded
        other = np.roll(j.base, 1)
        hdr2, ev2 = encode_rows(packed[:100], 3000, other)
        with pytest.raises(_lib.MsError, match="different base"):
            j.run_events_host(hdr2, ev2)
        # offsets that do not ascend from chunk boundary to chunk boundary are refused before anything is copied
        hdr3 = hdr.copy()
        hdr3["ev_off"][0] = len(ev) + 5
        with pytest.raises(_lib.MsError, match="not ascending"):
            j.run_events_host(hdr3, ev)
        # zero reads
        hdr0, ev0 = encode_rows(packed[:0], 3000, t.refseq)
        z = j.run_events_host(hdr0, ev0)
        assert len(z.variants) == 0
    finally:
        hd.close()
