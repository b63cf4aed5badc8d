"""Property tests of the CPU restatement (hypothesis): the invariants read off the reference's screenshots
 must hold for arbitrary alignments, not just the synthetic generator's."""
import numpy as np
from hypothesis import given, settings, strategies as st

STATES = np.array([0, 1, 2, 3, 4, 5, 7], dtype=np.uint8)


@st.composite
def alignments(draw):
    R = draw(st.integers(1, 60))
    L = draw(st.integers(3, 70))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    p = rng.dirichlet(np.ones(7))
    s = rng.choice(STATES, size=(R, L), p=p)
    s = s | (rng.integers(0, 2, size=(R, L), dtype=np.uint8) << 3)
    s = np.where((s & 7) == 7, np.uint8(7), s).astype(np.uint8)
    frame = draw(st.integers(0, 2))
    mask = np.zeros(L, dtype=np.uint8)
    mask[frame:max(frame, L - 2):3] = 1
    return s, mask, seed


@settings(max_examples=60, deadline=None)
@given(alignments())
def test_pileup_invariants(oracle, a):
    s, mask, seed = a
    R, L = s.shape
    col, codon = oracle.pileup(s, mask)
    span = ((s & 7) != 7).sum(axis=0)
    assert np.array_equal(col[:, :6].sum(axis=1), span) and np.array_equal(col[:, 7], span)       # C-1
    acgt = col[:, :4].sum(axis=1)
    cov = codon.sum(axis=1)
    for j in range(L - 2):
        if mask[j]:
            assert cov[j] <= min(acgt[j], acgt[j + 1], acgt[j + 2])                                 # C-2
        else:
            assert cov[j] == 0
    # permutation of the reads changes nothing; sharding and summing changes nothing
    perm = np.random.default_rng(seed).permutation(R)
    col2, codon2 = oracle.pileup(s[perm], mask)
    assert np.array_equal(col, col2) and np.array_equal(codon, codon2)
    k = R // 2
    ca, da = oracle.pileup(s[:k], mask) if k else (np.zeros_like(col), np.zeros_like(codon))
    cb, db = oracle.pileup(s[k:], mask)
    assert np.array_equal(ca + cb, col) and np.array_equal(da + db, codon)
    # multi-threaded == single-threaded
    col3, codon3 = oracle.pileup(s, mask, nthreads=3)
    assert np.array_equal(col, col3) and np.array_equal(codon, codon3)


@settings(max_examples=40, deadline=None)
@given(alignments(), st.integers(0, 40))
def test_phase_partition(oracle, a, V):
    s, mask, seed = a
    R, L = s.shape
    rng = np.random.default_rng(seed + 1)
    cols = rng.integers(0, L, size=V)          # may point past L-3: such variants make every read "partial"
    cods = rng.integers(0, 64, size=V)
    bits, flags = oracle.phase_bits(s, cols, cods)
    g = oracle.phase_group(bits, flags, V, min_reads=3)
    c = g["counters"]
    assert c["reported"] + c["insufficient"] + c["damaged"] == R                                    # C-4
    assert c["gaps"] + c["heteroduplex"] + c["partial"] >= c["damaged"]
    assert int(g["counts"].sum()) == R - c["damaged"]
    assert (np.diff(g["counts"][: g["nreported"]].astype(np.int64)) <= 0).all()                     # descending
    assert ((g["hap_id"] == -1) == (flags != 0)).all()
    if V:
        Cm = oracle.cooccurrence(bits, V)
        assert np.array_equal(Cm, Cm.T) and (np.diag(Cm) == np.array([(bits[:, v >> 5] >> (v & 31) & 1).sum() for v in range(V)])).all()
