"""Randomised end-to-end parity: the single-call juliet pass (pile-up -> call -> device-planned phasing) against the oracle on
shapes nobody picked by hand -- reference lengths 3 ... 12500 (whole rows, column segments, counter planes in shared memory,
logged and in-kernel rare path, DENSE), random gene layouts over one to three reading frames, ragged read counts, spans and
noise levels -- and the same reads again through the event-row entry point.  Fixed seeds, small cases (the oracle is the clock)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from minorseq_b200 import Handle, Juliet, device_rows, encode_states  # noqa: E402
from minorseq_b200.synth import SynthConfig, make_tables, pack_states, synth_states  # noqa: E402
from test_gpu_parity import mask_bytes, variants_equal  # noqa: E402


@pytest.fixture(scope="module")
def hd():
    h = Handle(0)
    yield h
    h.close()


def _random_case(seed):
    rng = np.random.default_rng(seed)
    L = int(rng.choice([rng.integers(9, 200), rng.integers(200, 3500), rng.integers(3500, 12500)]))
    R = int(rng.choice([rng.integers(0, 40), rng.integers(40, 1500), rng.integers(1500, 4000)])) if L < 6000 else int(rng.integers(0, 900))
    dense = rng.random() < 0.25 and L >= 300
    nminor = int(rng.integers(0, 4))
    cfg = SynthConfig(L=L, seed=int(seed) * 7 + 1, minor_fracs=tuple(float(x) for x in rng.uniform(0.03, 0.15, size=nminor)),
                      variants_per_minor=(1, 4), sub=float(rng.choice([5e-4, 5e-3])), dele=float(rng.choice([3e-3, 3e-2])),
                      ins=float(rng.choice([0.0, 1e-3])), n_rate=float(rng.choice([0.0, 2e-2, 0.1])), trunc=float(rng.choice([0.0, 0.02, 0.5])),
                      dense_sites=int(min(L // 6, rng.integers(20, 200))) if dense else 0, dense_strains=int(rng.integers(4, 20)),
                      frame=int(rng.integers(0, 3)) if L > 12 else 0)
    # genes: one to four intervals in random frames, possibly overlapping, possibly beyond the reference end
    genes = []
    for _ in range(int(rng.integers(1, 5))):
        b = int(rng.integers(1, max(2, L - 6)))
        e = int(min(L + 1 + rng.integers(0, 3), b + 3 * rng.integers(1, max(2, (L - b) // 3 + 1))))
        if e - b >= 3:
            genes.append((b, e))
    if not genes:
        genes = [(1, L - L % 3 + 1)]
    return cfg, R, genes


@pytest.mark.parametrize("seed", list(range(100, 196)))
def test_random_shapes_pass_equals_oracle(oracle, hd, seed):
    cfg, R, genes = _random_case(seed)
    L = cfg.L
    t = make_tables(cfg)
    st = synth_states(t, 0, R) if R else np.zeros((0, L), dtype=np.uint8)
    refseq = t.refseq if seed % 3 else None                     # reference-guided or against the major codon
    min_perc = 2.0
    j = Juliet(L, genes, refseq=refseq, mode_phasing=True, min_perc=min_perc, handle=hd)
    d = device_rows(pack_states(st)) if R else device_rows(np.zeros((0, 4 * ((L + 31) // 32)), dtype=np.uint32))
    res = j.run_device(d.data_ptr(), R, want_hap_id=True)
    col, codon = j.get_counts()
    ocol, ocodon = (oracle.pileup(st, mask_bytes(j.start_mask, L)) if R else
                    (np.zeros((L, 8), np.uint32), np.zeros((L, 64), np.uint32)))
    ocol = ocol.copy()
    ocol[:, 6] = 0
    assert np.array_equal(col, ocol), (seed, L, R, genes)
    assert np.array_equal(codon, ocodon), (seed, L, R, genes)
    ov = oracle.call(ocodon, genes, refseq=refseq, min_perc=min_perc)
    variants_equal(res.variants, ov)
    keys = sorted({(v.col, v.codon) for v in ov})
    assert res.keys == keys
    hp = res.haplotypes
    if R:
        obits, oflags = oracle.phase_bits(st, [c for c, _ in keys], [k for _, k in keys])
        g = oracle.phase_group(obits, oflags, len(keys))
        assert hp.counters == {kk: int(v) for kk, v in g["counters"].items()}, (seed, L, R)
        assert hp.ndistinct == g["H"] and hp.nreported == g["nreported"]
        k = len(hp.counts)
        assert np.array_equal(hp.counts, g["counts"][:k]) and np.array_equal(hp.patterns, g["patterns"][:k])
        assert np.array_equal(hp.hap_id, g["hap_id"])
        # the same reads as event rows from the host: same counts, calls and read ids
        if seed % 2 == 0:
            hdr, ev = encode_states(st, t.strain_base[0])
            j.set_base(t.strain_base[0])
            res2 = j.run_events_host(hdr, ev, want_hap_id=True)
            col2, codon2 = j.get_counts()
            assert np.array_equal(col2, ocol) and np.array_equal(codon2, ocodon)
            variants_equal(res2.variants, ov)
            assert np.array_equal(res2.haplotypes.hap_id, g["hap_id"])
    else:
        assert len(res.variants) == 0 and hp.nreported == 0
