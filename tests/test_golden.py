"""Committed golden vectors (tests/golden/small_case.npz, made by tests/golden/make_golden.py).
CPU: the oracle still reproduces them.  GPU: the CUDA path reproduces them through the C ABI."""
import os

import numpy as np
import pytest

from minorseq_b200.synth import unpack_states

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "small_case.npz"))
L = int(G["L"])
GENES = [tuple(int(x) for x in g) for g in G["genes"]]
REF = G["refseq"].tobytes().decode()


def _mask():
    m = np.zeros(L, dtype=np.uint8)
    for (b, e) in GENES:
        m[b - 1: min(e - 1, L) - 2: 3] = 1
    return m


def test_oracle_reproduces_golden(oracle):
    st = unpack_states(G["packed"], L)
    col, codon = oracle.pileup(st, _mask())
    assert np.array_equal(col, G["col"]) and np.array_equal(codon, G["codon"])
    v = oracle.call(codon, GENES, refseq=REF)
    got = np.array([(x.gene, x.codon_index, x.col, x.ref_codon, x.codon, x.count, x.coverage, x.expected, x.ntests) for x in v], dtype=np.int64)
    assert np.array_equal(got, G["variants"])
    assert np.allclose([x.pvalue for x in v], G["pvalues"], rtol=1e-12, atol=0)
    keys = [tuple(k) for k in G["keys"]]
    bits, flags = oracle.phase_bits(st, [k[0] for k in keys], [k[1] for k in keys])
    g = oracle.phase_group(bits, flags, len(keys))
    assert np.array_equal(g["counts"], G["hap_counts"]) and np.array_equal(g["patterns"], G["hap_patterns"])
    assert np.array_equal(g["hap_id"], G["hap_id"]) and g["nreported"] == int(G["nreported"])
    ocol, _ = oracle.pileup(st, None, codons=False)
    assert oracle.fuse(ocol) == G["consensus"].tobytes().decode()


@pytest.mark.gpu
def test_gpu_reproduces_golden():
    import torch
    from minorseq_b200 import Fuse, Juliet, device_rows
    d = device_rows(G["packed"])
    R = G["packed"].shape[0]
    j = Juliet(L, GENES, refseq=REF, mode_phasing=True)
    j.set_count_insertions(True)
    res = j.run_device(d.data_ptr(), R, want_hap_id=True)
    col, codon = j.get_counts()
    assert np.array_equal(col, G["col"]) and np.array_equal(codon, G["codon"])
    got = np.array([(x.gene, x.codon_index, x.col, x.ref_codon, x.codon, x.count, x.coverage, x.expected, x.ntests) for x in res.variants], dtype=np.int64)
    assert np.array_equal(got, G["variants"])
    for a, b in zip([x.pvalue for x in res.variants], G["pvalues"]):
        assert abs(a - b) <= 1e-9 * b
    h = res.haplotypes
    assert np.array_equal(h.counts, G["hap_counts"]) and np.array_equal(h.patterns, G["hap_patterns"]) and np.array_equal(h.hap_id, G["hap_id"])
    assert [h.counters[k] for k in ("reported", "insufficient", "damaged", "gaps", "heteroduplex", "partial")] == [int(x) for x in G["counters"]]
    f = Fuse(L, handle=j.hd)
    f.pileup_device(d.data_ptr(), R)
    assert f.consensus() == G["consensus"].tobytes().decode()
