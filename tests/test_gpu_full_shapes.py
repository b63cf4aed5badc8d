"""GPU parity at the REAL shapes of BASELINE.json configs[3] and configs[4] (reduced read counts so that the oracle finishes):

* C5: L = 6144, V = 2048 dense variant sites, 64 strains -- K1's DENSE instantiation on column segments,
  phase_bits_kernel with 64-word bit-vectors, ~R distinct patterns through the device merge/order, and the co-occurrence
  matrix through the 36-tile x split-K tcgen05 launch;
* C4: L = 9719 with the 15-gene / three-frame HXB2-style layout (logged rare path, 10-warp rows), call + phase.

Everything is compared with the CPU oracle (counts, variants, bit-vectors, flags, haplotype order, read ids); the
co-occurrence matrix with an exact float32 GEMM of the oracle's bit matrix (all partial sums are integers < 2^24).
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from minorseq_b200 import Handle, Juliet, _lib, host_rows  # noqa: E402
from minorseq_b200.synth import SynthConfig, make_tables  # noqa: E402
from test_gpu_parity import gpu_synth, mask_bytes, variants_equal  # noqa: E402

HIV_GENES = [(1, 634), (790, 1186), (1186, 1879), (1879, 1921), (1921, 2086), (2086, 2134), (2134, 2292), (2253, 2550),
             (2550, 4230), (4230, 5096), (5041, 5620), (5559, 5850), (6062, 6310), (6225, 8795), (8797, 9417)]


@pytest.fixture(scope="module")
def hd():
    h = Handle(0)
    yield h
    h.close()


class _V:
    def __init__(self, c, k):
        self.col, self.codon = c, k


def _device_results(j, hd, R, nw):
    from minorseq_b200.api import _as_tensor
    pb, pf, pn = C.c_void_p(), C.c_void_p(), C.c_int64()
    _lib.check(j.lib.ms_phase_device(hd.h, C.byref(pb), C.byref(pf), C.byref(pn)), hd.h)
    assert pn.value == R
    bits = _as_tensor(pb.value, (R * nw,), torch.int32, 0).cpu().numpy().view(np.uint32).reshape(R, nw)
    flags = _as_tensor(pf.value, (R,), torch.uint8, 0).cpu().numpy()
    return bits, flags


def test_c5_shape_pileup_phase_cooccurrence(oracle, hd):
    L, R = 6144, 40_000
    t = make_tables(SynthConfig(L=L, seed=20240005, dense_sites=2048, dense_strains=64, n_rate=2e-5, dele=2e-5, trunc=0.0))
    d = gpu_synth(hd, t, 0, R)
    st = oracle.unpack(host_rows(d, R, L), L, nthreads=8)
    genes = [(1, L + 1)]
    j = Juliet(L, genes, mode_phasing=True, handle=hd)
    j.pileup_device(d.data_ptr(), R)
    col, codon = j.get_counts()
    ocol, ocodon = oracle.pileup(st, mask_bytes(j.start_mask, L), nthreads=8)
    ocol[:, 6] = 0
    assert np.array_equal(col, ocol) and np.array_equal(codon, ocodon)

    sites = sorted({(c, k) for (_, c, k) in t.truth})
    V = len(sites)
    assert V == 2048
    nw = V // 32
    hap, keys = j.phase_device([_V(c, k) for c, k in sites], d.data_ptr(), R, want_hap_id=True, cap=R)
    assert keys == sites
    obits, oflags = oracle.phase_bits(st, [s[0] for s in sites], [s[1] for s in sites], nthreads=8)
    g = oracle.phase_group(obits, oflags, V)
    gbits, gflags = _device_results(j, hd, R, nw)
    assert np.array_equal(gflags, oflags) and np.array_equal(gbits, obits)
    assert hap.ndistinct == g["H"] > 10_000 and hap.nreported == g["nreported"]
    k = len(hap.counts)
    assert k == g["H"]
    assert np.array_equal(hap.counts, g["counts"]) and np.array_equal(hap.patterns, g["patterns"])
    assert hap.counters == {kk: int(v) for kk, v in g["counters"].items()}
    assert np.array_equal(hap.hap_id, g["hap_id"])

    # co-occurrence: auto picks the tcgen05 kernel here (V >= 256, >= 32768 reads): 8 x 8 tiles -> 36 upper-triangle tiles x split-K
    B = np.unpackbits(obits.view(np.uint8), axis=1, bitorder="little")[:, :V].astype(np.float32)
    want = (B.T @ B).astype(np.int64)
    assert want.max() < 2 ** 24
    for variant in (0, 2, 1):
        _lib.check(j.lib.ms_set_cooccurrence_variant(hd.h, variant), hd.h)
        try:
            got = j.cooccurrence().cpu().numpy().astype(np.int64)
        finally:
            _lib.check(j.lib.ms_set_cooccurrence_variant(hd.h, 0), hd.h)
        assert np.array_equal(got, want), variant


def test_c4_shape_full_genome_call_and_phase(oracle, hd):
    L, R = 9719, 24_000
    t = make_tables(SynthConfig(L=L, seed=20240004))
    d = gpu_synth(hd, t, 0, R)
    st = oracle.unpack(host_rows(d, R, L), L, nthreads=8)
    j = Juliet(L, HIV_GENES, mode_phasing=True, min_perc=0.5, handle=hd)
    res = j.run_device(d.data_ptr(), R, want_hap_id=True)
    col, codon = j.get_counts()
    ocol, ocodon = oracle.pileup(st, mask_bytes(j.start_mask, L), nthreads=8)
    ocol[:, 6] = 0
    assert np.array_equal(col, ocol) and np.array_equal(codon, ocodon)
    ov = oracle.call(codon, HIV_GENES, min_perc=0.5)
    variants_equal(res.variants, ov)
    assert len(ov) >= 6 and len({v.gene for v in ov}) >= 2      # planted variants fall into several genes / frames
    keys = sorted({(v.col, v.codon) for v in ov})
    assert res.keys == keys
    obits, oflags = oracle.phase_bits(st, [k[0] for k in keys], [k[1] for k in keys], nthreads=8)
    g = oracle.phase_group(obits, oflags, len(keys))
    hp = res.haplotypes
    nw = max(1, (len(keys) + 31) // 32)
    gbits, gflags = _device_results(j, hd, R, nw)
    assert np.array_equal(gflags, oflags) and np.array_equal(gbits, obits)
    k = len(hp.counts)
    assert hp.ndistinct == g["H"] and hp.nreported == g["nreported"] >= 3
    assert np.array_equal(hp.counts, g["counts"][:k]) and np.array_equal(hp.patterns, g["patterns"][:k])
    assert hp.counters == {kk: int(v) for kk, v in g["counters"].items()}
    assert np.array_equal(hp.hap_id, g["hap_id"])
    # the same pass from event rows (the e2e entry point) at this length
    from minorseq_b200 import encode_rows
    hdr, ev = encode_rows(host_rows(d, R, L), L, t.refseq)
    j.set_base(t.refseq)
    res2 = j.run_events_host(hdr, ev, want_hap_id=True)
    col2, codon2 = j.get_counts()
    assert np.array_equal(col2, ocol) and np.array_equal(codon2, ocodon)
    variants_equal(res2.variants, ov)
    assert np.array_equal(res2.haplotypes.hap_id, g["hap_id"])
