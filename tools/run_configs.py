"""Times the five BASELINE.json configs on one GPU (device-resident packed reads, CUDA events via the ABI
stopwatch).  Not the bench line -- bench.py is -- but the per-config table in BASELINE.md comes from here."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minorseq_b200 import Fuse, Handle, Juliet, _lib
from minorseq_b200._lib import SynthParams
from minorseq_b200.synth import SynthConfig, make_tables

lib = _lib.load()


def synth(hd, t, R):
    nw = lib.ms_row_words(t.cfg.L)
    d = torch.empty(((R + 7) // 8 * 8, nw), dtype=torch.int32, device="cuda")   # whole tiles (csrc/rows.cuh)
    sp = SynthParams(t.cfg.seed, t.cfg.L, t.nstrains, t.thr_N, t.thr_sub, t.thr_ins20, t.thr_trunc16)
    _lib.check(lib.ms_synth_dev(hd.h, C.byref(sp), t.strain_base.ctypes.data_as(C.c_void_p), t.thr_del.ctypes.data_as(C.c_void_p),
                                t.strain_cum.ctypes.data_as(C.c_void_p), 0, R, C.c_void_p(d.data_ptr())), hd.h)
    return d


def timed(hd, fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ms = C.c_double()
    _lib.check(lib.ms_timer_start(hd.h), hd.h)
    for _ in range(reps):
        out = fn()
    _lib.check(lib.ms_timer_stop(hd.h, C.byref(ms)), hd.h)
    return ms.value / reps, out


def k1_ms(hd):
    ms, rd = C.c_double(), C.c_int64()
    _lib.check(lib.ms_pileup_kernel_ms(hd.h, C.byref(ms), C.byref(rd)), hd.h)
    return ms.value


HIV_GENES = [(1, 634), (790, 1186), (1186, 1879), (1879, 1921), (1921, 2086), (2086, 2134), (2134, 2292), (2253, 2550),
             (2550, 4230), (4230, 5096), (5041, 5620), (5559, 5850), (6062, 6310), (6225, 8795), (8797, 9417)]
rows = []
hd = Handle(0)
_lib.check(lib.ms_set_timing(hd.h, 1), hd.h)
which = sys.argv[1:] or ["C1", "C2", "C3", "C4", "C5"]

if "C1" in which:
    t = make_tables(SynthConfig(L=3000, seed=20240001))
    R = 5000
    d = synth(hd, t, R)
    j = Juliet(3000, [(1, 3001)], refseq=t.refseq, mode_phasing=True, handle=hd)
    ms, res = timed(hd, lambda: j.run_device(d.data_ptr(), R))
    rows.append(dict(config="C1 juliet 5k x 3 kb", reads=R, L=3000, ms=ms, reads_per_s=R / ms * 1e3, k1_ms=k1_ms(hd), variants=len(res.variants),
                     haplotypes=res.haplotypes.nreported))
if "C2" in which:
    t = make_tables(SynthConfig(L=3000, seed=20240002))
    R = 50000
    d = synth(hd, t, R)
    f = Fuse(3000, handle=hd)

    def run():
        f.reset(); f.pileup_device(d.data_ptr(), R); return f.consensus()
    ms, seq = timed(hd, run)
    rows.append(dict(config="C2 fuse 50k x 3 kb", reads=R, L=3000, ms=ms, reads_per_s=R / ms * 1e3, k1_ms=k1_ms(hd), consensus_len=len(seq),
                     consensus_equals_major=(seq == t.refseq)))
if "C3" in which:
    t = make_tables(SynthConfig(L=3000, seed=20240003))
    R = 200000
    d = synth(hd, t, R)
    j = Juliet(3000, [(1, 3001)], refseq=t.refseq, mode_phasing=True, handle=hd)
    ms, res = timed(hd, lambda: j.run_device(d.data_ptr(), R))
    rows.append(dict(config="C3 juliet+phasing 200k x 3 kb, 4 strains", reads=R, L=3000, ms=ms, reads_per_s=R / ms * 1e3, k1_ms=k1_ms(hd),
                     variants=len(res.variants), haplotypes=res.haplotypes.nreported, counters=res.haplotypes.counters))
if "C4" in which:
    t = make_tables(SynthConfig(L=9719, seed=20240004))
    R = 1000000
    d = synth(hd, t, R)
    j = Juliet(9719, HIV_GENES, mode_phasing=True, min_perc=0.5, handle=hd)
    ms, res = timed(hd, lambda: j.run_device(d.data_ptr(), R), reps=3, warm=1)
    k1 = k1_ms(hd)
    rows.append(dict(config="C4 juliet full HIV genome 1M x 9719, 15 genes in 3 frames (one GPU)", reads=R, L=9719, ms=ms, reads_per_s=R / ms * 1e3, k1_ms=k1,
                     k1_GBps=R * 9719 / 2 / k1 / 1e6, variants=len(res.variants), haplotypes=res.haplotypes.nreported))
    del d
if "C5" in which:
    t = make_tables(SynthConfig(L=6144, seed=20240005, dense_sites=2048, dense_strains=64, n_rate=2e-5, dele=2e-5, trunc=0.0))
    R = 500000
    d = synth(hd, t, R)
    sites = sorted({(c, k) for (_, c, k) in t.truth})

    class V:
        def __init__(self, c, k):
            self.col, self.codon = c, k
    j = Juliet(6144, [(1, 6145)], mode_phasing=True, handle=hd)
    j.reset(); j.pileup_device(d.data_ptr(), R)
    pile = k1_ms(hd)
    vs = [V(c, k) for c, k in sites]
    ms_phase, (hap, keys) = timed(hd, lambda: j.phase_device(vs, d.data_ptr(), R, want_hap_id=False), reps=2, warm=1)
    ms_phase_ids, _ = timed(hd, lambda: j.phase_device(vs, d.data_ptr(), R, want_hap_id=True), reps=2, warm=1)
    ms_phase_host, (hap_h, _) = timed(hd, lambda: j.phase_device(vs, d.data_ptr(), R, want_hap_id=False, host_merge=True), reps=1, warm=1)
    assert hap_h.nreported == hap.nreported and hap_h.counters == hap.counters
    assert np.array_equal(hap_h.counts[: hap.nreported], hap.counts[: hap.nreported]) and np.array_equal(hap_h.patterns[: hap.nreported], hap.patterns[: hap.nreported])
    # stage split of the device-ordered pass
    t0 = time.perf_counter(); j.phase_device(vs, d.data_ptr(), R, want_hap_id=False); lib.ms_synchronize(hd.h); t_all = (time.perf_counter() - t0) * 1e3
    ms_co, Cm = timed(hd, lambda: j.cooccurrence(), reps=2, warm=1)
    rows.append(dict(config="C5 phasing stress 500k reads, V=%d sites (dense)" % len(sites), reads=R, L=6144, k1_ms=pile, phase_ms=ms_phase,
                     phase_with_read_ids_ms=ms_phase_ids, phase_host_merge_ms=ms_phase_host, phase_wall_ms=t_all,
                     cooccurrence_ms=ms_co, word_ops=len(sites) * (len(sites) + 1) / 2 * ((R + 31) // 32), distinct_patterns=hap.ndistinct,
                     haplotypes=hap.nreported, counters=hap.counters, diag_sum=int(Cm.diagonal().sum().item())))
for r in rows:
    print(json.dumps(r))
