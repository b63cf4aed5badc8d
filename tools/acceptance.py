"""The reference's own acceptance statements for juliet's statistical model, run through the GPU path.

/root/reference/doc/JULIET.md:34-36   "At a coverage of 6000 CCS reads with a predicted accuracy (RQ) of >=0.99, the false positive
                                      and false negative rates are below 1% and 0.001%"
/root/reference/doc/JULIET.md:233-237 minimal / reliable coverage per minor frequency: 1 % 2500x / 6000x, 5 % 500x / 1200x, 10 % 250x / 600x
                                      ("For the minimal coverage, FP/FN rates may increase")
/root/reference/doc/JULIET.md:249-251 "We tested clean samples, amplified in plasmids, and at 25000x there is not a single false positive call."

SURVEY.md section 6 / App. B U1 makes these the condition for accepting the restatement's unpinned constants (error-model rates,
Fisher table form, Bonferroni factor, alpha): a restatement that calls variants in clean data at 6000x, or misses 1 % minors there,
has the wrong defaults.  Every trial is an independent synthetic 3 kb amplicon (own reference, own variant positions, own noise),
piled up, tested and filtered by the CUDA kernels behind the C ABI with juliet's default options.

    python tools/acceptance.py [--trials 40] [--json out.json]        (needs a GPU)

Empirical rates from T trials cannot resolve 1e-5, so the false-negative column also gives the model's exact figure: the smallest
observed count k* the test calls at that coverage (found with the library's own test on synthetic histograms) and the binomial
probability that a minor of the stated frequency shows fewer than k* reads.
"""
import argparse
import ctypes as C
import json
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

L = 3000
GENES = [(1, L + 1)]


def _synth(hd, t, R):
    import torch
    from minorseq_b200 import _lib
    from minorseq_b200._lib import SynthParams
    lib = _lib.load()
    d = torch.empty(((R + 7) // 8 * 8, lib.ms_row_words(L)), dtype=torch.int32, device="cuda")   # whole tiles (csrc/rows.cuh)
    sp = SynthParams(t.cfg.seed, L, t.nstrains, t.thr_N, t.thr_sub, t.thr_ins20, t.thr_trunc16)
    _lib.check(lib.ms_synth_dev(hd.h, C.byref(sp), t.strain_base.ctypes.data_as(C.c_void_p), t.thr_del.ctypes.data_as(C.c_void_p),
                                t.strain_cum.ctypes.data_as(C.c_void_p), 0, R, C.c_void_p(d.data_ptr())), hd.h)
    return d


def clean_sample(hd, coverage, trials, seed0=910000):
    """One strain, sequencing noise only: every call is a false positive."""
    from minorseq_b200 import Juliet
    from minorseq_b200.synth import SynthConfig, make_tables
    calls, positions, samples_with_call = 0, 0, 0
    for k in range(trials):
        t = make_tables(SynthConfig(L=L, seed=seed0 + k, minor_fracs=(), variants_per_minor=(0, 0)))
        d = _synth(hd, t, coverage)
        j = Juliet(L, GENES, refseq=t.refseq, handle=hd)
        j.pileup_device(d.data_ptr(), coverage)
        v = j.call()
        calls += len(v)
        positions += L // 3
        samples_with_call += 1 if v else 0
    return dict(coverage=coverage, trials=trials, false_calls=calls, tested_positions=positions,
                fp_rate_per_position=calls / positions, samples_with_a_false_call=samples_with_call)


def minor_sample(hd, frac, coverage, trials, seed0=920000):
    """Major + one minor strain at `frac` carrying 3-5 private codon substitutions: planted variants found / missed, plus
    every call that is not a planted variant."""
    from minorseq_b200 import Juliet
    from minorseq_b200.synth import SynthConfig, make_tables
    planted = found = false_calls = 0
    for k in range(trials):
        t = make_tables(SynthConfig(L=L, seed=seed0 + 1000 * int(frac * 1000) + k, minor_fracs=(frac,), variants_per_minor=(3, 5)))
        d = _synth(hd, t, coverage)
        j = Juliet(L, GENES, refseq=t.refseq, handle=hd)
        j.pileup_device(d.data_ptr(), coverage)
        got = {(v.col, v.codon) for v in j.call()}
        truth = {(c, cod) for (_, c, cod) in t.truth}
        planted += len(truth)
        found += len(truth & got)
        false_calls += len(got - truth)
    return dict(minor_frequency=frac, coverage=coverage, trials=trials, planted=planted, found=found, missed=planted - found,
                detection_rate=found / planted, false_calls=false_calls)


def call_threshold(hd, n, ref_codon=0, codon=1):
    """Smallest observed count of one single-substitution codon that the library calls at codon coverage n with juliet's defaults
    (one 1000-codon gene): the test is run on a synthetic histogram written into the count tensor."""
    import torch
    from minorseq_b200 import Juliet
    refseq = "A" * L     # every reference codon is AAA (index 0); `codon` = AAC
    j = Juliet(L, GENES, refseq=refseq, handle=hd)
    lo, hi = 1, n
    while lo < hi:
        k = (lo + hi) // 2
        hist = np.zeros((L, 64), dtype=np.uint32)
        hist[::3, ref_codon] = n
        hist[0, ref_codon] = n - k
        hist[0, codon] = k
        ct = j.counts_tensor()
        ct[L * 8:] = torch.from_numpy(hist.reshape(-1).view(np.int32)).cuda()
        torch.cuda.synchronize()
        if any(v.col == 0 and v.codon == codon for v in j.call()):
            hi = k
        else:
            lo = k + 1
    return lo


def binom_cdf_below(kstar, n, f):
    """P(X < kstar), X ~ Binomial(n, f), summed in log space."""
    if kstar <= 0:
        return 0.0
    lf, l1 = math.log(f), math.log1p(-f)
    tot = 0.0
    for k in range(kstar):
        tot += math.exp(math.lgamma(n + 1) - math.lgamma(k + 1) - math.lgamma(n - k + 1) + k * lf + (n - k) * l1)
    return tot


def run(trials=40, hd=None):
    from minorseq_b200 import Handle
    own = hd is None
    hd = hd or Handle(0)
    try:
        out = dict(model="restatement defaults: substitution 5e-4, deletion 3e-3, alpha 0.01, Bonferroni factor = codons of the gene (1000)",
                   noise="synthetic CCS-like reads: substitution 5e-4, deletion 3e-3 (x10 in homopolymers), insertion 1e-3, QV-filtered N 2e-2, 2 % truncated",
                   clean=[], minors=[])
        for cov in (6000, 25000):
            out["clean"].append(clean_sample(hd, cov, trials))
        clean_codon = (1 - 2e-2 - 3e-3 - 5e-4) ** 3     # share of reads whose codon is all A/C/G/T and undamaged (about 0.93)
        for frac, cov, kind in ((0.01, 2500, "minimal"), (0.01, 6000, "reliable"), (0.05, 500, "minimal"), (0.05, 1200, "reliable"),
                                (0.10, 250, "minimal"), (0.10, 600, "reliable")):
            r = minor_sample(hd, frac, cov, trials)
            n = int(round(cov * clean_codon))
            ks = call_threshold(hd, n)
            r.update(kind=kind, codon_coverage=n, smallest_called_count=ks, expected_count=frac * n,
                     model_false_negative_rate=binom_cdf_below(ks, n, frac))
            out["minors"].append(r)
        return out
    finally:
        if own:
            hd.close()


def table(res):
    lines = ["| Sample | Coverage | Trials | Result | Doc statement |", "|---|---|---|---|---|"]
    for c in res["clean"]:
        doc = "FP < 1 % (`doc/JULIET.md:34-36`)" if c["coverage"] == 6000 else "not a single false positive call (`:249-251`)"
        lines.append(f"| clean, one strain | {c['coverage']}x | {c['trials']} | {c['false_calls']} false calls in {c['tested_positions']} tested codon positions "
                     f"(FP rate {100 * c['fp_rate_per_position']:.3g} %) | {doc} |")
    for m in res["minors"]:
        doc = "FN < 0.001 % (`:34-36`)" if (m["minor_frequency"], m["coverage"]) == (0.01, 6000) else f"{m['kind']} coverage (`:233-237`)"
        lines.append(f"| {100 * m['minor_frequency']:g} % minor | {m['coverage']}x | {m['trials']} | {m['found']}/{m['planted']} planted variants called, "
                     f"{m['false_calls']} other calls; smallest called count {m['smallest_called_count']} of {m['codon_coverage']} "
                     f"(expected {m['expected_count']:.1f}) -> model FN {m['model_false_negative_rate']:.2g} | {doc} |")
    return "\n".join(lines)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=40)
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    res = run(a.trials)
    print(table(res))
    if a.json:
        json.dump(res, open(a.json, "w"), indent=1)
