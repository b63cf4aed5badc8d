"""Regenerates tests/golden/small_case.npz: a small seeded alignment plus the oracle's answers for it.

The reference ships no fixtures (documentation-only repo), so these vectors pin the RESTATEMENT: they make
an accidental change of the oracle (or of the generator) visible, and give the GPU tests a fixed known
answer that does not depend on building the oracle on the GPU box.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding  # noqa: E402
from minorseq_b200.synth import SynthConfig, make_tables, pack_states, synth_states  # noqa: E402


def main():
    oracle = oracle_binding.load()
    cfg = SynthConfig(L=240, seed=424242, n_rate=5e-3, dele=2e-3, trunc=0.05, ins=2e-3)
    t = make_tables(cfg)
    st = synth_states(t, 0, 1500)
    genes = [(1, 121), (100, 241), (2, 239)]
    mask = np.zeros(240, dtype=np.uint8)
    for (b, e) in genes:
        mask[b - 1: min(e - 1, 240) - 2: 3] = 1
    col, codon = oracle.pileup(st, mask)
    v = oracle.call(codon, genes, refseq=t.refseq)
    var = np.array([(x.gene, x.codon_index, x.col, x.ref_codon, x.codon, x.count, x.coverage, x.expected, x.ntests) for x in v], dtype=np.int64)
    pval = np.array([x.pvalue for x in v], dtype=np.float64)
    keys = sorted({(x.col, x.codon) for x in v})
    bits, flags = oracle.phase_bits(st, [k[0] for k in keys], [k[1] for k in keys])
    g = oracle.phase_group(bits, flags, len(keys))
    ocol, _ = oracle.pileup(st, None, codons=False)
    consensus = oracle.fuse(ocol)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "small_case.npz")
    np.savez_compressed(out, packed=pack_states(st), L=240, genes=np.array(genes), refseq=np.frombuffer(t.refseq.encode(), dtype=np.uint8),
                        col=col, codon=codon, variants=var, pvalues=pval, keys=np.array(keys), hap_counts=g["counts"],
                        hap_patterns=g["patterns"], hap_id=g["hap_id"], nreported=g["nreported"],
                        counters=np.array([g["counters"][k] for k in ("reported", "insufficient", "damaged", "gaps", "heteroduplex", "partial")]),
                        consensus=np.frombuffer(consensus.encode(), dtype=np.uint8))
    print("wrote", out, os.path.getsize(out), "bytes;", len(v), "variants,", g["nreported"], "haplotypes")


if __name__ == "__main__":
    main()
