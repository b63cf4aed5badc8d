"""C++ host layer (minorseq_b200/host): CPU-only self test, and -- on a GPU -- the juliet / fuse binaries
run end to end on BAM files written by an independent Python BAM writer, checked against the oracle."""
import json
import re
import os
import subprocess

import numpy as np
import pytest

import bam_util
from minorseq_b200 import build as msbuild
from minorseq_b200.synth import SynthConfig, make_tables, synth_states

BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "minorseq_b200", "bin")


@pytest.fixture(scope="module")
def binaries():
    msbuild.build_library()
    msbuild.build_host()
    return BIN


def test_host_selftest(binaries, tmp_path):
    out = subprocess.run([os.path.join(binaries, "host_selftest"), str(tmp_path / "t.bam")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "host selftest ok" in out.stdout


def test_cli_fails_loudly_without_gpu(binaries, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = subprocess.run([os.path.join(binaries, "juliet"), "x.bam", str(tmp_path / "o.json")], capture_output=True, text=True)
    assert out.returncode != 0 and "no CPU fallback" in out.stderr
    assert subprocess.run([os.path.join(binaries, "juliet"), "--help"], capture_output=True).returncode == 0


def test_mixdata_follows_the_mixing_rule(binaries, tmp_path):
    """doc/MIXDATA.md:9-22: first file is the major, each minor contributes PERCENTAGE % of COVERAGE reads."""
    paths = []
    for k, base in enumerate("ACG"):
        recs = [bam_util.record(f"{base}/{i}", 0, i % 7, [(30, "=")], base * 30) for i in range(400)]
        p = str(tmp_path / f"clone{k}.bam")
        bam_util.write_bam(p, "ref", 100, recs)
        paths.append(p)
    env = dict(os.environ, COVERAGE="300", PERCENTAGE="10", OUTPUT_PREFIX=str(tmp_path / "mixed"))
    res = subprocess.run([os.path.join(binaries, "mixdata")] + paths, capture_output=True, text=True, env=env)
    assert res.returncode == 0, res.stderr
    # read the result back with the same independent logic: count read-name prefixes via a tiny parse
    import gzip
    raw = gzip.open(str(tmp_path / "mixed.bam")).read()      # BGZF is a valid multi-member gzip stream
    assert raw[:4] == b"BAM\x01"
    counts = {b: raw.count(f"{b}/".encode()) for b in "ACG"}
    assert counts == {"A": 240, "C": 30, "G": 30}


def _make_bam(path, t, st, insertions=None, n_via_qv=False):
    recs = []
    ref = t.strain_base[0]
    for r in range(st.shape[0]):
        flag = 0x10 if r % 5 == 0 else 0
        rec = bam_util.states_to_record(f"m/{r}/ccs", st[r], ref, insertions.get(r) if insertions else None, flag, n_via_qv)
        if rec is not None:
            recs.append(rec)
    # records juliet must ignore: secondary and unmapped (doc/JULIET.md:58)
    recs.append(bam_util.record("secondary", 0x100, 0, [(30, "=")], "A" * 30))
    recs.append(bam_util.record("unmapped", 0x4, 0, [], "ACGT", ref_id=-1))
    bam_util.write_bam(path, "synthetic_ref", t.cfg.L, recs)


@pytest.mark.gpu
@pytest.mark.parametrize("n_via_qv", [False, True], ids=["N-in-seq", "N-by-richQV"])
def test_juliet_cli_matches_oracle(binaries, oracle, tmp_path, n_via_qv):
    cfg = SynthConfig(L=600, seed=77, n_rate=2e-3, trunc=0.05, ins=0.0)
    t = make_tables(cfg)
    st = synth_states(t, 0, 6000)
    bam = str(tmp_path / "in.bam")
    _make_bam(bam, t, st, n_via_qv=n_via_qv)
    conf = {"genes": [{"name": "geneA", "begin": 1, "end": 301, "drms": [{"name": "drugX", "positions": [str(p) for p in range(1, 101)]}]},
                      {"name": "geneB", "begin": 302, "end": 599}],
            "referenceName": "synthetic_ref", "referenceSequence": t.refseq, "version": "test"}
    cpath = tmp_path / "cfg.json"
    cpath.write_text(json.dumps(conf))
    oj, oh = str(tmp_path / "out.json"), str(tmp_path / "out.html")
    tj = str(tmp_path / "timing.json")
    res = subprocess.run([os.path.join(binaries, "juliet"), "--config", str(cpath), "--mode-phasing", "--timing-json", tj, bam, oj, oh],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    rep = json.load(open(oj))
    genes = [(1, 301), (302, 599)]
    keep = (st & 7 != 7).any(axis=1)
    sto = st[keep]
    mask = np.zeros(600, dtype=np.uint8)
    for (b, e) in genes:
        mask[b - 1: e - 3: 3] = 1
    if res.returncode == 0:
        tim = json.load(open(tj))
        assert tim["reads"] > 0 and len(tim["stages"]) >= 4 and tim["total_ms"] >= sum(x["ms"] for x in tim["stages"]) * 0.99
    col, codon = oracle.pileup(sto, mask)
    ov = oracle.call(codon, genes, refseq=t.refseq)
    got = []
    for gi, g in enumerate(rep["genes"]):
        for vp in g["variant_positions"]:
            for va in vp["variant_amino_acids"]:
                for vc in va["variant_codons"]:
                    got.append((gi, vp["ref_position"], vc["codon"], vc["count"], vp["coverage"], vc["pValue"]))
    want = [(v.gene, v.codon_index + 1, "".join("ACGT"[(v.codon >> s) & 3] for s in (4, 2, 0)), v.count, v.coverage, v.pvalue) for v in ov]
    assert sorted(x[:5] for x in got) == sorted(x[:5] for x in want)
    for a, b in zip(sorted(got), sorted(want)):
        assert abs(a[5] - b[5]) <= 1e-9 * b[5]
    # MSA context rows come from the column counts
    vp = rep["genes"][0]["variant_positions"][0]
    c0 = 3 * (vp["ref_position"] - 1)
    for row in vp["msa"]:
        j = c0 + row["rel_pos"]
        assert [row[k] for k in "ACGT-N"] == [int(x) for x in col[j, :6]]
    # haplotypes: counts, order, categories and read names against the oracle
    keys = sorted({(v.col, v.codon) for v in ov})
    bits, flags = oracle.phase_bits(sto, [k[0] for k in keys], [k[1] for k in keys])
    g = oracle.phase_group(bits, flags, len(keys))
    assert [h["reads"] for h in rep["haplotypes"]] == [int(x) for x in g["counts"][: g["nreported"]]]
    assert [h["name"] for h in rep["haplotypes"]] == [oracle.hap_name(i) for i in range(g["nreported"])]
    cats = rep["haplotype_read_categories"]
    assert (cats["reported"], cats["insufficient_coverage"], cats["unsuitable"]) == (g["counters"]["reported"], g["counters"]["insufficient"], g["counters"]["damaged"])
    names = np.array([f"m/{r}/ccs" for r in range(st.shape[0])])[keep]
    for k, h in enumerate(rep["haplotypes"]):
        assert h["read_names"] == list(names[g["hap_id"] == k])
    # DRM annotation: every geneA variant at AA positions 1..100 is labelled drugX
    for vp in rep["genes"][0]["variant_positions"]:
        for va in vp["variant_amino_acids"]:
            for vc in va["variant_codons"]:
                assert (vc["known_drm"] == "drugX") == (vp["ref_position"] <= 100)
    html = open(oh).read()
    # the page is a 1:1 rendering of the JSON (doc/JULIET.md:68-107): four sections in the documented order, the input block's four
    # fields, the per-gene table header of the screenshots, one haplotype column per reported haplotype with its percentage and
    # read count, the read-category tooltip (:372-381), one row per variant codon and the -3..+5 context rows (juliet_hiv-context.png)
    order = [html.index(f"<summary>{t}</summary>") for t in ("Input data", "Target config", "Variant Discovery", "Drug Summaries")]
    assert order == sorted(order)
    for field in ("Timestamp:", "Input File:", "Command Line Call:", "Juliet Version:"):
        assert field in html
    assert rep["input"]["command_line"] in html and "--mode-phasing" in rep["input"]["command_line"]
    assert "<th>Codon</th><th>AA</th><th>Pos</th><th>AA</th><th>Codon</th><th>%</th><th>Coverage</th><th>Affected Drugs</th>" in html
    assert "Sample Variants" in html and "Haplotypes %" in html and "geneA" in html and "geneB" in html and "synthetic_ref" in html
    for hp in rep["haplotypes"]:
        assert f'<th title="{hp["reads"]} reads">{hp["percentage"]}</th>' in html
        assert re.search(rf'<th class="gene" style="color:[^"]+">{hp["name"]}</th>', html)
    tip = (f'Reported: {cats["reported"]} | Insufficient coverage: {cats["insufficient_coverage"]} | Unsuitable: {cats["unsuitable"]} '
           f'(gaps {cats["unsuitable_gaps"]}, heteroduplexes {cats["unsuitable_heteroduplex"]}, partial {cats["unsuitable_partial"]})')
    assert tip in html
    assert cats["unsuitable_gaps"] == g["counters"]["gaps"] and cats["unsuitable_heteroduplex"] == g["counters"]["heteroduplex"] and cats["unsuitable_partial"] == g["counters"]["partial"]
    assert html.count('<tr class="var"') == len(got)
    nctx = sum(len(vp["msa"]) for gg in rep["genes"] for vp in gg["variant_positions"])
    assert html.count('<tr class="msa m') == nctx + sum(len(gg["variant_positions"]) for gg in rep["genes"])      # + one header row per position
    assert "drugX" in html and html.count("<b>drugX</b>") == 1                                                 # drug summary entry
    nhit = sum(sum(vc["haplotype_hit"]) for gg in rep["genes"] for vp in gg["variant_positions"] for va in vp["variant_amino_acids"] for vc in va["variant_codons"])
    assert html.count('<td class="hap" style="background:') == nhit
    # --drm-only, --region, --min-perc
    res = subprocess.run([os.path.join(binaries, "juliet"), "-c", str(cpath), "--drm-only", "--region", "1-200", "--min-perc", "2", bam, oj], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    rep2 = json.load(open(oj))
    ov2 = oracle.call(codon, genes, refseq=t.refseq, region=(1, 200), min_perc=2.0)
    got2 = sorted((gi, vp["ref_position"], vc["codon"]) for gi, g2 in enumerate(rep2["genes"]) for vp in g2["variant_positions"]
                  for va in vp["variant_amino_acids"] for vc in va["variant_codons"])
    want2 = sorted((v.gene, v.codon_index + 1, "".join("ACGT"[(v.codon >> s) & 3] for s in (4, 2, 0))) for v in ov2 if v.gene == 0 and v.codon_index < 100)
    assert got2 == want2


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
def test_juliet_and_fuse_cli_two_gpus_equal_one(binaries, tmp_path):
    """`juliet --gpus 2` (reads sharded over two GPUs, one NCCL all-reduce, device-side haplotype merge) writes the same JSON and
    HTML, byte for byte, as `--gpus 1`; likewise fuse's FASTA.  MS_FIXED_TIMESTAMP pins the only run-dependent field."""
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs")
    cfg = SynthConfig(L=900, seed=78, n_rate=2e-3, trunc=0.05, ins=2e-3)
    t = make_tables(cfg)
    st = synth_states(t, 0, 20001)
    bam = str(tmp_path / "in.bam")
    _make_bam(bam, t, st)
    env = dict(os.environ, MS_FIXED_TIMESTAMP="2026-01-01T00:00:00.000Z")
    outs = {}
    for n in (1, 2):
        oj, oh, of = str(tmp_path / f"o{n}.json"), str(tmp_path / f"o{n}.html"), str(tmp_path / f"o{n}.fasta")
        res = subprocess.run([os.path.join(binaries, "juliet"), "--mode-phasing", "--gpus", str(n), bam, oj, oh], capture_output=True, text=True, env=env)
        assert res.returncode == 0, res.stderr
        res = subprocess.run([os.path.join(binaries, "fuse"), "--gpus", str(n), bam, of], capture_output=True, text=True, env=env)
        assert res.returncode == 0, res.stderr
        rep = json.load(open(oj))
        cmd = rep["input"]["command_line"]
        outs[n] = (open(oj).read().replace(cmd, "CMD"), open(oh).read().replace(cmd, "CMD"), open(of).read())
        assert len(rep["haplotypes"]) >= 2 and sum(len(g["variant_positions"]) for g in rep["genes"]) >= 3
    for a, b in zip(outs[1], outs[2]):
        a2 = a.replace("o1.json", "X").replace("o1.html", "X")
        b2 = b.replace("o2.json", "X").replace("o2.html", "X")
        assert a2 == b2


@pytest.mark.gpu
def test_juliet_cli_rejects_cigar_m(binaries, tmp_path):
    bam = str(tmp_path / "m.bam")
    bam_util.write_bam(bam, "r", 100, [bam_util.record("bad", 0, 0, [(50, "M")], "A" * 50)])
    res = subprocess.run([os.path.join(binaries, "juliet"), bam, str(tmp_path / "o.json")], capture_output=True, text=True)
    assert res.returncode != 0 and "cigar M is forbidden" in res.stderr


@pytest.mark.gpu
def test_fuse_cli_matches_oracle(binaries, oracle, tmp_path):
    cfg = SynthConfig(L=900, seed=5, ins=0.0, trunc=0.02)
    t = make_tables(cfg)
    st = synth_states(t, 0, 2000)
    st[:, 300:303] = np.where(np.arange(2000)[:, None] % 4 != 0, np.uint8(4), st[:, 300:303])     # major deletion
    ins = {}
    for r in range(2000):
        d = {}
        if r % 10 < 8 and (st[r, 450] & 7) != 7:
            d[450] = "GGC"
        if r % 10 < 7 and (st[r, 460] & 7) != 7:
            d[460] = "TTT"           # too close to the accepted one
        if r % 10 < 9 and (st[r, 600] & 7) != 7:
            d[600] = "AC"            # not in frame
        for c in d:
            st[r, c] |= 8
        ins[r] = d
    bam = str(tmp_path / "f.bam")
    _make_bam(bam, t, st, insertions=ins)
    fa = str(tmp_path / "out.fasta")
    res = subprocess.run([os.path.join(binaries, "fuse"), bam, fa], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    lines = open(fa).read().split("\n")
    assert lines[0].startswith(">f.bam|fuse|synthetic_ref")
    got = "".join(lines[1:])
    keep = (st & 7 != 7).any(axis=1)
    ocol, _ = oracle.pileup(st[keep], None, codons=False)
    ev = [(c, s) for r in range(2000) if keep[r] for c, s in sorted(ins[r].items())]
    pool = "".join(s for _, s in ev).encode()
    il = [len(s) for _, s in ev]
    io = np.concatenate([[0], np.cumsum(il)[:-1]]) if ev else []
    want = oracle.fuse(ocol, [c for c, _ in ev], io, il, pool)
    assert got == want
    assert "GGC" in got and len(got) == 900 - 3 + 3


@pytest.mark.gpu
@pytest.mark.parametrize("combined", [False, True])
def test_cleric_cli_matches_oracle(binaries, oracle, tmp_path, combined):
    """`cleric in.bam reference.fasta new_ref.fasta out.bam` and the combined-FASTA form (doc/CLERIC.md:27-38): GPU
    Needleman-Wunsch of the two references + host projection, every output record against the restatement."""
    from test_oracle_cleric import _random_case
    rng = np.random.default_rng(77)
    a = "".join(rng.choice(list("ACGT"), size=1500))
    b = []
    for ch in a:
        u = rng.random()
        if u < 0.01: continue
        if u < 0.02: b.append(str(rng.choice(list("ACGT"))))
        b.append(ch if rng.random() > 0.03 else str(rng.choice(list("ACGT"))))
    b = "".join(b)
    recs, meta = [], []
    for r in range(300):
        pos = int(rng.integers(0, len(a) - 400))
        span = int(rng.integers(50, 400))
        cig, seq, i = [], [], pos
        if r % 5 == 0:
            cig.append((3, "S")); seq.append("NNN")
        while i < pos + span:
            u = rng.random()
            if u < 0.02 and cig and cig[-1][1] in "=X" and i + 1 < pos + span: cig.append((1, "D")); i += 1
            elif u < 0.04 and cig and cig[-1][1] in "=X": cig.append((2, "I")); seq.append("GT")
            elif u < 0.06: cig.append((1, "X")); seq.append([c for c in "ACGT" if c != a[i]][0]); i += 1
            else: cig.append((1, "=")); seq.append(a[i]); i += 1
        while cig[-1][1] in "DI":
            if cig[-1][1] == "I": seq.pop()
            cig.pop()
        merged = []
        for l, o in cig:
            if merged and merged[-1][1] == o: merged[-1] = (merged[-1][0] + l, o)
            else: merged.append((l, o))
        seq = "".join(seq)
        aux = bam_util.aux_z("MD", "10A5") + b"NMC\x03" + bam_util.aux_z("sq", "I" * len(seq)) if r % 3 == 0 else bam_util.aux_z("sq", "I" * len(seq))
        recs.append(bam_util.record(f"read/{r}/ccs", 0 if r % 7 else 2048, pos, merged, seq, aux=aux))
        meta.append((pos, merged, seq))
    recs.append(bam_util.record("unmapped/1/ccs", 4, -1, [], "ACGT", ref_id=-1))
    bam_util.write_bam(str(tmp_path / "in.bam"), "orig", len(a), recs, sort_order="coordinate" if combined else "unknown")
    (tmp_path / "a.fa").write_text(">orig some description\n" + "\n".join(a[i:i + 60] for i in range(0, len(a), 60)) + "\n")
    (tmp_path / "b.fa").write_text(">target\n" + b.lower() + "\n")
    if combined:
        (tmp_path / "c.fa").write_text((tmp_path / "b.fa").read_text() + (tmp_path / "a.fa").read_text())   # order must not matter
        cmd = [os.path.join(binaries, "cleric"), str(tmp_path / "in.bam"), str(tmp_path / "c.fa"), str(tmp_path / "out.bam")]
    else:
        cmd = [os.path.join(binaries, "cleric"), str(tmp_path / "in.bam"), str(tmp_path / "a.fa"), str(tmp_path / "b.fa"), str(tmp_path / "out.bam")]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    text, refs, got = bam_util.read_bam(str(tmp_path / "out.bam"))
    assert refs == [("target", len(b))] and "SN:target" in text and "SN:orig" not in text and "ID:cleric" in text
    assert len(got) == len(recs)
    ops, score = oracle.nw_align(a, b)
    assert f"alignment score {score})" in out.stderr
    for g, (pos, cig, seq) in zip(got[:-1], meta):
        want = oracle.project_read(ops, b, pos, cig, seq)
        assert want is not None
        assert (g["pos"], g["cigar"], g["seq"], g["ref_id"]) == (want[0], want[1], seq, 0), g["name"]
        # tags that describe the alignment to the old reference are gone, the QV track stays; the bin is the new interval's
        assert g["aux"] == bam_util.aux_z("sq", "I" * len(seq)), g["name"]
        reflen = sum(l for l, o in want[1] if o in "=XDNM")
        assert g["bin"] == bam_util.reg2bin(want[0], want[0] + max(1, reflen)), g["name"]
    assert got[-1]["flag"] & 4 and got[-1]["ref_id"] == -1 and got[0]["flag"] == 2048 and got[-1]["bin"] == 4680
    # no record lost its place here, so a sort order in the header survives
    assert ("SO:coordinate" in text) == bool(combined)
    # the BAM's reference must be one of the two sequences (doc/CLERIC.md:16-17); cigar M is refused (:14-15)
    (tmp_path / "x.fa").write_text(">other\nACGT\n")
    bad = subprocess.run([os.path.join(binaries, "cleric"), str(tmp_path / "in.bam"), str(tmp_path / "x.fa"), str(tmp_path / "b.fa"), str(tmp_path / "o2.bam")],
                         capture_output=True, text=True)
    assert bad.returncode != 0 and "must match the reference name" in bad.stderr
    bam_util.write_bam(str(tmp_path / "m.bam"), "orig", len(a), [bam_util.record("m/1/ccs", 0, 0, [(4, "M")], "ACGT")])
    bad = subprocess.run([os.path.join(binaries, "cleric"), str(tmp_path / "m.bam"), str(tmp_path / "a.fa"), str(tmp_path / "b.fa"), str(tmp_path / "o3.bam")],
                         capture_output=True, text=True)
    assert bad.returncode != 0 and "cigar M is forbidden" in bad.stderr
