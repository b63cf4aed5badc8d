"""cleric restatement (oracle/ms_oracle.c, choices U13): Needleman-Wunsch of the two references and the transitive
re-expression of a read's alignment (/root/reference/doc/CLERIC.md:19-23,41-44).  Known answers worked out by hand plus
invariants every projection must keep."""
import numpy as np
import pytest

A = "ACGTACGTTTGACCAGTACGATCGATTACAGGCT"


def cig_query_len(cig):
    return sum(l for l, o in cig if o in "=XIS")


def cig_ref_len(cig):
    return sum(l for l, o in cig if o in "=XD")


def test_nw_identity_and_single_edits(oracle):
    ops, score = oracle.nw_align(A, A)
    assert ops == "M" * len(A) and score == 2 * len(A)
    b = A[:10] + A[11:]                       # B lacks A[10]
    ops, score = oracle.nw_align(A, b)
    assert ops.count("D") == 1 and ops.count("I") == 0 and ops.count("M") == len(A) - 1 and score == 2 * (len(A) - 1) - 4
    b = A[:10] + "G" + A[10:]                 # B has an extra base
    ops, score = oracle.nw_align(A, b)
    assert ops.count("I") == 1 and ops.count("D") == 0 and score == 2 * len(A) - 4
    b = A[:5] + ("A" if A[5] != "A" else "C") + A[6:]
    ops, score = oracle.nw_align(A, b)
    assert ops == "M" * len(A) and score == 2 * (len(A) - 1) - 3
    assert oracle.nw_align("", "ACG") == ("III", -12) and oracle.nw_align("AC", "") == ("DD", -8)


def test_nw_score_matches_bruteforce_dp(oracle):
    rng = np.random.default_rng(5)
    for _ in range(30):
        a = "".join(rng.choice(list("ACGT"), size=int(rng.integers(0, 40))))
        b = "".join(rng.choice(list("ACGT"), size=int(rng.integers(0, 40))))
        ops, score = oracle.nw_align(a, b)
        # the path is consistent with both sequences and reproduces the score
        i = j = s = 0
        for o in ops:
            if o == "M":
                s += 2 if a[i] == b[j] else -3; i += 1; j += 1
            elif o == "D":
                s -= 4; i += 1
            else:
                s -= 4; j += 1
        assert (i, j, s) == (len(a), len(b), score)
        # independent DP for the optimum
        H = np.zeros((len(a) + 1, len(b) + 1), dtype=np.int64)
        H[:, 0] = -4 * np.arange(len(a) + 1); H[0, :] = -4 * np.arange(len(b) + 1)
        for x in range(1, len(a) + 1):
            for y in range(1, len(b) + 1):
                H[x, y] = max(H[x - 1, y - 1] + (2 if a[x - 1] == b[y - 1] else -3), H[x - 1, y] - 4, H[x, y - 1] - 4)
        assert score == H[len(a), len(b)]


def test_projection_known_answers(oracle):
    read = A[4:20]
    # same reference: nothing changes
    ops, _ = oracle.nw_align(A, A)
    assert oracle.project_read(ops, A, 4, [(16, "=")], read) == (4, [(16, "=")])
    # B lacks A[10]: the read's base there becomes an insertion, everything behind it shifts left by one
    b = A[:10] + A[11:]
    ops, _ = oracle.nw_align(A, b)
    gone = ops.index("D")                      # the unpaired A column (NW may shift it inside a homopolymer)
    pos, cig = oracle.project_read(ops, b, 4, [(16, "=")], read)
    assert pos == 4 and cig == [(gone - 4, "="), (1, "I"), (16 - (gone - 4) - 1, "=")]
    # a read that already had a deletion exactly there simply loses it
    pos, cig = oracle.project_read(ops, b, 4, [(gone - 4, "="), (1, "D"), (15 - (gone - 4), "=")], read[:gone - 4] + read[gone - 4 + 1:])
    assert pos == 4 and cig == [(15, "=")]
    # B has an extra base: the read gets a deletion there
    b = A[:10] + "G" + A[10:]
    ops, _ = oracle.nw_align(A, b)
    extra = ops.index("I")
    pos, cig = oracle.project_read(ops, b, 4, [(16, "=")], read)
    assert pos == 4 and cig == [(extra - 4, "="), (1, "D"), (16 - (extra - 4), "=")]
    # a read starting behind the extra base only shifts
    pos, cig = oracle.project_read(ops, b, 12, [(8, "=")], A[12:20])
    assert pos == 13 and cig == [(8, "=")]
    # substitution in B: '=' turns into 'X' and an 'X' that carried B's base turns into '='
    b = A[:6] + ("A" if A[6] != "A" else "C") + A[7:]
    ops, _ = oracle.nw_align(A, b)
    assert oracle.project_read(ops, b, 4, [(16, "=")], read) == (4, [(2, "="), (1, "X"), (13, "=")])
    mut = read[:2] + b[6] + read[3:]
    assert oracle.project_read(ops, b, 4, [(2, "="), (1, "X"), (13, "=")], mut) == (4, [(16, "=")])
    # clips and insertions are carried along; an insertion at the very start becomes a clip
    assert oracle.project_read(ops, b, 4, [(3, "S"), (2, "="), (1, "X"), (4, "="), (2, "I"), (9, "="), (2, "S")],
                               "NNN" + mut[:7] + "GG" + mut[7:] + "NN") == (4, [(3, "S"), (7, "="), (2, "I"), (9, "="), (2, "S")])
    assert oracle.project_read(ops, A, 4, [(2, "I"), (16, "=")], "GG" + read)[1][0] == (2, "S")
    with pytest.raises(ValueError):
        oracle.project_read(ops, b, 4, [(16, "M")], read)          # doc/CLERIC.md:14-15: cigar M is forbidden


def test_projection_invariants_random(oracle):
    """Random references with indels and substitutions, random reads with =/X/I/D/S: the query length never changes,
    every '=' really matches B and every 'X' really differs, and the reference span stays inside B."""
    rng = np.random.default_rng(11)
    for _ in range(60):
        a = "".join(rng.choice(list("ACGT"), size=int(rng.integers(60, 200))))
        b = []
        for ch in a:
            u = rng.random()
            if u < 0.03: continue
            if u < 0.06: b.append(str(rng.choice(list("ACGT"))))
            b.append(ch if rng.random() > 0.05 else str(rng.choice(list("ACGT"))))
        b = "".join(b)
        ops, _ = oracle.nw_align(a, b)
        for _ in range(10):
            pos = int(rng.integers(0, len(a) - 30))
            span = int(rng.integers(10, len(a) - pos))
            cig, seq, i = [], [], pos
            if rng.random() < 0.3:
                k = int(rng.integers(1, 5)); cig.append((k, "S")); seq.append("N" * k)
            while i < pos + span:
                u = rng.random()
                if u < 0.05 and cig and cig[-1][1] in "=X" and i + 1 < pos + span:
                    cig.append((1, "D")); i += 1
                elif u < 0.10 and cig and cig[-1][1] in "=X":
                    k = int(rng.integers(1, 4)); cig.append((k, "I")); seq.append("".join(rng.choice(list("ACGT"), size=k)))
                elif u < 0.15:
                    alt = [c for c in "ACGT" if c != a[i]][int(rng.integers(0, 3))]
                    cig.append((1, "X")); seq.append(alt); i += 1
                else:
                    cig.append((1, "=")); seq.append(a[i]); i += 1
            while cig and cig[-1][1] in "DI":
                if cig[-1][1] == "I": seq.pop()
                cig.pop()
            seq = "".join(seq)
            out = oracle.project_read(ops, b, pos, cig, seq)
            if out is None:
                continue
            npos, ncig = out
            assert cig_query_len(ncig) == len(seq) == cig_query_len(cig)
            assert ncig[0][1] not in "DI" and ncig[-1][1] not in "DI" and all(l > 0 for l, _ in ncig)
            assert all(ncig[k][1] != ncig[k + 1][1] for k in range(len(ncig) - 1))
            assert 0 <= npos and npos + cig_ref_len(ncig) <= len(b)
            q, j = 0, npos
            for l, o in ncig:
                if o == "S" or o == "I": q += l
                elif o == "D": j += l
                else:
                    for t in range(l):
                        assert (seq[q + t] == b[j + t]) == (o == "=")
                    q += l; j += l


def _random_case(rng):
    a = "".join(rng.choice(list("ACGT"), size=int(rng.integers(40, 160))))
    b = []
    for ch in a:
        u = rng.random()
        if u < 0.04: continue
        if u < 0.08: b.append(str(rng.choice(list("ACGT"))))
        b.append(ch if rng.random() > 0.05 else str(rng.choice(list("ACGT"))))
    b = "".join(b)
    pos = int(rng.integers(0, len(a) - 20))
    span = int(rng.integers(5, len(a) - pos + 1))
    cig, seq, i = [], [], pos
    if rng.random() < 0.3:
        k = int(rng.integers(1, 5)); cig.append((k, "H" if rng.random() < 0.3 else "S"))
        if cig[-1][1] == "S": seq.append("N" * k)
    if rng.random() < 0.15:
        k = int(rng.integers(1, 3)); cig.append((k, "I")); seq.append("".join(rng.choice(list("ACGT"), size=k)))
    while i < pos + span:
        u = rng.random()
        if u < 0.07:
            cig.append((1, "D")); i += 1
        elif u < 0.13:
            k = int(rng.integers(1, 4)); cig.append((k, "I")); seq.append("".join(rng.choice(list("ACGT"), size=k)))
        elif u < 0.2:
            cig.append((1, "X")); seq.append([c for c in "ACGT" if c != a[i]][int(rng.integers(0, 3))]); i += 1
        else:
            cig.append((1, "=")); seq.append(a[i]); i += 1
    if rng.random() < 0.3:
        k = int(rng.integers(1, 5)); cig.append((k, "S")); seq.append("N" * k)
    merged = []
    for l, o in cig:
        if merged and merged[-1][1] == o: merged[-1] = (merged[-1][0] + l, o)
        else: merged.append((l, o))
    return a, b, pos, merged, "".join(seq)


def test_host_projection_equals_oracle(oracle, tmp_path):
    """The product's C++ projection (minorseq_b200/host/cleric.hpp) against the restatement on random alignments,
    including reads that begin / end with insertions or deletions and reads made of clips only around a tiny body."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "cleric_project_tool")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-o", exe, os.path.join(root, "tests", "cleric_project_tool.cpp"), "-lz"])
    rng = np.random.default_rng(2024)
    cases, want = [], []
    for _ in range(400):
        a, b, pos, cig, seq = _random_case(rng)
        ops, _ = oracle.nw_align(a, b)
        try:
            out = oracle.project_read(ops, b, pos, cig, seq)
            want.append("unmapped" if out is None else "ok %d %s" % (out[0], "".join("%d%s" % (l, o) for l, o in out[1])))
        except ValueError:
            want.append("inconsistent")
        cases.append("%s %s %d %s %s" % (ops or "-", b or "-", pos, "".join("%d%s" % (l, o) for l, o in cig), seq or "-"))
    cases.append("MMMM ACGT 0 4M ACGT"); want.append("unsupported")
    got = subprocess.run([exe], input="\n".join(cases) + "\n", capture_output=True, text=True, check=True).stdout.strip().split("\n")
    assert len(got) == len(want)
    bad = [(c, g, w) for c, g, w in zip(cases, got, want) if g != w]
    assert not bad, bad[:3]
    assert sum(w.startswith("ok") for w in want) > 300
