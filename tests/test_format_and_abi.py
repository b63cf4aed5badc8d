"""Host logic: packed format round trips, CIGAR expansion, and that the C-ABI library loads and
exports every symbol include/minorseq_b200.h declares (no GPU compute here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from minorseq_b200 import _lib
from minorseq_b200.synth import SynthConfig, make_tables, pack_states, synth_states, unpack_states, start_mask_words

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(mslib):
    hdr = open(os.path.join(ROOT, "include", "minorseq_b200.h")).read()
    names = set(re.findall(r"\b(ms_[a-z0-9_]+)\s*\(", hdr))
    assert names, "no declarations parsed"
    for n in sorted(names):
        assert hasattr(mslib, n), f"{n} declared in the header but not exported"
    assert names == set(_lib.EXPORTED_SYMBOLS)


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C99 and as C++ on its own (no torch / CUDA types)."""
    import shutil, subprocess
    hdr = os.path.join(ROOT, "include", "minorseq_b200.h")
    for cc, lang, std in (("gcc", "c", "c99"), ("g++", "c++", "c++11")):
        exe = shutil.which(cc, path="/usr/bin") or shutil.which(cc)
        if not exe:
            pytest.skip(f"{cc} not found")
        out = subprocess.run([exe, "-x", lang, f"-std={std}", "-fsyntax-only", "-Wall", "-Werror", hdr], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
    text = open(hdr).read()
    assert "#include <torch" not in text and "cuda_runtime" not in text and "ATen" not in text


def test_create_fails_loudly_without_gpu(mslib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = mslib.ms_create(0, C.byref(h))
    assert rc == -3 and not h.value
    assert b"no CPU fallback" in mslib.ms_last_error(None)


@pytest.mark.parametrize("L", [3, 31, 32, 33, 96, 3000, 9719])
def test_pack_roundtrip(mslib, oracle, L):
    rng = np.random.default_rng(L)
    R = 17
    st = rng.choice(np.array([0, 1, 2, 3, 4, 5, 7], dtype=np.uint8), size=(R, L)) | (rng.integers(0, 2, size=(R, L), dtype=np.uint8) << 3)
    st = np.where((st & 7) == 7, np.uint8(7), st).astype(np.uint8)
    nw = mslib.ms_row_words(L)
    assert nw == 4 * ((L + 31) // 32)
    packed = np.zeros((R, nw), dtype=np.uint32)
    assert mslib.ms_pack_states(st.ctypes.data_as(C.c_void_p), R, L, packed.ctypes.data_as(C.c_void_p)) == 0
    assert np.array_equal(packed, pack_states(st))                       # C++ packer == numpy packer
    assert np.array_equal(oracle.unpack(packed, L), st)                  # oracle's own unpacker
    back = np.zeros_like(st)
    assert mslib.ms_unpack_states(packed.ctypes.data_as(C.c_void_p), R, L, back.ctypes.data_as(C.c_void_p)) == 0
    assert np.array_equal(back, st)
    assert np.array_equal(unpack_states(packed, L), st)
    # columns past L in the last block are "not spanned"
    if L % 32:
        last = packed.reshape(R, -1, 4)[:, -1, :]
        hi = np.uint32((0xFFFFFFFF << (L % 32)) & 0xFFFFFFFF)
        assert ((last[:, 0] & hi) == hi).all() and ((last[:, 2] & hi) == hi).all() and ((last[:, 3] & hi) == 0).all()


def test_read_admission_filter(mslib):
    """doc/JULIET.md:58: only primary and supplementary alignments are used."""
    assert mslib.ms_read_admitted(0) == 1 and mslib.ms_read_admitted(16) == 1      # primary fwd / rev
    assert mslib.ms_read_admitted(0x800) == 1 and mslib.ms_read_admitted(0x810) == 1  # supplementary
    assert mslib.ms_read_admitted(0x100) == 0 and mslib.ms_read_admitted(0x4) == 0    # secondary, unmapped
    assert mslib.ms_read_admitted(0x904) == 0


def test_pack_rejects_reserved_state(mslib):
    st = np.array([[6, 0, 1]], dtype=np.uint8)
    packed = np.zeros((1, 4), dtype=np.uint32)
    assert mslib.ms_pack_states(st.ctypes.data_as(C.c_void_p), 1, 3, packed.ctypes.data_as(C.c_void_p)) == -5


def _expand(mslib, cigar, pos, seq, L, qv=None, want_ins=True):
    ops = "MIDNSHP=X"
    enc = np.array([(n << 4) | ops.index(o) for n, o in cigar], dtype=np.uint32)
    row = np.zeros(mslib.ms_row_words(L), dtype=np.uint32)
    ic, io, il = np.zeros(16, np.int32), np.zeros(16, np.int64), np.zeros(16, np.int32)
    pool = np.zeros(256, np.uint8)
    nins, used = C.c_int64(0), C.c_int64(0)
    q = None if qv is None else np.asarray(qv, dtype=np.uint8)
    rc = mslib.ms_expand_cigar(enc.ctypes.data_as(C.c_void_p), len(enc), pos, seq.encode(),
                               None if q is None else q.ctypes.data_as(C.c_void_p), len(seq), L,
                               row.ctypes.data_as(C.c_void_p), ic.ctypes.data_as(C.c_void_p),
                               io.ctypes.data_as(C.c_void_p), il.ctypes.data_as(C.c_void_p), 16, C.byref(nins),
                               pool.ctypes.data_as(C.c_void_p), 256, C.byref(used))
    ins = [(int(ic[i]), pool[io[i]: io[i] + il[i]].tobytes().decode()) for i in range(nins.value)]
    return rc, (unpack_states(row[None, :], L)[0] if rc == 0 else None), ins


def test_expand_cigar(mslib):
    L = 40
    #        pos=2: 3=, 2I, 1X, 2D, 2=, 1S
    rc, st, ins = _expand(mslib, [(3, "="), (2, "I"), (1, "X"), (2, "D"), (2, "="), (1, "S")], 2, "ACGTTACGA", L)
    assert rc == 0
    want = np.full(L, 7, dtype=np.uint8)
    want[2:5] = [0, 1, 2]; want[4] |= 8          # insertion follows column 4
    want[5] = 0                                   # X: 'A'
    want[6:8] = 4
    want[8:10] = [1, 2]
    assert np.array_equal(st, want)
    assert ins == [(4, "TT")]
    # QV-filtered base becomes N (doc/JULIET.md:256-259)
    rc, st, _ = _expand(mslib, [(4, "=")], 0, "ACGT", L, qv=[0, 1, 0, 0])
    assert rc == 0 and list(st[:4]) == [0, 5, 2, 3]
    # CIGAR M is forbidden (doc/JULIET.md:53)
    rc, _, _ = _expand(mslib, [(4, "M")], 0, "ACGT", L)
    assert rc == -5
    # clipping at the reference end and negative overhang is tolerated
    rc, st, _ = _expand(mslib, [(6, "=")], 37, "ACGTAC", L)
    assert rc == 0 and list(st[37:40]) == [0, 1, 2]


def test_synth_generator_statistics():
    cfg = SynthConfig(L=600, seed=11)
    t = make_tables(cfg)
    st = synth_states(t, 0, 4000)
    s = st & 7
    assert abs((s == 5).mean() - cfg.n_rate) < 3e-3
    assert 0.5 * cfg.ins < ((st >> 3) & 1)[s != 7].mean() < 2 * cfg.ins
    assert 0.005 < (s == 7).any(axis=1).mean() < 0.05            # ~2 % truncated reads
    # the planted minor variants show up at roughly their mixture fraction
    fr = dict(zip(range(1, 4), cfg.minor_fracs))
    for strain, col, codon in t.truth:
        cod = 16 * s[:, col].astype(int) + 4 * s[:, col + 1] + s[:, col + 2]
        clean = (s[:, col: col + 3] < 4).all(axis=1)
        f = (cod[clean] == codon).mean()
        assert 0.4 * fr[strain] < f < 2.0 * fr[strain]
    # determinism and chunk independence
    assert np.array_equal(synth_states(t, 100, 50, chunk=7), st[100:150])


def test_start_mask_words():
    m = start_mask_words(100, [(1, 31), (35, 50)])
    bits = [j for j in range(128) if (int(m[j >> 5]) >> (j & 31)) & 1]
    assert bits == list(range(0, 30, 3)) + list(range(34, 49 - 2, 3))
    m2 = start_mask_words(100, [(1, 31)], region=(4, 20))
    bits2 = [j for j in range(128) if (int(m2[j >> 5]) >> (j & 31)) & 1]
    assert bits2 == [3, 6, 9, 12, 15]


def test_expand_cigar_randomised_against_python(mslib):
    """Random =/X/I/D/S alignments with and without a QV mask, including reads that start left of the reference
    window (negative pos is impossible in BAM, so: pos 0 with the window cut on the right) -- against a plain Python walk."""
    rng = np.random.default_rng(31)
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    for _ in range(200):
        L = int(rng.integers(20, 120))
        pos = int(rng.integers(0, L))
        cig, seq = [], []
        span = int(rng.integers(1, 100))
        used = 0
        if rng.random() < 0.3:
            cig.append((2, "S")); seq.append("GG")
        while used < span:
            u = rng.random()
            l = int(rng.integers(1, 6))
            if u < 0.15 and cig and cig[-1][1] in "=X":
                cig.append((l, "D")); used += l
            elif u < 0.3 and cig and cig[-1][1] in "=X":
                cig.append((l, "I")); seq.append("".join(rng.choice(list("ACGT"), size=l)))
            else:
                cig.append((l, "=" if u < 0.8 else "X")); seq.append("".join(rng.choice(list("ACGTN"), size=l))); used += l
        seq = "".join(seq)
        qv = None if rng.random() < 0.4 else (rng.random(len(seq)) < 0.3).astype(np.uint8)
        want = np.full(L, 7, dtype=np.uint8)
        c, q = pos, 0
        for l, o in cig:
            if o in "=X":
                for t in range(l):
                    if 0 <= c + t < L:
                        want[c + t] = 5 if (qv is not None and qv[q + t]) else code.get(seq[q + t], 5)
                c += l; q += l
            elif o == "D":
                for t in range(l):
                    if 0 <= c + t < L: want[c + t] = 4
                c += l
            elif o == "I":
                if pos <= c - 1 < L and (want[c - 1] & 7) != 7: want[c - 1] |= 8
                q += l
            else:
                q += l
        merged = []
        for l, o in cig:
            if merged and merged[-1][1] == o: merged[-1] = (merged[-1][0] + l, o)
            else: merged.append((l, o))
        rc, st, _ = _expand(mslib, merged, pos, seq, L, qv=qv)
        assert rc == 0 and np.array_equal(st, want), (merged, pos, seq)


def test_tile_layout_round_trip_and_definition(mslib):
    """plain rows <-> device tiles (csrc/rows.cuh): the C helpers, the numpy twins and the formula in the header agree"""
    from minorseq_b200.synth import tile_rows, untile_rows
    rng = np.random.default_rng(11)
    for R, L in [(1, 40), (8, 3000), (13, 100), (77, 9719), (0, 64)]:
        rw = mslib.ms_row_words(L)
        nblk = rw // 4
        rows = rng.integers(0, 2 ** 32, size=(R, rw), dtype=np.uint32)
        t = tile_rows(rows) if R else np.zeros(0, np.uint32)
        assert t.size == mslib.ms_tiled_words(L, R) == (R + 7) // 8 * 8 * rw
        t2 = np.empty_like(t)
        assert mslib.ms_tile_rows(rows.ctypes.data_as(C.c_void_p), R, L, t2.ctypes.data_as(C.c_void_p)) == 0
        assert np.array_equal(t, t2)
        back = np.empty_like(rows)
        assert mslib.ms_untile_rows(t.ctypes.data_as(C.c_void_p), R, L, back.ctypes.data_as(C.c_void_p)) == 0
        assert np.array_equal(back, rows)
        if R:
            assert np.array_equal(untile_rows(t, R, L), rows)
            t4 = t.reshape(-1, 4)
            for r, b in [(0, 0), (R - 1, nblk - 1), (R // 2, nblk // 3)]:
                assert np.array_equal(t4[((r >> 3) * nblk + b) * 8 + ((r & 7) ^ (b & 7))], rows[r, 4 * b: 4 * b + 4])
            if R % 8:    # padding slots of the last tile are "not spanned"
                assert np.array_equal(t4[((R >> 3) * nblk + 0) * 8 + ((R & 7) ^ 0)], np.array([2 ** 32 - 1] * 3 + [0], dtype=np.uint32))
