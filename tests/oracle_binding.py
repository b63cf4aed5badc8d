"""ctypes binding of oracle/libms_oracle.so -- the CPU restatement used ONLY as the checker."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libms_oracle.so")


class Gene(C.Structure):
    _fields_ = [("begin", C.c_int32), ("end", C.c_int32)]


class CallParams(C.Structure):
    _fields_ = [("substitution_rate", C.c_double), ("deletion_rate", C.c_double), ("alpha", C.c_double),
                ("min_perc", C.c_double), ("max_perc", C.c_double), ("region_begin", C.c_int32),
                ("region_end", C.c_int32)]


class Variant(C.Structure):
    _fields_ = [("gene", C.c_int32), ("codon_index", C.c_int32), ("col", C.c_int32), ("ref_codon", C.c_int32),
                ("codon", C.c_int32), ("count", C.c_uint32), ("coverage", C.c_uint32), ("expected", C.c_uint32),
                ("ntests", C.c_uint32), ("pvalue", C.c_double)]


class PhaseCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("reported", "insufficient", "damaged", "gaps", "heteroduplex", "partial")]


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("L", C.c_int32), ("nstrains", C.c_int32), ("thr_N", C.c_uint32),
                ("thr_sub", C.c_uint32), ("thr_ins20", C.c_uint32), ("thr_trunc16", C.c_uint32)]


class FuseParams(C.Structure):
    _fields_ = [("min_coverage", C.c_int32), ("ins_fraction", C.c_double), ("ins_distance", C.c_int32)]


def build():
    src = [os.path.join(ORACLE_DIR, f) for f in ("ms_oracle.c", "ms_oracle.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB


_P = C.c_void_p


def _p(a):
    return None if a is None else a.ctypes.data_as(_P)


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build())
        L = self.lib
        L.mso_fisher_greater.restype = C.c_double
        L.mso_fisher_greater.argtypes = [C.c_uint32] * 4
        L.mso_codon_error_prob.restype = C.c_double
        L.mso_codon_error_prob.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
        L.mso_call.restype = C.c_int64
        L.mso_call.argtypes = [_P, C.c_int32, C.POINTER(Gene), C.c_int32, C.c_char_p, C.POINTER(CallParams),
                               C.POINTER(Variant), C.c_int64]
        L.mso_pileup.argtypes = [_P, C.c_int64, C.c_int32, _P, _P, _P, C.c_int]
        L.mso_unpack_planar_mt.argtypes = [_P, C.c_int64, C.c_int32, _P, C.c_int]
        L.mso_phase_bits_mt.argtypes = [_P, C.c_int64, C.c_int32, _P, _P, C.c_int32, _P, _P, C.c_int]
        L.mso_phase_group.restype = C.c_int64
        L.mso_phase_group.argtypes = [_P, _P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, C.c_int64,
                                      C.POINTER(C.c_int64), C.POINTER(PhaseCounters)]
        L.mso_haplotype_name.argtypes = [C.c_int64, C.c_char_p]
        L.mso_cooccurrence.argtypes = [_P, C.c_int64, C.c_int32, _P]
        L.mso_fuse.restype = C.c_int64
        L.mso_fuse.argtypes = [_P, C.c_int32, _P, _P, _P, C.c_int64, _P, C.POINTER(FuseParams), _P, C.c_int64]

    def synth_states(self, t, read0, R, nthreads=1):
        """C twin of minorseq_b200.synth.synth_states (t: SynthTables) -- the CPU arm's own workload generator."""
        p = SynthParams(t.cfg.seed, t.cfg.L, t.nstrains, t.thr_N, t.thr_sub, t.thr_ins20, t.thr_trunc16)
        out = np.empty((R, t.cfg.L), dtype=np.uint8)
        self.lib.mso_synth_states(C.byref(p), _p(np.ascontiguousarray(t.strain_base)), _p(np.ascontiguousarray(t.thr_del)),
                                  _p(np.ascontiguousarray(t.strain_cum)), C.c_int64(read0), C.c_int64(R), _p(out), nthreads)
        return out

    def unpack(self, packed, L, nthreads=1):
        R = packed.shape[0]
        out = np.empty((R, L), dtype=np.uint8)
        self.lib.mso_unpack_planar_mt(_p(np.ascontiguousarray(packed)), R, L, _p(out), nthreads)
        return out

    def pileup(self, states, start_mask_bytes=None, codons=True, nthreads=1):
        R, L = states.shape
        col = np.zeros((L, 8), dtype=np.uint32)
        codon = np.zeros((L, 64), dtype=np.uint32) if codons else None
        states = np.ascontiguousarray(states)
        self.lib.mso_pileup(_p(states), R, L, _p(start_mask_bytes), _p(col), _p(codon), nthreads)
        return col, codon

    def fisher(self, a, b, c, d):
        return self.lib.mso_fisher_greater(a, b, c, d)

    def call(self, codon, genes, refseq=None, sub=5e-4, dele=3e-3, alpha=0.01, min_perc=-1.0, max_perc=-1.0,
             region=(0, 0), cap=1 << 16):
        L = codon.shape[0]
        g = (Gene * len(genes))(*[Gene(b, e) for (b, e) in genes])
        prm = CallParams(sub, dele, alpha, min_perc, max_perc, region[0], region[1])
        out = (Variant * cap)()
        n = self.lib.mso_call(_p(np.ascontiguousarray(codon)), L, g, len(genes),
                              refseq.encode() if refseq else None, C.byref(prm), out, cap)
        assert n <= cap
        return [out[i] for i in range(n)]

    def phase_bits(self, states, var_col, var_codon, nthreads=1):
        R, L = states.shape
        V = len(var_col)
        nw = max(1, (V + 31) // 32)
        bits = np.zeros((R, nw), dtype=np.uint32)
        flags = np.zeros(R, dtype=np.uint8)
        vc = np.ascontiguousarray(var_col, dtype=np.int32)
        vd = np.ascontiguousarray(var_codon, dtype=np.int32)
        if V == 0:
            bits[:] = 0
        self.lib.mso_phase_bits_mt(_p(np.ascontiguousarray(states)), R, L, _p(vc), _p(vd), V, _p(bits), _p(flags), nthreads)
        if V == 0:
            bits = np.zeros((R, 1), dtype=np.uint32)
        return bits, flags

    def phase_group(self, bits, flags, V, min_reads=10):
        R = bits.shape[0]
        nw = max(1, (V + 31) // 32)
        # the C side uses ceil(V/32) words; with V == 0 there are zero words per read
        cbits = np.ascontiguousarray(bits[:, : (V + 31) // 32]) if V > 0 else np.zeros((R, 0), dtype=np.uint32)
        hap = np.empty(R, dtype=np.int32)
        cap = max(1, R)
        pat = np.zeros((cap, max(1, (V + 31) // 32)), dtype=np.uint32)
        cnt = np.zeros(cap, dtype=np.uint64)
        nrep = C.c_int64()
        ctr = PhaseCounters()
        H = self.lib.mso_phase_group(_p(cbits), _p(np.ascontiguousarray(flags)), R, V, min_reads, _p(hap), _p(pat),
                                     _p(cnt), cap, C.byref(nrep), C.byref(ctr))
        pat = pat[:H].reshape(H, pat.shape[1])
        if V == 0:
            pat = np.zeros((H, nw), dtype=np.uint32)
        return dict(H=H, nreported=nrep.value, patterns=pat, counts=cnt[:H], hap_id=hap,
                    counters={k: getattr(ctr, k) for k, _ in PhaseCounters._fields_})

    def hap_name(self, i):
        buf = C.create_string_buffer(3)
        self.lib.mso_haplotype_name(i, buf)
        return buf.value.decode()

    def cooccurrence(self, bits, V):
        Cm = np.zeros((V, V), dtype=np.int32)
        self.lib.mso_cooccurrence(_p(np.ascontiguousarray(bits)), bits.shape[0], V, _p(Cm))
        return Cm

    # ---- cleric restatement
    def nw_align(self, a: str, b: str):
        cap = len(a) + len(b) + 1
        ops = C.create_string_buffer(cap)
        score = C.c_int64()
        self.lib.mso_nw_align.restype = C.c_int64
        n = self.lib.mso_nw_align(a.encode(), len(a), b.encode(), len(b), ops, cap, C.byref(score))
        assert n >= 0
        return ops.raw[:n].decode(), score.value

    def project_read(self, ops: str, b: str, pos: int, cigar, seq: str):
        """cigar: list of (len, op-char).  Returns (new_pos, new cigar list) or None when nothing lands on B."""
        co = "".join(o for _, o in cigar).encode()
        cl = np.array([l for l, _ in cigar], dtype=np.int32)
        cap = len(seq) + len(ops) + 8
        no = C.create_string_buffer(cap)
        nl = np.zeros(cap, dtype=np.int32)
        npos = C.c_int32(-1)
        self.lib.mso_project_read.restype = C.c_int64
        n = self.lib.mso_project_read(ops.encode(), len(ops), b.encode(), len(b), pos, co, _p(cl), len(cigar), seq.encode(), len(seq),
                                      no, _p(nl), cap, C.byref(npos))
        if n < 0:
            raise ValueError(f"mso_project_read failed ({n})")
        if n == 0:
            return None
        return npos.value, [(int(nl[i]), no.raw[i:i + 1].decode()) for i in range(n)]

    def fuse(self, col, ins_col=None, ins_off=None, ins_len=None, pool=b"", min_coverage=50, ins_fraction=0.5,
             ins_distance=20):
        L = col.shape[0]
        nins = 0 if ins_col is None else len(ins_col)
        ic = np.ascontiguousarray(ins_col, dtype=np.int32) if nins else None
        io = np.ascontiguousarray(ins_off, dtype=np.int64) if nins else None
        il = np.ascontiguousarray(ins_len, dtype=np.int32) if nins else None
        pl = np.frombuffer(pool, dtype=np.uint8) if nins else None
        cap = L + (int(il.sum()) if nins else 0) + 16
        seq = np.empty(cap, dtype=np.uint8)
        prm = FuseParams(min_coverage, ins_fraction, ins_distance)
        n = self.lib.mso_fuse(_p(np.ascontiguousarray(col)), L, _p(ic), _p(io), _p(il), nins, _p(pl), C.byref(prm),
                              _p(seq), cap)
        return seq[:n].tobytes().decode()


_oracle = None


def load():
    global _oracle
    if _oracle is None:
        _oracle = Oracle()
    return _oracle
