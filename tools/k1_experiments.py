"""K1 what-if timings on 1M x 3 kb: fuse mode (no codon work), juliet mode with and without exceptions."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minorseq_b200 import Fuse, Handle, Juliet, _lib
from minorseq_b200._lib import SynthParams
from minorseq_b200.synth import SynthConfig, make_tables
lib = _lib.load()
hd = Handle(0)
_lib.check(lib.ms_set_timing(hd.h, 1), hd.h)
R, L = 1_000_000, 3000


def synth(t):
    d = torch.empty(((R + 7) // 8 * 8, lib.ms_row_words(L)), dtype=torch.int32, device="cuda")   # whole tiles (csrc/rows.cuh)
    sp = SynthParams(t.cfg.seed, L, t.nstrains, t.thr_N, t.thr_sub, t.thr_ins20, t.thr_trunc16)
    _lib.check(lib.ms_synth_dev(hd.h, C.byref(sp), t.strain_base.ctypes.data_as(C.c_void_p), t.thr_del.ctypes.data_as(C.c_void_p),
                                t.strain_cum.ctypes.data_as(C.c_void_p), 0, R, C.c_void_p(d.data_ptr())), hd.h)
    return d


def k1(obj, d, reps=8):
    ts = []
    for _ in range(reps):
        obj.reset(); obj.pileup_device(d.data_ptr(), R)
        ms, rd = C.c_double(), C.c_int64()
        _lib.check(lib.ms_pileup_kernel_ms(hd.h, C.byref(ms), C.byref(rd)), hd.h)
        ts.append(ms.value)
    return float(np.median(ts[2:]))


t = make_tables(SynthConfig(L=L, seed=20240003))
d = synth(t)
print("juliet default        %.4f ms" % k1(Juliet(L, [(1, 3001)], handle=hd), d))
print("juliet 3 frames       %.4f ms" % k1(Juliet(L, [(1, 3001), (2, 3001), (3, 3001)], handle=hd), d))
print("fuse (no codons)      %.4f ms" % k1(Fuse(L, handle=hd), d))
jb = Juliet(L, [(1, 3001)], handle=hd); jb.set_count_insertions(True)
print("juliet + ins (8 masks) %.4f ms" % k1(jb, d)); jb.set_count_insertions(False)
t0 = make_tables(SynthConfig(L=L, seed=20240003, sub=0.0, minor_fracs=(), variants_per_minor=(0, 0)))
d0 = synth(t0)
print("juliet no exceptions  %.4f ms" % k1(Juliet(L, [(1, 3001)], handle=hd), d0))
