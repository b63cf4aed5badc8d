"""A/B: the shared-memory-atomic histogram (north_star's sketch) vs the carry-save kernel on 200k x 3 kb."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minorseq_b200 import Handle, Juliet, _lib
from minorseq_b200._lib import SynthParams
from minorseq_b200.synth import SynthConfig, make_tables
lib = _lib.load(); hd = Handle(0)
R, L = 200_000, 3000
t = make_tables(SynthConfig(L=L, seed=20240003))
d = torch.empty(((R + 7) // 8 * 8, lib.ms_row_words(L)), dtype=torch.int32, device="cuda")   # whole tiles (csrc/rows.cuh)
sp = SynthParams(t.cfg.seed, L, t.nstrains, t.thr_N, t.thr_sub, t.thr_ins20, t.thr_trunc16)
_lib.check(lib.ms_synth_dev(hd.h, C.byref(sp), t.strain_base.ctypes.data_as(C.c_void_p), t.thr_del.ctypes.data_as(C.c_void_p),
                            t.strain_cum.ctypes.data_as(C.c_void_p), 0, R, C.c_void_p(d.data_ptr())), hd.h)
j = Juliet(L, [(1, 3001)], handle=hd)
res = {}
for variant in (0, 1):
    _lib.check(lib.ms_set_pileup_variant(hd.h, variant), hd.h)
    ts = []
    for _ in range(4):
        j.reset(); lib.ms_synchronize(hd.h)
        ms = C.c_double(); lib.ms_timer_start(hd.h)
        j.pileup_device(d.data_ptr(), R)
        lib.ms_timer_stop(hd.h, C.byref(ms)); ts.append(ms.value)
    res[variant] = (float(np.median(ts[1:])), j.get_counts())
_lib.check(lib.ms_set_pileup_variant(hd.h, 0), hd.h)
a, b = res[0][1], res[1][1]
same = np.array_equal(a[1], b[1]) and np.array_equal(a[0][:, :6], b[0][:, :6])
print("csa    %.3f ms  %.0f GB/s" % (res[0][0], R * L / 2 / res[0][0] / 1e6))
print("atomic %.3f ms  %.0f GB/s   identical counts: %s" % (res[1][0], R * L / 2 / res[1][0] / 1e6, same))
