"""Host-side wall-clock breakdown of one juliet step (debug aid, not a benchmark)."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minorseq_b200 import Juliet, _lib
from minorseq_b200._lib import SynthParams
from minorseq_b200.synth import SynthConfig, make_tables

L, R = 3000, int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
t = make_tables(SynthConfig(L=L, seed=20240003))
j = Juliet(L, [(1, 3001)], refseq=t.refseq, mode_phasing=True)
d = torch.empty(((R + 7) // 8 * 8, j.row_words), dtype=torch.int32, device="cuda")   # whole tiles (csrc/rows.cuh)
sp = SynthParams(t.cfg.seed, L, t.nstrains, t.thr_N, t.thr_sub, t.thr_ins20, t.thr_trunc16)
_lib.check(j.lib.ms_synth_dev(j.hd.h, C.byref(sp), t.strain_base.ctypes.data_as(C.c_void_p), t.thr_del.ctypes.data_as(C.c_void_p),
                              t.strain_cum.ctypes.data_as(C.c_void_p), 0, R, C.c_void_p(d.data_ptr())), j.hd.h)
torch.cuda.synchronize()


def tick(name, f, acc):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize()
    acc.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
    return r


acc = {}
for it in range(6):
    tick("reset", j.reset, acc)
    tick("pileup", lambda: j.pileup_device(d.data_ptr(), R), acc)
    v = tick("call", j.call, acc)
    tick("phase", lambda: j.phase_device(v, d.data_ptr(), R, want_hap_id=False), acc)
for k, x in acc.items():
    print(f"{k:8s} " + " ".join(f"{y:8.3f}" for y in x))
