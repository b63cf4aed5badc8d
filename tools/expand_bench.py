"""Times expand_events_kernel alone (device-resident event rows, CUDA events on the handle's stream) and the e2e pass
from pinned host event rows, 1 M x 3 kb.  Usage: python tools/expand_bench.py [reads] [L]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minorseq_b200 import Juliet, _lib, encode_rows
from minorseq_b200._lib import SynthParams
from minorseq_b200.synth import SynthConfig, make_tables

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
lib = _lib.load()
t = make_tables(SynthConfig(L=L, seed=20240003))
j = Juliet(L, [(1, L - L % 3 + 1)], refseq=t.refseq, mode_phasing=True, min_perc=0.5)
nw = j.row_words
d = torch.empty(((R + 7) // 8 * 8, nw), dtype=torch.int32, device="cuda")   # whole tiles (csrc/rows.cuh)
sp = SynthParams(t.cfg.seed, L, t.nstrains, t.thr_N, t.thr_sub, t.thr_ins20, t.thr_trunc16)
_lib.check(lib.ms_synth_dev(j.hd.h, C.byref(sp), t.strain_base.ctypes.data_as(C.c_void_p), t.thr_del.ctypes.data_as(C.c_void_p),
                            t.strain_cum.ctypes.data_as(C.c_void_p), 0, R, C.c_void_p(d.data_ptr())), j.hd.h)
from minorseq_b200 import host_rows
rows = host_rows(d, R, L)
t0 = time.perf_counter()
hdr, ev = encode_rows(rows, L, t.refseq)
print(f"host encode: {time.perf_counter() - t0:.2f} s single thread, {len(ev) / R:.1f} event bytes + 8 header bytes = {(hdr.nbytes + ev.nbytes) / R:.1f} B/read")
j.set_base(t.refseq)
dh = torch.from_numpy(hdr.view(np.uint8)).cuda()
de = torch.from_numpy(ev).cuda()
out = torch.zeros_like(d)
ms = C.c_double()
for rep in range(3):
    _lib.check(lib.ms_timer_start(j.hd.h), j.hd.h)
    _lib.check(lib.ms_expand_events_dev(j.hd.h, C.c_void_p(dh.data_ptr()), C.c_void_p(de.data_ptr()), R, C.c_void_p(out.data_ptr())), j.hd.h)
    _lib.check(lib.ms_timer_stop(j.hd.h, C.byref(ms)), j.hd.h)
print(f"expand_events_kernel: {ms.value:.3f} ms for {R} reads = {R * nw * 4 / ms.value / 1e6:.0f} GB/s written, equal to rows: {bool(torch.equal(out, d))}")
th = torch.from_numpy(hdr.view(np.uint8)).pin_memory()
te = torch.from_numpy(ev).pin_memory()
hp, ep = th.numpy().view(hdr.dtype), te.numpy()
for mb in (os.environ.get("MS_EVENTS_CHUNK_MB", "24"),):
    for _ in range(2):
        j.run_events_host(hp, ep)
    t0 = time.perf_counter()
    for _ in range(5):
        j.run_events_host(hp, ep)
        j.get_counts()
    print(f"e2e from pinned event rows (chunk {mb} MB): {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per pass")
