"""Host-side wall-clock split of the e2e pass (debug aid, not a benchmark): the C-ABI call, the Python wrapper around it, get_counts."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minorseq_b200 import Juliet, _lib, encode_rows, host_rows, synth_device
from minorseq_b200.synth import SynthConfig, make_tables

L, R = 3000, int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
t = make_tables(SynthConfig(L=L, seed=20240003))
j = Juliet(L, [(1, 3001)], refseq=t.refseq, mode_phasing=True, min_perc=0.5)
lib = j.lib
d = synth_device(j.hd, t, 0, R)
hdr, ev = encode_rows(host_rows(d, R, L), L, t.refseq)
del d
th = torch.from_numpy(hdr.view(np.uint8)).pin_memory()
te = torch.from_numpy(ev).pin_memory()
hp, ep = th.numpy().view(hdr.dtype), te.numpy()
j.set_base(t.refseq)
_lib.check(lib.ms_set_timing(j.hd.h, 1), j.hd.h)
for _ in range(3):
    j.run_events_host(hp, ep); j.get_counts()
real = lib.ms_juliet_pass_events_host
acc = {"c_call": [], "wrapper_total": [], "get_counts": [], "pass_device": [], "upload": []}


class Timed:
    def __call__(self, *a):
        t0 = time.perf_counter(); rc = real(*a); acc["c_call"].append((time.perf_counter() - t0) * 1e3); return rc


for _ in range(8):
    lib.ms_juliet_pass_events_host = Timed()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); j.run_events_host(hp, ep); t1 = time.perf_counter(); j.get_counts(); t2 = time.perf_counter()
    lib.ms_juliet_pass_events_host = real
    acc["wrapper_total"].append((t1 - t0) * 1e3); acc["get_counts"].append((t2 - t1) * 1e3)
    ms = C.c_double()
    for k, s in (("upload", 6), ("pass_device", 7)):
        acc[k].append(ms.value if lib.ms_stage_kernel_ms(j.hd.h, s, C.byref(ms)) == 0 else float("nan"))
for k, v in acc.items():
    print(f"{k:14s} " + " ".join(f"{x:7.3f}" for x in v))
