"""cleric's alignment step: GPU Needleman-Wunsch (ms_align_refs) against the CPU restatement on HIV-genome-sized references,
and the cleric binary on a synthetic BAM.    python tools/cleric_timing.py [L]"""
import ctypes as C, json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bam_util, oracle_binding
from minorseq_b200 import Handle, _lib
L = int(sys.argv[1]) if len(sys.argv) > 1 else 9719
rng = np.random.default_rng(1)
a = "".join(rng.choice(list("ACGT"), size=L))
b = []
for ch in a:
    u = rng.random()
    if u < 0.01: continue
    if u < 0.02: b.append(str(rng.choice(list("ACGT"))))
    b.append(ch if rng.random() > 0.05 else str(rng.choice(list("ACGT"))))
b = "".join(b)
lib = _lib.load(); hd = Handle(0)
cap = len(a) + len(b) + 1
ops = C.create_string_buffer(cap); n, score = C.c_int64(), C.c_int64()
ts = []
for _ in range(4):
    t0 = time.perf_counter()
    _lib.check(lib.ms_align_refs(hd.h, a.encode(), len(a), b.encode(), len(b), ops, cap, C.byref(n), C.byref(score)), hd.h)
    ts.append(time.perf_counter() - t0)
gpu_ops = ops.raw[: n.value].decode()
oracle = oracle_binding.load()
t0 = time.perf_counter(); cpu_ops, cpu_score = oracle.nw_align(a, b); t_cpu = time.perf_counter() - t0
out = dict(la=len(a), lb=len(b), cells=len(a) * len(b), gpu_ms_incl_traceback_and_25MB_readback=min(ts[1:]) * 1e3, cpu_restatement_ms=t_cpu * 1e3,
           same_path=gpu_ops == cpu_ops, same_score=score.value == cpu_score, gcups_gpu=len(a) * len(b) / min(ts[1:]) / 1e9)
# the binary on 20 k reads of ~3 kb
R = 20000
d = tempfile.mkdtemp()
recs = []
for r in range(R):
    pos = int(rng.integers(0, L - 3000)); seq = a[pos:pos + 3000]
    recs.append(bam_util.record(f"r/{r}/ccs", 0, pos, [(3000, "=")], seq))
bam_util.write_bam(os.path.join(d, "in.bam"), "orig", L, recs)
open(os.path.join(d, "a.fa"), "w").write(">orig\n" + a + "\n"); open(os.path.join(d, "b.fa"), "w").write(">target\n" + b + "\n")
t0 = time.perf_counter()
subprocess.check_call([os.path.join(ROOT, "minorseq_b200", "bin", "cleric"), os.path.join(d, "in.bam"), os.path.join(d, "a.fa"), os.path.join(d, "b.fa"), os.path.join(d, "out.bam")])
out["cleric_cli_s_20k_reads_x_3kb"] = time.perf_counter() - t0
print(json.dumps(out))
