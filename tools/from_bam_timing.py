"""R3: wall-clock of the juliet / fuse binaries starting from a BAM file (inflate + CIGAR walk + H2D + kernels +
report).  Host-bound by construction; reported, not a target."""
import json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from minorseq_b200.synth import SynthConfig, make_tables, pack_states, synth_states
BIN = os.path.join(ROOT, "minorseq_b200", "bin")
R = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 5            # runs per variant (min and median are reported)
VARIANTS = (("", {}),) if len(sys.argv) > 3 and sys.argv[3] == "overlap-only" else (("", {}), ("_serial_load", {"MS_SERIAL_LOAD": "1"}))
t = make_tables(SynthConfig(L=3000, seed=20240003))
d = tempfile.mkdtemp()
packed = pack_states(synth_states(t, 0, R))
packed.tofile(os.path.join(d, "p.bin"))
open(os.path.join(d, "ref.txt"), "w").write(t.refseq + "\n")
subprocess.check_call([os.path.join(BIN, "packed2bam"), os.path.join(d, "p.bin"), "3000", str(R), os.path.join(d, "ref.txt"), os.path.join(d, "in.bam")])
cfg = {"genes": [{"name": "pol", "begin": 1, "end": 3001}], "referenceName": "synthetic_ref", "referenceSequence": t.refseq}
json.dump(cfg, open(os.path.join(d, "cfg.json"), "w"))
out = {"reads": R, "bam_bytes": os.path.getsize(os.path.join(d, "in.bam"))}
for name, cmd in (("juliet", [os.path.join(BIN, "juliet"), "-c", os.path.join(d, "cfg.json"), "--mode-phasing", "--min-perc", "0.5", os.path.join(d, "in.bam"), os.path.join(d, "o.json")]),
                  ("fuse", [os.path.join(BIN, "fuse"), os.path.join(d, "in.bam"), os.path.join(d, "o.fasta")])):
    for tag, extra in VARIANTS:   # context/decode overlap vs serial, same box
        ts = []
        for _ in range(REPS):
            t0 = time.perf_counter(); subprocess.check_call(cmd, env=dict(os.environ, MS_TIMING="1", **extra)); ts.append(time.perf_counter() - t0)
        out[name + tag + "_s"] = min(ts); out[name + tag + "_median_s"] = sorted(ts)[len(ts) // 2]
    out[name + "_reads_per_s"] = R / out[name + "_s"]
print(json.dumps(out))
