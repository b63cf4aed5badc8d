#!/bin/bash
# One GPU box, N=1: the GPU test suite, the bench lines of the BASELINE configs and the ncu launch list of a 2-step bench
# run (outputs under gpurun_out/<tag>_*).  usage: bash tools/final_measure.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/${tag}_gputests.log
for c in T C4 C5; do
  python bench.py --config $c > gpurun_out/${tag}_bench_n1_$c.json 2> gpurun_out/${tag}_bench_n1_$c.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
python - <<PY
import json
for c in ("T", "C4", "C5"):
    try:
        d = json.loads(open("gpurun_out/${tag}_bench_n1_%s.json" % c).read().strip().splitlines()[-1]); e = d["e2e"]
        print(c, round(d["ms_per_step"], 4), round(d["roofline"]["frac"], 3),
              {k: (round(e[k], 3) if isinstance(e[k], float) else e[k]) for k in ("ms_per_step", "upload_ms", "pass_device_ms", "plain_h2d_copy_ms", "h2d_bytes_per_read")},
              [(k["kernel"], round(k.get("kernel_ms") or 0, 3), round(k.get("frac") or 0, 3)) for k in d["roofline"]["kernels"]])
    except Exception as ex:
        print(c, "failed:", ex)
PY
