#!/bin/bash
# A/B of the device-resident step between the in-tree library and another build of the same ABI (MS_LIB_PATH) at N ranks on one box.
# usage: bash tools/ab_step_n.sh <N> <other.so>
N=${1:-2}; OTHER=$2
run() {
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $N --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-digest 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],4), d.get('exchanges'))"
}
for rep in 1 2; do
  unset MS_LIB_PATH; run in-tree 2951$rep
  export MS_LIB_PATH=$OTHER; run other 2952$rep
done
