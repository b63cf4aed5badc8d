"""BASELINE.json configs[4]: phasing stress test, 500k reads x 2048 dense variant sites, 1/2/4/8-GPU scaling.

    python tools/c5_scaling.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/c5_scaling.py

The 500k reads are split over the ranks (strong scaling: the total is fixed).  Per stage the time is the maximum over
ranks of a CUDA-event stopwatch on the handle's stream (ms_timer_start/stop): pileup + count all-reduce, phasing
(bit-vectors, local grouping, all-gather of the ranks' compact lists, device merge + order), co-occurrence
(popcount-AND + all-reduce of the V x V matrix).  Rank 0 prints one JSON line.
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minorseq_b200 import Juliet, _lib  # noqa: E402
from minorseq_b200._lib import SynthParams  # noqa: E402
from minorseq_b200.synth import SynthConfig, make_tables  # noqa: E402

TOTAL = int(os.environ.get("C5_READS", "500000"))
REPS = 3


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if os.environ.get("BENCH_WATCHDOG"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["BENCH_WATCHDOG"]), exit=True)
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    lib = _lib.load()
    t = make_tables(SynthConfig(L=6144, seed=20240005, dense_sites=2048, dense_strains=64, n_rate=2e-5, dele=2e-5, trunc=0.0))
    sites = sorted({(c, k) for (_, c, k) in t.truth})
    per = (TOTAL + world - 1) // world
    lo, hi = rank * per, min(TOTAL, (rank + 1) * per)
    R = hi - lo
    j = Juliet(6144, [(1, 6145)], device=local, mode_phasing=True)
    if world > 1:
        assert j.hd.attach_comm()
    d = torch.empty(((R + 7) // 8 * 8, j.row_words), dtype=torch.int32, device=f"cuda:{local}")   # whole tiles (csrc/rows.cuh)
    sp = SynthParams(t.cfg.seed, 6144, t.nstrains, t.thr_N, t.thr_sub, t.thr_ins20, t.thr_trunc16)
    _lib.check(lib.ms_synth_dev(j.hd.h, C.byref(sp), t.strain_base.ctypes.data_as(C.c_void_p), t.thr_del.ctypes.data_as(C.c_void_p),
                                t.strain_cum.ctypes.data_as(C.c_void_p), lo, R, C.c_void_p(d.data_ptr())), j.hd.h)
    torch.cuda.synchronize()

    class V:
        def __init__(self, c, k):
            self.col, self.codon = c, k
    vs = [V(c, k) for c, k in sites]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def stage(fn):
        fn()                               # warm-up (buffers grow, tables find their size)
        barrier()
        ms = C.c_double()
        _lib.check(lib.ms_timer_start(j.hd.h), j.hd.h)
        for _ in range(REPS):
            out = fn()
        _lib.check(lib.ms_timer_stop(j.hd.h, C.byref(ms)), j.hd.h)
        v = torch.tensor([ms.value / REPS], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item()), out

    def pile():
        j.reset(); j.pileup_device(d.data_ptr(), R)
        _lib.check(lib.ms_allreduce_counts(j.hd.h), j.hd.h)

    ms_pile, _ = stage(pile)
    ms_phase, (hap, keys) = stage(lambda: j.phase_device(vs, d.data_ptr(), R, want_hap_id=False))
    ms_ids, (hap2, _) = stage(lambda: j.phase_device(vs, d.data_ptr(), R, want_hap_id=True))
    ms_co, Cm = stage(lambda: j.cooccurrence())
    diag = int(Cm.diagonal().sum().item())
    total = ms_pile + ms_phase + ms_co
    if rank == 0:
        print(json.dumps(dict(config="C5 phasing stress: %d reads x L=6144, V=%d dense sites, read-sharded x%d (strong scaling)" % (TOTAL, len(sites), world),
                              n_gpus=world, reads=TOTAL, reads_per_gpu=per, pileup_allreduce_ms=ms_pile, phase_ms=ms_phase,
                              phase_with_read_ids_ms=ms_ids, cooccurrence_ms=ms_co, total_ms=total, reads_per_s=TOTAL / total * 1e3,
                              distinct_patterns=hap.ndistinct, haplotypes=hap.nreported, counters=hap.counters, diag_sum=diag,
                              first_counts=[int(x) for x in hap.counts[:8]], hap_id_checksum=int(np.int64(hap2.hap_id).sum()) if world == 1 else None)),
              flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
