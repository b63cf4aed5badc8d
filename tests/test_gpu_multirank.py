"""Read-sharded 2-GPU run (NCCL) must reproduce the 1-GPU result bit for bit.
Skipped on boxes with fewer than 2 GPUs; run with `gpurun --gpus 2`."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L, R = 3000, 60000


def _pass(device, read0, nreads, native=False):
    sys.path.insert(0, ROOT)
    import ctypes as C
    from minorseq_b200 import Juliet, _lib
    from minorseq_b200._lib import SynthParams
    from minorseq_b200.synth import SynthConfig, make_tables
    t = make_tables(SynthConfig(L=L, seed=20240003, n_rate=2e-3))
    j = Juliet(L, [(1, 3001), (2, 3000)], refseq=t.refseq, device=device, mode_phasing=True)
    if native:
        assert j.hd.attach_comm()
    from minorseq_b200 import synth_device
    d = synth_device(j.hd, t, read0, nreads)
    res = j.run_device(d.data_ptr(), nreads, want_hap_id=True)
    col, codon = j.get_counts()
    v = [(x.gene, x.col, x.codon, x.count, x.coverage, x.expected, x.ntests, x.pvalue) for x in res.variants]
    h = res.haplotypes
    return dict(col=col, codon=codon, variants=v, patterns=h.patterns, counts=h.counts, nreported=h.nreported,
                counters=h.counters, hap_id=h.hap_id)


def _worker(rank, world, port, q, native):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        per = R // world
        q.put((rank, _pass(rank, rank * per, per, native)))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("native", [True, False], ids=["library-nccl", "torch-distributed"])
def test_two_gpus_equal_one_gpu(native):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, native)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=500) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    one = _pass(0, 0, R)
    for r in range(world):
        m = out[r]
        assert np.array_equal(m["col"], one["col"]) and np.array_equal(m["codon"], one["codon"])   # all-reduced counts
        assert m["variants"] == one["variants"]                                                   # replicated K2, bit-identical p
        assert np.array_equal(m["patterns"], one["patterns"]) and np.array_equal(m["counts"], one["counts"])
        assert m["nreported"] == one["nreported"] and m["counters"] == one["counters"]
    hap = np.concatenate([out[r]["hap_id"] for r in range(world)])
    assert np.array_equal(hap, one["hap_id"])


# ---- phasing merge with thousands of distinct patterns (device-side merge + bitonic order), BASELINE configs[4] shape ----
def _phase_many(device, lo, hi, native, total=16000):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from minorseq_b200 import Juliet
    from minorseq_b200.synth import pack_states
    from test_gpu_parity import _many_pattern_states

    class V:
        def __init__(self, c, k):
            self.col, self.codon = c, k
    st, cols, cods = _many_pattern_states(total, 900, 70, 25, seed=99)
    j = Juliet(900, [(1, 901)], device=device, mode_phasing=True)
    if native:
        assert j.hd.attach_comm()
    from minorseq_b200 import device_rows
    d = device_rows(pack_states(st[lo:hi]), device)
    hap, keys = j.phase_device([V(c, k) for c, k in zip(cols, cods)], d.data_ptr(), hi - lo, cap=6000)
    return dict(patterns=hap.patterns, counts=hap.counts, nreported=hap.nreported, counters=hap.counters, hap_id=hap.hap_id,
                ndistinct=hap.ndistinct)


def _worker_many(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        per = 16000 // world
        q.put((rank, _phase_many(rank, rank * per, (rank + 1) * per, True)))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpus_many_patterns_equal_one_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_many, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=500) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    one = _phase_many(0, 0, 16000, False)
    assert one["ndistinct"] > 8192          # the merged list needs the sort path, each rank's own list more than the first capacity guess
    for r in range(world):
        m = out[r]
        assert m["ndistinct"] == one["ndistinct"] and m["nreported"] == one["nreported"] and m["counters"] == one["counters"]
        assert np.array_equal(m["patterns"], one["patterns"]) and np.array_equal(m["counts"], one["counts"])
    assert np.array_equal(np.concatenate([out[r]["hap_id"] for r in range(world)]), one["hap_id"])
