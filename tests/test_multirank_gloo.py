"""N>1 host logic on CPU: world_size-2 gloo runs of the two exchange steps of the sharded path
(SURVEY.md 8e): the integer all-reduce of the count tensor and the all-gather + merge of the
per-rank haplotype lists.  The per-rank inputs come from the oracle (no GPU here); the merge
code under test is the product's (minorseq_b200.api._gather_groups + ms_haplotype_order)."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle_binding
        from minorseq_b200 import _lib
        from minorseq_b200.api import _gather_groups
        from minorseq_b200._lib import PhaseCounters
        from minorseq_b200.synth import SynthConfig, make_tables, synth_states
        oracle = oracle_binding.load()
        lib = _lib.load()
        L, R = 300, 6000
        t = make_tables(SynthConfig(L=L, seed=99, n_rate=1e-3, dele=1e-3, trunc=0.05))
        per = R // world
        st = synth_states(t, rank * per, per)                       # contiguous read ranges per rank
        mask = np.zeros(L, dtype=np.uint8); mask[0:L - 2:3] = 1
        # --- exchange 1: integer sum of the count tensor
        col, codon = oracle.pileup(st, mask)
        ct = torch.from_numpy(np.concatenate([col.reshape(-1), codon.reshape(-1)]).view(np.int32).copy())
        dist.all_reduce(ct, op=dist.ReduceOp.SUM)
        # --- exchange 2: haplotype lists
        sites = sorted({(c, k) for (_, c, k) in t.truth})
        vc, vk = [s[0] for s in sites], [s[1] for s in sites]
        V = len(sites)
        bits, flags = oracle.phase_bits(st, vc, vk)
        g = oracle.phase_group(bits, flags, V, min_reads=1 << 30)   # local list, any order
        order = np.lexsort(g["patterns"].T[::-1]) if g["H"] else np.array([], dtype=int)
        pat, cnt = g["patterns"][order], g["counts"][order]          # ascending pattern, as ms_phase_groups returns
        c = g["counters"]
        marg = np.array([c["damaged"], c["gaps"], c["heteroduplex"], c["partial"]], dtype=np.int64)
        pat, cnt, marg = _gather_groups(pat, cnt, marg, device=0)
        pat, cnt = np.ascontiguousarray(pat), np.ascontiguousarray(cnt)
        Hm, nrep, c2 = C.c_int64(), C.c_int64(), PhaseCounters()
        assert lib.ms_haplotype_order(pat.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), len(cnt), V, 10,
                                      C.byref(Hm), C.byref(nrep), C.byref(c2)) == 0
        q.put((rank, ct.numpy().copy(), pat[:Hm.value].copy(), cnt[:Hm.value].copy(), nrep.value,
               (c2.reported, c2.insufficient, int(marg[0]), int(marg[1]), int(marg[2]), int(marg[3]))))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_exchange_matches_single_rank(oracle):
    from minorseq_b200.synth import SynthConfig, make_tables, synth_states
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-rank answer from the oracle on all reads
    L, R = 300, 6000
    t = make_tables(SynthConfig(L=L, seed=99, n_rate=1e-3, dele=1e-3, trunc=0.05))
    st = synth_states(t, 0, R)
    mask = np.zeros(L, dtype=np.uint8); mask[0:L - 2:3] = 1
    col, codon = oracle.pileup(st, mask)
    want_ct = np.concatenate([col.reshape(-1), codon.reshape(-1)]).view(np.int32)
    sites = sorted({(c, k) for (_, c, k) in t.truth})
    bits, flags = oracle.phase_bits(st, [s[0] for s in sites], [s[1] for s in sites])
    g = oracle.phase_group(bits, flags, len(sites))
    c = g["counters"]
    for (rank, ct, pat, cnt, nrep, ctr) in results:
        assert np.array_equal(ct, want_ct)                        # identical on every rank
        assert np.array_equal(pat, g["patterns"]) and np.array_equal(cnt, g["counts"]) and nrep == g["nreported"]
        assert ctr == (c["reported"], c["insufficient"], c["damaged"], c["gaps"], c["heteroduplex"], c["partial"])
