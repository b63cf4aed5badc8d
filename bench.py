#!/usr/bin/env python
"""bench.py -- juliet pileup + call + phase throughput (aligned CCS reads/s) on N B200s.

A "step" is one full juliet --mode-phasing pass over one batch of synthetic reads:
pivot sample -> K1 pileup -> (N>1: one NCCL all-reduce of the count tensor) -> K2 codon test
-> K3 phasing (bit-vectors, grouping, all-gather of the haplotype lists).
Reads are sharded across ranks (weak scaling: --reads-per-gpu fixed, default 1M x 3 kb, the
size BASELINE.json's target is quoted on).

  value     reads/s with the packed reads resident in HBM (CUDA events, max over ranks)
  e2e       same pass through the host-buffer entry point (pinned host -> H2D -> ... -> D2H)
  roofline  K1 (pileup_csa_kernel): algorithmic bytes L/2 per read / its CUDA-event time,
            against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the CPU restatement (oracle/, "port": the reference ships no source),
            single-threaded, on a bounded sample of the same reads, rank 0 at N=1

`--impl reference` times that CPU restatement with all host threads (unpack, pileup and the
phasing bit-vectors are OpenMP over reads; call and grouping are serial) on a bounded sample per step.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aligned CCS reads/sec pileup+call+phase"
MIN_PERC = 0.5   # juliet --min-perc (doc/JULIET.md:342-344): keeps the called set (V = 13) independent of the total read count
UNIT = "reads/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--L", type=int, default=3000)
    ap.add_argument("--seed", type=int, default=20240003)
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU sample (0 = auto, ~10 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_config(args):
    return {"workload": f"juliet --mode-phasing, synthetic HIV-like amplicon: {args.reads_per_gpu} CCS reads x {args.L} columns per GPU, "
                        "4 strains (major + 10/5/1 % minors), one gene in frame 0, reference-guided calling, --min-perc 0.5",
            "reads_per_gpu": args.reads_per_gpu, "L": args.L, "strains": 4, "seed": args.seed, "min_perc": MIN_PERC,
            "l2_policy": "packed input per GPU (%.2f GB) is larger than the 126 MB L2" % (args.reads_per_gpu * ((args.L + 31) // 32) * 16 / 1e9),
            "parallelism": f"read-sharded x{args.gpus}, one NCCL all-reduce of the count tensor"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons (B200_PROFILING.md).  The subprocess is started BEFORE CUDA / NCCL
    are initialised (forking a process that already holds NCCL state can wedge later collectives) and keeps
    sampling every 20 ms; finish(t0, t1) reports the samples that fall inside the timed window."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def run(self):
        if not self.proc:
            return
        try:
            for line in self.proc.stdout:
                self.rows.append((time.time(), [x.strip() for x in line.split(",")]))
        except Exception:
            pass

    def finish(self, t0, t1):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        inside = [r for (ts, r) in self.rows if t0 - 0.02 <= ts <= t1 + 0.02]
        sm, mx, reasons = [], [], set()
        for r in inside:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "window": "timed steps + e2e steps"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_pass(oracle, packed, L, genes, refseq, nthreads):
    """One juliet pass of the CPU restatement on packed reads; returns seconds."""
    from minorseq_b200.synth import start_mask_words
    words = start_mask_words(L, genes)
    mask = np.array([(int(words[j >> 5]) >> (j & 31)) & 1 for j in range(L)], dtype=np.uint8)
    t0 = time.perf_counter()
    st = oracle.unpack(packed, L, nthreads=nthreads)
    col, codon = oracle.pileup(st, mask, nthreads=nthreads)
    v = oracle.call(codon, genes, refseq=refseq, min_perc=MIN_PERC)
    keys = sorted({(x.col, x.codon) for x in v})
    bits, flags = oracle.phase_bits(st, [k[0] for k in keys], [k[1] for k in keys], nthreads=nthreads)
    oracle.phase_group(bits, flags, len(keys))
    return time.perf_counter() - t0


def run_reference(args, rank):
    """The CPU arm: the restatement in oracle/ (the reference ships documentation only)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding
    from minorseq_b200.synth import SynthConfig, make_tables, pack_states, synth_states
    oracle = oracle_binding.load()
    cfg = SynthConfig(L=args.L, seed=args.seed)
    t = make_tables(cfg)
    sample = args.cpu_sample or 100000
    packed = pack_states(synth_states(t, 0, sample))
    genes = [(1, args.L - args.L % 3 + 1)]
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_pass(oracle, packed, args.L, genes, t.refseq, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pass(oracle, packed, args.L, genes, t.refseq, cores)
    dt = (time.perf_counter() - t0) / args.steps
    v = sample / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 counts / f64 p-values", "data": "synthetic", "config": workload_config(args),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} reads of the same generator per step (unpack, pileup, phase bits OpenMP x{cores}; call and grouping serial)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa(local_rank):
    """Pin this rank's threads to the CPUs NVML names as local to its GPU (when the container's cpuset allows), so that
    the pinned staging buffers of the e2e path are first-touched on that GPU's NUMA node.  Returns the CPU count or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        ideal = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = ideal & allowed
        if len(cpus) >= 2 and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def main():
    args = parse()
    if os.environ.get("BENCH_WATCHDOG"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["BENCH_WATCHDOG"]), exit=True)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    affinity = bind_to_gpu_numa(local_rank) if world > 1 else None   # pinned buffers land on the GPU's NUMA node
    sampler = ClockSampler(local_rank)   # before CUDA / NCCL come up
    sampler.start()
    import torch
    import torch.distributed as dist
    from minorseq_b200 import Juliet, _lib
    from minorseq_b200._lib import SynthParams
    from minorseq_b200.synth import SynthConfig, make_tables

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    L, Rg = args.L, args.reads_per_gpu
    cfg = SynthConfig(L=L, seed=args.seed)
    t = make_tables(cfg)
    genes = [(1, L - L % 3 + 1)]
    j = Juliet(L, genes, refseq=t.refseq, device=local_rank, mode_phasing=True, min_perc=MIN_PERC)
    lib = j.lib
    if world > 1:
        j.hd.attach_comm()     # native ncclAllReduce / ncclAllGather on the handle's stream
    nw = j.row_words
    d_packed = torch.empty((Rg, nw), dtype=torch.int32, device=f"cuda:{local_rank}")
    sp = SynthParams(t.cfg.seed, L, t.nstrains, t.thr_N, t.thr_sub, t.thr_ins20, t.thr_trunc16)
    _lib.check(lib.ms_synth_dev(j.hd.h, C.byref(sp), t.strain_base.ctypes.data_as(C.c_void_p), t.thr_del.ctypes.data_as(C.c_void_p),
                                t.strain_cum.ctypes.data_as(C.c_void_p), rank * Rg, Rg, C.c_void_p(d_packed.data_ptr())), j.hd.h)
    torch.cuda.synchronize()
    _lib.check(lib.ms_set_timing(j.hd.h, 1), j.hd.h)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return j.run_device(d_packed.data_ptr(), Rg, want_hap_id=False)

    res = None
    for _ in range(max(3, args.warmup)):
        res = step()
    # ---- timed region: device-resident inputs
    k1_ms, launches0 = [], j.hd.launches
    barrier()
    t_window0 = time.time()
    _lib.check(lib.ms_timer_start(j.hd.h), j.hd.h)      # CUDA events on the stream the kernels are launched on
    for _ in range(args.steps):
        res = step()
        ms, rd = C.c_double(), C.c_int64()
        _lib.check(lib.ms_pileup_kernel_ms(j.hd.h, C.byref(ms), C.byref(rd)), j.hd.h)
        k1_ms.append(ms.value)
    el = C.c_double()
    _lib.check(lib.ms_timer_stop(j.hd.h, C.byref(el)), j.hd.h)
    barrier()
    elapsed_ms = el.value
    launches = j.hd.launches - launches0
    tm = torch.tensor([elapsed_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_per_step = float(tm.item()) / args.steps
    value = world * Rg / (ms_per_step / 1e3)

    # ---- e2e: the host-buffer entry point (pinned host -> H2D -> kernels -> D2H results)
    e2e = None
    if not args.no_e2e:
        host = torch.empty((Rg, nw), dtype=torch.int32, pin_memory=True)
        host.copy_(d_packed)
        torch.cuda.synchronize()
        hp = host.numpy().view(np.uint32)
        for _ in range(2):
            j.run_host(hp)
        barrier()
        esteps = max(2, min(args.steps, 5))
        d2h = 0
        t0 = time.perf_counter()
        for _ in range(esteps):
            r = j.run_host(hp)
            col, codon = j.get_counts()
            d2h = col.nbytes + codon.nbytes + len(r.variants) * 48 + r.haplotypes.patterns.nbytes + r.haplotypes.counts.nbytes
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / esteps], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * Rg / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(world * Rg * nw * 4),
               "d2h_bytes_per_step": int(world * d2h), "ms_per_step": float(dt.item()) * 1e3, "steps": esteps}
        del host

    clocks = sampler.finish(t_window0, time.time())

    # ---- roofline of the dominant kernel (K1)
    peak, peak_src = measured_peak()
    k1 = float(np.mean(k1_ms))
    alg_bytes = Rg * (L / 2.0)
    achieved = alg_bytes / (k1 / 1e3) / 1e9
    traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel on this workload
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_k1_traffic.json")))
        if tj["workload"] == {"reads_per_gpu": Rg, "L": L}:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    except Exception:
        pass
    roof = {"kernel": "pileup_csa_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "algorithmic_bytes": alg_bytes, "peak_source": peak_src, "algorithmic_bytes_per_read": L / 2.0, "kernel_ms": k1,
            "kernel_share_of_step": k1 / ms_per_step}

    # ---- CPU restatement beside it (rank 0, N=1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_binding
        oracle = oracle_binding.load()
        probe = d_packed[:5000].cpu().numpy().view(np.uint32)
        tp = cpu_pass(oracle, probe, L, genes, t.refseq, 1)
        sample = args.cpu_sample or int(min(Rg, max(5000, 5000 * 12.0 / max(tp, 1e-3))))
        samp = d_packed[:sample].cpu().numpy().view(np.uint32)
        ts = cpu_pass(oracle, samp, L, genes, t.refseq, 1)
        cpu = {"value": sample / ts, "unit": UNIT, "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
               "sample": f"first {sample} reads of the same device-generated batch, one pass, {ts:.1f} s"}

    if rank == 0:
        hp = res.haplotypes
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u32 counts / f64 p-values", "data": "synthetic", "config": workload_config(args),
                "clocks": clocks, "e2e": e2e, "host_affinity_cpus": affinity, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
                "result_check": {"variants": len(res.variants), "haplotypes_reported": hp.nreported if hp else None,
                                 "counters": hp.counters if hp else None}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
