"""A/B of the count tensor's cross-GPU sum (SURVEY.md 5 asks to measure it): ncclAllReduce against one-shot / multimem
all-reduces over symmetric (peer-mapped, NVLS-capable) memory, at the path's message sizes (0.86 MB at 3 kb, 2.8 MB at 9.7 kb).
Run under torch.distributed.run, one rank per GPU:
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/allreduce_ab.py
The symmetric-memory kernels are PyTorch's (library code): this is a measurement that decides whether a native peer-memory
exchange would pay, not a product path.  Integer sums are the requirement (bit-exact counts), so float variants are timed for
latency reference only."""
import json, os, sys
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
dev = f"cuda:{local}"
out = {"world": world}


def timeit(fn, iters=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())          # microseconds per call, max over ranks


for name, L in (("3kb", 3000), ("9.7kb", 9719)):
    n = L * 72
    res = {"bytes": n * 4}
    a = torch.ones(n, dtype=torch.int32, device=dev)
    res["nccl_allreduce_i32_us"] = timeit(lambda: dist.all_reduce(a))
    f = torch.ones(n, dtype=torch.float32, device=dev)
    res["nccl_allreduce_f32_us"] = timeit(lambda: dist.all_reduce(f))
    g = torch.empty(world * 4096, dtype=torch.int32, device=dev)
    s = torch.ones(4096, dtype=torch.int32, device=dev)
    res["nccl_allgather_16KBx_us"] = timeit(lambda: dist.all_gather_into_tensor(g, s))
    try:
        import torch.distributed._symmetric_memory as symm
        gname = dist.group.WORLD.group_name
        for dt, key in ((torch.int32, "i32"), (torch.float32, "f32")):
            try:
                t = symm.empty(n, dtype=dt, device=dev)
                symm.rendezvous(t, gname)
                t.fill_(1)
                res[f"symm_one_shot_{key}_us"] = timeit(lambda: torch.ops.symm_mem.one_shot_all_reduce(t, "sum", gname))
                chk = torch.ops.symm_mem.one_shot_all_reduce(t, "sum", gname)
                res[f"symm_one_shot_{key}_ok"] = bool((chk == world).all().item())
            except Exception as e:
                res[f"symm_one_shot_{key}_error"] = str(e)[:200]
            try:
                t2 = symm.empty(n, dtype=dt, device=dev)
                symm.rendezvous(t2, gname)
                t2.fill_(0)
                res[f"symm_multimem_{key}_us"] = timeit(lambda: torch.ops.symm_mem.multimem_all_reduce_(t2, "sum", gname))
            except Exception as e:
                res[f"symm_multimem_{key}_error"] = str(e)[:200]
            try:
                t3 = symm.empty(n, dtype=dt, device=dev)
                symm.rendezvous(t3, gname)
                t3.fill_(0)
                res[f"symm_two_shot_{key}_us"] = timeit(lambda: torch.ops.symm_mem.two_shot_all_reduce_(t3, "sum", gname))
            except Exception as e:
                res[f"symm_two_shot_{key}_error"] = str(e)[:200]
    except Exception as e:
        res["symm_error"] = str(e)[:300]
    out[name] = res
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
