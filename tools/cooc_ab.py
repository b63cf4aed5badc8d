"""A/B of the two co-occurrence kernels (popcount-AND vs tcgen05 u8 MMA): exact equality of C and timings.
    python tools/cooc_ab.py [R V ...]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minorseq_b200 import Handle, Juliet, _lib  # noqa: E402

lib = _lib.load()
hd = Handle(0)
cases = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)] or [(5000, 300), (40000, 256), (100000, 700), (500000, 2048)]


def run(variant, reps):
    _lib.check(lib.ms_set_cooccurrence_variant(hd.h, variant), hd.h)
    p = C.c_void_p()
    _lib.check(lib.ms_cooccurrence(hd.h, C.byref(p)), hd.h)
    _lib.check(lib.ms_synchronize(hd.h), hd.h)
    ms = C.c_double()
    _lib.check(lib.ms_timer_start(hd.h), hd.h)
    for _ in range(reps):
        _lib.check(lib.ms_cooccurrence(hd.h, C.byref(p)), hd.h)
    _lib.check(lib.ms_timer_stop(hd.h, C.byref(ms)), hd.h)
    from minorseq_b200.api import _as_tensor
    return ms.value / reps, _as_tensor(p.value, (V * V,), torch.int32, 0).clone()


for R, V in cases:
    # random bit matrix straight into the handle's phasing buffers: L is irrelevant here
    nw = (V + 31) // 32
    g = torch.Generator(device="cuda"); g.manual_seed(R + V)
    bits = torch.randint(-2**31, 2**31 - 1, (R, nw), dtype=torch.int32, device="cuda", generator=g)
    if V % 32:
        bits[:, -1] &= (1 << (V % 32)) - 1
    _lib.check(lib.ms_set_layout(hd.h, 3 * V + 3, None), hd.h)
    vc = np.arange(V, dtype=np.int32) * 3
    vk = np.zeros(V, dtype=np.int32)
    _lib.check(lib.ms_phase_begin(hd.h, vc.ctypes.data_as(C.c_void_p), vk.ctypes.data_as(C.c_void_p), V, R), hd.h)
    rows = torch.zeros(((R + 7) // 8 * 8, lib.ms_row_words(3 * V + 3)), dtype=torch.int32, device="cuda")
    _lib.check(lib.ms_phase_dev(hd.h, C.c_void_p(rows.data_ptr()), R), hd.h)      # sizes the buffers, sets phase_n
    pb, pf, pn = C.c_void_p(), C.c_void_p(), C.c_int64()
    _lib.check(lib.ms_phase_device(hd.h, C.byref(pb), C.byref(pf), C.byref(pn)), hd.h)
    _lib.check(lib.ms_synchronize(hd.h), hd.h)
    from minorseq_b200.api import _as_tensor
    _as_tensor(pb.value, (R * nw,), torch.int32, 0).copy_(bits.reshape(-1))
    torch.cuda.synchronize()
    t1, c1 = run(1, 3)
    t2, c2 = run(2, 3)
    # independent check of a corner of C with torch
    sub = min(V, 64)
    b = ((bits[:, :2].to(torch.int64).unsqueeze(-1) >> torch.arange(32, device="cuda")) & 1).reshape(R, -1)[:, :sub].to(torch.float64)
    ref = (b.T @ b).to(torch.int32)
    ok_ref = bool(torch.equal(c1.view(V, V)[:sub, :sub], ref))
    print(json.dumps(dict(R=R, V=V, popcount_ms=t1, tensor_ms=t2, equal=bool(torch.equal(c1, c2)), popcount_matches_torch=ok_ref,
                          mismatches=int((c1 != c2).sum().item()), speedup=t1 / t2)), flush=True)
