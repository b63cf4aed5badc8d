"""K1 time against the number of column segments (MS_K1_NSEG) for one reference length; one process per setting.
    python tools/k1_segments.py L R [frames [nseg,nseg,...]]      # prints one line per nseg"""
import ctypes as C, os, subprocess, sys
if os.environ.get("K1SEG_CHILD"):
    import numpy as np, torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from minorseq_b200 import Handle, Juliet, _lib
    from minorseq_b200._lib import SynthParams
    from minorseq_b200.synth import SynthConfig, make_tables
    L, R, frames = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    lib = _lib.load()
    hd = Handle(0)
    _lib.check(lib.ms_set_timing(hd.h, 1), hd.h)
    t = make_tables(SynthConfig(L=L, seed=20240004))
    d = torch.empty(((R + 7) // 8 * 8, lib.ms_row_words(L)), dtype=torch.int32, device="cuda")   # whole tiles (csrc/rows.cuh)
    sp = SynthParams(t.cfg.seed, L, t.nstrains, t.thr_N, t.thr_sub, t.thr_ins20, t.thr_trunc16)
    _lib.check(lib.ms_synth_dev(hd.h, C.byref(sp), t.strain_base.ctypes.data_as(C.c_void_p), t.thr_del.ctypes.data_as(C.c_void_p),
                                t.strain_cum.ctypes.data_as(C.c_void_p), 0, R, C.c_void_p(d.data_ptr())), hd.h)
    try:
        j = Juliet(L, [(f + 1, L + 1) for f in range(frames)], handle=hd)
    except Exception as e:
        print("nseg", os.environ.get("MS_K1_NSEG"), "not possible:", str(e)[:60]); sys.exit(0)
    ts = []
    for _ in range(8):
        j.reset(); j.pileup_device(d.data_ptr(), R)
        ms, rd = C.c_double(), C.c_int64()
        _lib.check(lib.ms_pileup_kernel_ms(hd.h, C.byref(ms), C.byref(rd)), hd.h)
        ts.append(ms.value)
    k1 = float(np.median(ts[2:]))
    print("L %d frames %d nseg %s  K1 %.4f ms  %.0f GB/s" % (L, frames, os.environ.get("MS_K1_NSEG", "auto"), k1, R * L / 2 / k1 / 1e6), flush=True)
else:
    L, R = sys.argv[1], sys.argv[2]
    frames = sys.argv[3] if len(sys.argv) > 3 else "1"
    for nseg in (sys.argv[4].split(",") if len(sys.argv) > 4 else ["auto", "1", "2", "3", "4", "5", "6", "7", "8"]):
        env = dict(os.environ, K1SEG_CHILD="1")
        if nseg != "auto":
            env["MS_K1_NSEG"] = nseg
        subprocess.run([sys.executable, __file__, L, R, frames], env=env, timeout=120)
