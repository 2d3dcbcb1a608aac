"""Experiment: per-phase clock64 breakdown of the onesweep kernel.

    make -C cloudbrush_b200/csrc timing
    CB_LIB=cloudbrush_b200/libcloudbrush_gpu_timing.so python tools/os_timing.py

Results of round 1: profiles/r01_onesweep_phases.md."""
import ctypes as C, sys, os
sys.path.insert(0, ".")
import numpy as np, torch
from cloudbrush_b200 import BrushConfig, BrushGpu, synth, _lib
sfa, m = synth.make_config("cfg2")
dev = torch.from_numpy(sfa).cuda()
g = BrushGpu(BrushConfig(k=m["k"], readlen=m["readlen"]))
L = _lib.load()
out = (C.c_ulonglong * 16)()
names = ["tile id+zero+load issue", "wait key loads", "ranking", "barrier after ranking", "digit scan+publish", "block scan",
         "staging keys+vals", "look-back(d0)", "barrier after look-back", "write-out"]
for it in range(3):
    g.borrow_sfa_device(dev.data_ptr(), dev.numel()); g.preprocess()
    L.cb_debug_os_prof(out, 1)
    g.build_overlap()
    L.cb_debug_os_prof(out, 1)
    v = list(out); tiles = v[14]; tot = sum(v[:10])
    print(f"--- run {it}: tiles {tiles}, cycles/tile {tot/tiles:.0f}")
    print(f"  look-back per tile: round trips {v[11]/tiles:.2f}, of which spins {v[12]/tiles:.2f}, depth {v[13]/tiles:.2f}")
    for i, nm in enumerate(names):
        print(f"  {nm:28s} {v[i]/tiles:8.0f}  {100*v[i]/tot:5.1f}%")
