"""Debug only: per-phase clock64 totals of an instrumented throwaway variant library."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from grid_ndt_b200 import TwoDmap, synthetic, lib
cloud = torch.from_numpy(synthetic.cfg2(10_000_000)).cuda()
m = TwoDmap(0.2, 0.1); m.setInterval(0.08)
m.chatterCallback(cloud, "slope"); torch.cuda.synchronize()
L = lib(); L.gndt_debug_phases.argtypes = [C.c_void_p, C.c_int]
L.gndt_debug_phases(None, 1)
m.chatterCallback(cloud, "slope"); torch.cuda.synchronize()
out = (C.c_ulonglong * 64)(); L.gndt_debug_phases(out, 0)
names = sys.argv[1].split(",")
tiles = int(sys.argv[2])
tot = sum(out[k] for k in range(len(names)))
print(f"avg CTA lifetime {tot / tiles:8.0f} cyc | " + " | ".join(f"{names[k]} {out[k] / tiles:6.0f}" for k in range(len(names))))
