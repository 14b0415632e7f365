"""Profiling driver: N map builds of one synthetic cloud on cuda:0 (run under ncu)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from grid_ndt_b200 import TwoDmap, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="cfg2")
ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--builds", type=int, default=2)
ap.add_argument("--grid", type=float, default=None)
a = ap.parse_args()
spec = synthetic.CONFIGS[a.cfg]
kw = {}
if a.cfg == "cfg2" and a.n != spec.n:
    kw["scale"] = (a.n / spec.n) ** 0.5
if a.cfg in ("cfg3", "cfg5") and a.n != spec.n:
    kw["extent"] = (224.0 if a.cfg == "cfg3" else 500.0) * (a.n / spec.n) ** 0.5
cloud = torch.from_numpy(synthetic.make(a.cfg, a.n, **kw)).cuda()
m = TwoDmap(a.grid or spec.grid_len, spec.z_len)
m.setInterval(spec.slope_interval)
m.stage_timing(True)
for _ in range(a.builds):
    m.chatterCallback(cloud, "slope")
    torch.cuda.synchronize()
print(m.counts(), m.stage_ms())
