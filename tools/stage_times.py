"""Stage timings (min over builds) of one synthetic config; used to compare tuning variants."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from grid_ndt_b200 import TwoDmap, synthetic
ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="cfg2"); ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--builds", type=int, default=8); ap.add_argument("--grid", type=float, default=None)
a = ap.parse_args()
spec = synthetic.CONFIGS[a.cfg]
kw = {}
if a.cfg == "cfg2" and a.n != spec.n: kw["scale"] = (a.n / spec.n) ** 0.5
if a.cfg in ("cfg3", "cfg5") and a.n != spec.n: kw["extent"] = (224.0 if a.cfg == "cfg3" else 500.0) * (a.n / spec.n) ** 0.5
cloud = torch.from_numpy(synthetic.make(a.cfg, a.n, **kw)).cuda()
m = TwoDmap(a.grid or spec.grid_len, spec.z_len); m.setInterval(spec.slope_interval); m.stage_timing(True); m.stage_timing(True)
best = None
for _ in range(a.builds):
    m.chatterCallback(cloud, "slope"); torch.cuda.synchronize()
    s = m.stage_ms()
    best = s if best is None else {k: min(best[k], s[k]) for k in s}
c = m.counts()
print(os.environ.get("GNDT_LIB", "default"), a.cfg, a.n, {k: round(v, 4) for k, v in best.items()}, c["n_voxels"], c["n_slopes"], c["n_columns"])
