"""One full-size oracle comparison per BASELINE config: the GPU build of the WHOLE
cloud against both oracle modes with the complete parity bar of tests/parity.py (integer fields bit-exact,
means / scatters / eigenvalues within 1e-5, labels and reach bits identical or threshold-adjacent).
The oracle needs minutes at these sizes, so this is run once per round through gpurun and its JSON kept
under profiles/.   usage: full_size_parity.py cfg2|cfg3|cfg5 [points] [grid_len]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from grid_ndt_b200 import synthetic
from grid_ndt_b200._abi import default_params
from tests import parity

name = sys.argv[1]
spec = synthetic.CONFIGS[name]
n = int(sys.argv[2]) if len(sys.argv) > 2 else spec.n
gl = float(sys.argv[3]) if len(sys.argv) > 3 else (0.1 if name == "cfg5" else spec.grid_len)
kw = {}
if name == "cfg2" and n != spec.n: kw["scale"] = (n / spec.n) ** 0.5
if name in ("cfg3", "cfg5") and n != spec.n: kw["extent"] = (224.0 if name == "cfg3" else 500.0) * (n / spec.n) ** 0.5
t0 = time.time()
cloud = synthetic.make(name, n, **kw)
t1 = time.time()
rep = parity.run_case(cloud, default_params(gl, spec.z_len, spec.slope_interval), "slope")
t2 = time.time()
out = {"config": name, "points": n, "grid_len": gl, "z_len": spec.z_len, "ok": rep["ok"], "fail": rep["fail"], "counts": rep["counts"],
       "seconds": {"generate": round(t1 - t0, 1), "gpu_build_plus_two_oracle_runs_plus_compare": round(t2 - t1, 1)}}
for k, v in rep.items():
    if k.startswith(("exact_", "col_", "mean_", "scatter_", "evals_", "rough_", "normal_", "label_", "reach_")):
        out[k] = v
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/full_parity_{name}_{n}.json", "w"), indent=1, default=str)
print(json.dumps(out, default=str))
sys.exit(0 if rep["ok"] else 1)
