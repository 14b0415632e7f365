"""Generate the committed golden vectors from the REFERENCE ITSELF as compiled here
(oracle/_ref/libgndt_ref.so = /root/reference/src/receiver.cpp + include/*.h against inert
shims).  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Outputs (small, committed):
  kat_keys.json            countMorton / mortonToXY / transMortonXYZ known answers
  bridge_ground.json       counters + digests of the genePcd.cpp fixture build
  cfg1_50k.json            counters + digests of a 50k-point synthetic build (both demands)
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from grid_ndt_b200 import synthetic  # noqa: E402
from grid_ndt_b200._abi import default_params  # noqa: E402
from oracle import oracle as O  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:32]


def table_digests(m):
    v = m.voxels
    out = {f: digest(v[f]) for f in ("sx", "sy", "sz", "count", "mean", "scatter", "evals", "normal", "rough")}
    out["flags_fitted_slope_reach"] = digest(v["flags"] & 0xF3)  # `up`/`down` of non-Slopes are not observable
    out["slope_down"] = digest((v["flags"] & 0x08)[(v["flags"] & 0x02) != 0])
    out["morton_list"] = digest(m.morton_list)
    out["first_seen_rank_order"] = digest(np.argsort(v["first_index"], kind="stable").astype(np.uint32))
    return out


def main():
    rng = np.random.default_rng(20260001)
    kat = {"count_morton": [], "trans": []}
    pairs = [(5, 7), (1, 1), (1, 2), (2, 1), (3, 3), (100, 37), (32767, 32767), (32768, 1), (65536, 1), (1, 32768)]
    pairs += [tuple(int(x) for x in rng.integers(1, 32768, 2)) for _ in range(40)]
    for a, b in pairs:
        s = O.ref_count_morton(a, b)
        kat["count_morton"].append({"a": a, "b": b, "string": s, "inverse": list(O.ref_morton_to_xy(int(s))) if abs(int(s)) < 2**31 else None})
    cases = [((1, 1, 1), 0.1, 0.05, (1.05, 1.2, 1.3)), ((1, 1, 1), 0.2, 0.1, (1.4, 0.8, 1.0)), ((1, 1, 1), 0.1, 0.05, (3.0, 1.6, 1.1)),
             ((0, 0, 0), 0.5, 0.1, (-0.0, 0.0, -1e-9)), ((10.5, -3.25, 2.0), 0.25, 0.1, (10.5, -3.25, 2.0))]
    for _ in range(200):
        o = rng.uniform(-50, 50, 3).astype(np.float32)
        gl, zl = np.float32(rng.choice([0.05, 0.1, 0.2, 0.25, 0.5, 1.0])), np.float32(rng.choice([0.05, 0.1]))
        k = rng.integers(-200, 200, 3)
        pos = (o + np.array([k[0] * gl, k[1] * gl, k[2] * zl], np.float32) + rng.choice([0, 0, 1e-6, -1e-6, 0.013], 3).astype(np.float32)).astype(np.float32)
        cases.append((tuple(float(x) for x in o), float(gl), float(zl), tuple(float(x) for x in pos)))
    for o, gl, zl, pos in cases:
        key, z = O.ref_trans(o, gl, zl, pos)
        kat["trans"].append({"origin": o, "grid_len": gl, "z_len": zl, "pos": pos, "key": key, "z": z})
    json.dump(kat, open(os.path.join(HERE, "kat_keys.json"), "w"), indent=0)

    pts, assigned = O.bridge_ground()
    p = default_params(0.1, 0.05, 0.08, "slope")  # launch/parameters.txt:53-59
    m = O.ref_build(pts, p)
    g = {"assigned_points": assigned, "cloud_sha": digest(pts), "params": [0.1, 0.05, 0.08, "slope"], "counts": m.counts, "digests": table_digests(m)}
    json.dump(g, open(os.path.join(HERE, "bridge_ground.json"), "w"), indent=1)

    cloud = synthetic.cfg1(50_000)
    out = {"cloud_sha": digest(cloud)}
    for demand in ("slope", "true"):
        m = O.ref_build(cloud, default_params(0.2, 0.1, 0.08, demand))
        out[demand] = {"counts": m.counts, "digests": table_digests(m)}
    json.dump(out, open(os.path.join(HERE, "cfg1_50k.json"), "w"), indent=1)
    print("golden vectors written")


if __name__ == "__main__":
    main()
