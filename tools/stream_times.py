"""cfg4 (streaming): 100 k-point scans fused into a resident map, at two resident sizes (cfg2: 0.55 M voxels,
cfg3: 6.5 M voxels).  Reports p50 / p99 device ms per gndt_update, wall ms, scans/s, cells changed per scan,
and the same for gndt_remove.  Writes gpurun_out/stream_<tag>.json."""
import json, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from grid_ndt_b200 import TwoDmap, synthetic

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
n_scans = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rows = []


def discs(cloud, n_scans, n_per, radius, rng):
    """Scans = the points of an independent cloud of the same scene inside a disc around a moving pose."""
    lo, hi = cloud[1:, :2].min(axis=0), cloud[1:, :2].max(axis=0)
    out = []
    pose = lo + 0.3 * (hi - lo)
    step = (hi - lo) * 0.4 / n_scans
    for _ in range(n_scans):
        pose = pose + step
        d2 = ((cloud[:, :2] - pose) ** 2).sum(axis=1)
        idx = np.nonzero(d2 < radius * radius)[0]
        idx = idx[rng.permutation(len(idx))[:n_per]]
        out.append(np.ascontiguousarray(cloud[np.sort(idx)]))
    return out


def run(name, base, scans, gl, zl):
    m = TwoDmap(gl, zl); m.setInterval(0.08)
    m.chatterCallback(torch.from_numpy(base).cuda(), "slope"); torch.cuda.synchronize()
    v0 = m.counts()["n_voxels"]
    dev = [torch.from_numpy(s).cuda() for s in scans]
    lat_dev, lat_wall, changed, new_vox = [], [], [], []
    for s in dev:
        before = m.counts()["n_voxels"]
        t0 = time.perf_counter()
        m.change2DMap(s); torch.cuda.synchronize()
        lat_wall.append((time.perf_counter() - t0) * 1e3)
        lat_dev.append(m.stage_ms()["total"])
        changed.append(len(m.changed_columns))
        new_vox.append(m.counts()["n_voxels"] - before)
    rem_dev = []
    for s in reversed(dev[-10:]):
        m.del2DMap(s); torch.cuda.synchronize()
        m.counts()
        rem_dev.append(m.stage_ms()["total"])
    pct = lambda a, q: sorted(a)[min(len(a) - 1, int(q * len(a)))]
    row = {"resident": name, "resident_points": int(base.shape[0]), "resident_voxels": v0, "voxels_after": m.counts()["n_voxels"],
           "scan_points": int(scans[0].shape[0]), "n_scans": len(scans),
           "update_device_ms_p50": statistics.median(lat_dev), "update_device_ms_p99": pct(lat_dev, 0.99),
           "update_wall_ms_p50": statistics.median(lat_wall), "scans_per_s_wall": 1e3 / statistics.mean(lat_wall),
           "cells_changed_per_scan_median": statistics.median(changed), "new_voxels_per_scan_median": statistics.median(new_vox),
           "remove_device_ms_p50": statistics.median(rem_dev)}
    rows.append(row)
    print(json.dumps(row), flush=True)
    m.close(); del dev
    torch.cuda.empty_cache()


rng = np.random.default_rng(11)
cfg2 = synthetic.cfg2(10_000_000)
run("cfg2 10M pts, 0.2 m", cfg2, list(synthetic.scans(n_scans, 100_000)), 0.2, 0.1)
del cfg2
cfg3 = synthetic.cfg3(50_000_000)
other = synthetic.cfg3(20_000_000, cfg=13)
run("cfg3 50M pts, 0.1 m", cfg3, discs(other, n_scans, 100_000, 20.0, rng), 0.1, 0.1)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"tag": tag, "rows": rows}, open(f"gpurun_out/stream_{tag}.json", "w"), indent=1)
