"""All BASELINE.json configs on one GPU: device ms, points/s, algorithmic GB/s and roofline
fraction per config; streaming scans/s with p50/p99; the cell-size sweep with the skew
variant.  Writes profiles/configs_<tag>.json and prints a markdown table.

  python tools/bench_configs.py [--tag r1] [--big]   (--big: cfg3 at 50 M, cfg5 at 100 M)
"""
import argparse, json, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from grid_ndt_b200 import TwoDmap, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--tag", default="r1")
ap.add_argument("--big", action="store_true")
ap.add_argument("--builds", type=int, default=6)
a = ap.parse_args()
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
rows = []


def run(name, cloud, gl, zl, builds=a.builds):
    dev = torch.from_numpy(cloud).cuda()
    m = TwoDmap(gl, zl); m.setInterval(0.08); m.stage_timing(True)
    best = None
    for _ in range(builds):
        m.chatterCallback(dev, "slope"); torch.cuda.synchronize()
        s = m.stage_ms()
        best = s if best is None or s["total"] < best["total"] else best
    c = m.counts()
    n = cloud.shape[0]
    b_alg = 16.0 * n + 96.0 * c["n_voxels"]
    row = {"config": name, "points": n, "grid_len": gl, "z_len": zl, "device_ms": best["total"], "stage_ms": best,
           "mpts_per_s": n / best["total"] / 1e3, "voxels": c["n_voxels"], "columns": c["n_columns"], "slopes": c["n_slopes"],
           "b_alg_gb": b_alg / 1e9, "achieved_gbs": b_alg / best["total"] / 1e6, "frac_of_measured_peak": b_alg / best["total"] / 1e6 / PEAK,
           "frac_of_8tbs": b_alg / best["total"] / 1e6 / 8000.0}
    rows.append(row)
    print(json.dumps(row), flush=True)
    m.close(); del dev
    torch.cuda.empty_cache()
    return row


run("cfg1 ramp+step", synthetic.cfg1(1_000_000), 0.2, 0.1)
cfg2 = synthetic.cfg2(10_000_000)
run("cfg2 multi-level", cfg2, 0.2, 0.1)
n3 = 50_000_000 if a.big else 20_000_000
run(f"cfg3 terrain {n3 // 1_000_000}M", synthetic.cfg3(n3, extent=224.0 * (n3 / 50e6) ** 0.5), 0.1, 0.1, builds=4)

# cfg4: streaming fusion of 100k-point scans into the resident cfg2 map
m = TwoDmap(0.2, 0.1); m.setInterval(0.08); m.stage_timing(True)
m.chatterCallback(torch.from_numpy(cfg2).cuda(), "slope"); torch.cuda.synchronize()
n_scans = 200
lat_dev, lat_wall = [], []
scans = [torch.from_numpy(s).cuda() for s in synthetic.scans(n_scans, 100_000)]
for s in scans:
    t0 = time.perf_counter()
    m.change2DMap(s); torch.cuda.synchronize()
    lat_wall.append((time.perf_counter() - t0) * 1e3)
    lat_dev.append(m.stage_ms()["total"])
c = m.counts()
row4 = {"config": "cfg4 streaming", "scan_points": 100_000, "n_scans": n_scans, "resident_voxels": c["n_voxels"],
        "device_ms_p50": statistics.median(lat_dev), "device_ms_p99": sorted(lat_dev)[int(0.99 * n_scans) - 1],
        "wall_ms_p50": statistics.median(lat_wall), "wall_ms_p99": sorted(lat_wall)[int(0.99 * n_scans) - 1],
        "scans_per_s_wall": 1e3 / statistics.mean(lat_wall), "points_total": c["n_input"]}
rows.append(row4)
print(json.dumps(row4), flush=True)
m.close(); del scans
torch.cuda.empty_cache()

# cfg5: cell-size sweep (+ skew) on the 500 m x 500 m terrain
n5 = 100_000_000 if a.big else 20_000_000
ext5 = 500.0 * (n5 / 100e6) ** 0.5
base5 = synthetic.cfg5(n5, extent=ext5)
for gl in (0.05, 0.1, 0.2, 0.5, 1.0):
    run(f"cfg5 sweep {n5 // 1_000_000}M cell {gl}", base5, gl, 0.1, builds=3)
del base5
run(f"cfg5 skew {n5 // 1_000_000}M cell 0.2", synthetic.cfg5(n5, extent=ext5, skew=True), 0.2, 0.1, builds=3)

os.makedirs("profiles", exist_ok=True)
json.dump({"tag": a.tag, "peak_gbs": PEAK, "rows": rows}, open(f"profiles/configs_{a.tag}.json", "w"), indent=1)
print("\n| config | N | cell (m) | device ms | Mpts/s | B_alg (GB) | achieved GB/s | % of measured peak | % of 8 TB/s |")
print("|---|---|---|---|---|---|---|---|---|")
for r in rows:
    if "device_ms" in r:
        print(f"| {r['config']} | {r['points'] / 1e6:.0f} M | {r['grid_len']} | {r['device_ms']:.3f} | {r['mpts_per_s']:.0f} | {r['b_alg_gb']:.3f} | {r['achieved_gbs']:.0f} | {100 * r['frac_of_measured_peak']:.2f} | {100 * r['frac_of_8tbs']:.2f} |")
    else:
        print(f"| {r['config']} | 100 k/scan x {r['n_scans']} | 0.2 | p50 {r['device_ms_p50']:.3f} / p99 {r['device_ms_p99']:.3f} (wall p50 {r['wall_ms_p50']:.3f}) | {r['scans_per_s_wall']:.0f} scans/s | | | | |")
