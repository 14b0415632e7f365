"""Host-side logic that needs no GPU: synthetic clouds, the strip all-gather under gloo with
world_size 2, the bench's reference arm."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from grid_ndt_b200 import synthetic
from grid_ndt_b200._abi import VOXEL_DTYPE, default_params
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synthetic_clouds_are_deterministic_and_shaped():
    for fn, kw in ((synthetic.cfg1, {}), (synthetic.cfg2, {"scale": 0.1}), (synthetic.cfg3, {"extent": 10.0})):
        a, b = fn(40_000, **kw), fn(40_000, **kw)
        assert a.shape == (40_000, 4) and a.dtype == np.float32 and np.array_equal(a, b)
        assert np.all(a[:, 3] == 1.0)
        nz = np.abs(a[:, :3]).sum(axis=1) == 0
        assert 0.015 < nz.mean() < 0.025 and nz[-1] and not nz[0]  # 2 % trailing (0,0,0) padding
    sk = synthetic.cfg5(50_000, extent=50.0, skew=True)
    assert sk.shape == (50_000, 4)
    scans = list(synthetic.scans(3, 1000))
    assert len(scans) == 3 and scans[0].shape == (1000, 4)


def test_tile_filter_partitions_the_map():
    """Strips built independently (oracle with tile_lo/hi) concatenate to the untiled map."""
    cloud = synthetic.cfg2(150_000, scale=0.15)
    p = default_params(0.2, 0.1, 0.08)
    p.origin_is_first_point = 0
    for i in range(3):
        p.origin[i] = cloud[0, i]
    full = O.oracle_build(cloud, p)
    cx = np.where(full.voxels["sx"] > 0, full.voxels["sx"] - 1, full.voxels["sx"])
    cut = int(np.median(cx))
    parts = []
    for lo, hi in ((-40000, cut), (cut, 40000)):
        q = default_params(0.2, 0.1, 0.08)
        q.origin_is_first_point = 0
        for i in range(3):
            q.origin[i] = cloud[0, i]
        q.tile_lo, q.tile_hi = lo, hi
        parts.append(O.oracle_build(cloud, q))
    assert parts[0].counts["n_binned"] + parts[1].counts["n_binned"] == full.counts["n_binned"]
    both = np.concatenate([parts[0].voxels, parts[1].voxels])
    for f in ("sx", "sy", "sz", "count", "first_index", "mean", "scatter", "rough"):
        assert np.array_equal(both[f], full.voxels[f]), f
    assert np.array_equal(both["flags"] & 0x0F, full.voxels["flags"] & 0x0F)


_WORKER = r"""
import os, pickle, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from grid_ndt_b200 import synthetic, _abi
from grid_ndt_b200._abi import default_params
from grid_ndt_b200.tiles import allgather_bytes
from oracle import oracle as O
from tests import parity
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
# the one collective of the native exchange: fixed-size buffer handles, in rank order
blob = bytes([rank]) * 128
every = allgather_bytes(blob, world)
assert every == [bytes([r]) * 128 for r in range(world)], every
cloud = synthetic.cfg2(120_000, scale=0.12)
def params(lo, hi):
    p = default_params(0.2, 0.1, 0.08); p.origin_is_first_point = 0
    for i in range(3): p.origin[i] = cloud[0, i]
    p.tile_lo, p.tile_hi = lo, hi
    return p
full = O.oracle_build(cloud, params(0, 0))
cx = np.where(full.voxels["sx"] > 0, full.voxels["sx"] - 1, full.voxels["sx"])
cut = int(np.quantile(cx, 0.4))
for cuts in ([-32768, cut, 32768], [-32768, -32768, 32768]):   # the second: strip 0 is EMPTY
    lo, hi = cuts[rank], cuts[rank + 1]
    if lo >= hi: lo, hi = _abi.TILE_EMPTY                       # equal cuts must keep nothing (lo >= hi = filter off)
    mine = O.oracle_build(cloud, params(lo, hi))
    if cuts[rank] >= cuts[rank + 1]: assert len(mine.voxels) == 0
    got = [None] * world
    dist.all_gather_object(got, (mine.voxels, mine.columns))
    vox, cols = parity.assemble_strips(got)
    assert len(vox) == len(full.voxels) and len(cols) == len(full.columns)
    for f in ("sx", "sy", "sz", "count", "first_index", "mean", "scatter", "evals", "rough", "column", "slope"):
        assert np.array_equal(vox[f], full.voxels[f]), f
    assert np.array_equal(vox["flags"] & 0x10F, full.voxels["flags"] & 0x10F)
    for f in ("sx", "sy", "first_index", "voxel_begin", "voxel_count", "slope_begin", "slope_count"):
        assert np.array_equal(cols[f], full.columns[f]), f
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_strips_world2_gloo(tmp_path):
    """The N>1 host logic with two gloo ranks on CPU: the handle exchange of the native strip
    exchange, the empty-strip rule, and the index fix-up rule of the device-side push (stated on
    the host in tests/parity.assemble_strips); strips come from the oracle, no GPU involved."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    port = 29500 + (os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"rank {r} ok" in o, o


def test_bench_reference_arm_prints_contract_line():
    env = dict(os.environ)
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                                  env=env, timeout=600).decode().strip().splitlines()[-1]
    line = json.loads(out)
    assert line["impl"] == "reference" and line["unit"] == "points/s" and line["value"] > 1e4
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["metric"] == "ndt_map_build_points_per_sec" and line["higher_is_better"] is True


def test_bench_reference_arm_nonzero_rank_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"], env=env, timeout=120)
    assert out.strip() == b""


def test_integer_lookups_on_oracle_tables():
    """gndt_find_column / gndt_find_slope / gndt_neighbor_column (host helpers of libgndt.so, no
    GPU involved) on tables produced by the oracle: every cell is found at its own index, empty
    cells and index 0 give -1, neighbours are one step in contiguous index space with the
    quadrant crossing between -1 and +1 (countLRFB, map2D.h:197-263)."""
    import ctypes as C
    import numpy as np
    from grid_ndt_b200 import lib, synthetic
    from grid_ndt_b200._abi import SLOPE_DTYPE, default_params
    from oracle import oracle as O
    cloud = synthetic.cfg2(120_000, scale=0.12)
    cloud[:, :2] -= cloud[0, :2] * np.float32(0.5)  # all four quadrants around the first point
    cloud[0, :2] = cloud[1:, :2].mean(axis=0)
    o = O.oracle_build(cloud, default_params(0.2, 0.1, 0.08))
    cols, vox = o.columns, o.voxels
    assert len({(int(np.sign(c["sx"])), int(np.sign(c["sy"]))) for c in cols}) == 4, "want all four quadrants"
    idx = np.nonzero(vox["flags"] & 2)[0]
    sl = np.zeros(len(idx), SLOPE_DTYPE)
    for f in ("sx", "sy", "sz"):
        sl[f] = vox[f][idx]
    L = lib()
    table = {(int(c["sx"]), int(c["sy"])): i for i, c in enumerate(cols)}
    cont = lambda s: s - 1 if s > 0 else s
    sgn = lambda c: c + 1 if c >= 0 else c
    step = {0: (0, -1), 1: (0, 1), 2: (1, 0), 3: (-1, 0)}
    rng = np.random.default_rng(5)
    for i in rng.choice(len(cols), 3000, replace=False):
        sx, sy = int(cols[i]["sx"]), int(cols[i]["sy"])
        assert L.gndt_find_column(cols.ctypes.data, len(cols), sx, sy) == i
        for d, (dx, dy) in step.items():
            want = table.get((sgn(cont(sx) + dx), sgn(cont(sy) + dy)), -1)
            assert L.gndt_neighbor_column(cols.ctypes.data, len(cols), sx, sy, d) == want
        for s in range(cols[i]["slope_begin"], cols[i]["slope_begin"] + cols[i]["slope_count"]):
            assert L.gndt_find_slope(cols.ctypes.data, len(cols), sl.ctypes.data, sx, sy, int(sl[s]["sz"])) == s
        assert L.gndt_find_slope(cols.ctypes.data, len(cols), sl.ctypes.data, sx, sy, 32000) == -1
    assert L.gndt_find_column(cols.ctypes.data, len(cols), 0, 5) == -1
    assert L.gndt_find_column(cols.ctypes.data, len(cols), 30000, 30000) == -1
    assert L.gndt_neighbor_column(cols.ctypes.data, len(cols), 1, 1, 7) == -1
    assert L.gndt_find_column(None, 0, 1, 1) == -1
