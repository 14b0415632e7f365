"""Streaming fusion (gndt_update / TwoDmap.change2DMap) — BASELINE configs[3].

The reference's incremental path is dead code with an inconsistent formula, so
the contract is batch equivalence: a map built from cloud A and updated with scans B1..Bk
equals one build over the concatenation A+B1+..+Bk, with the same parity bar as a build."""
import numpy as np
import pytest

from grid_ndt_b200 import _abi, synthetic
from grid_ndt_b200._abi import default_params

pytestmark = pytest.mark.gpu


def _gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test on a box without CUDA (no fallback exists)")
    return torch


def _check_against_batch(m, cloud, p):
    from oracle import oracle as O
    from tests import parity
    o32 = O.oracle_build(cloud, p, "faithful32")
    o64 = O.oracle_build(cloud, p, "truth64")
    rep = parity.compare(m.voxels, m.columns, m.slopes, m.counts(), o32, o64, p)
    assert rep["ok"], rep["fail"]
    return rep


@pytest.mark.parametrize("demand", ["slope", "true"])
def test_update_equals_batch_build(demand):
    _gpu()
    from grid_ndt_b200 import TwoDmap
    cloud = synthetic.cfg2(900_000, scale=0.3)
    cuts = [0, 400_000, 520_000, 521_000, 760_000, 900_000]
    m = TwoDmap(0.2, 0.1)
    m.setInterval(0.08)
    m.chatterCallback(cloud[: cuts[1]], demand)
    for a, b in zip(cuts[1:-1], cuts[2:]):
        m.change2DMap(cloud[a:b])
    p = default_params(0.2, 0.1, 0.08, demand)
    rep = _check_against_batch(m, cloud, p)
    assert rep["counts"] if "counts" in rep else True
    assert m.counts()["n_input"] == 900_000 and m.counts()["n_binned"] == 899_999
    m.close()


def test_update_grows_capacity_and_accepts_device_scans():
    torch = _gpu()
    from grid_ndt_b200 import TwoDmap
    cloud = synthetic.cfg3(600_000, extent=40.0)
    m = TwoDmap(0.1, 0.1)
    m.setInterval(0.08)
    m.chatterCallback(cloud[:20_000], "slope")  # tiny resident map, then 29x more voxels arrive
    v0 = m.counts()["n_voxels"]
    m.change2DMap(torch.from_numpy(cloud[20_000:300_000]).cuda())
    m.change2DMap(cloud[300_000:])
    assert m.counts()["n_voxels"] > 5 * v0
    _check_against_batch(m, cloud, default_params(0.1, 0.1, 0.08))
    m.close()


def test_scans_from_cfg4_generator():
    """BASELINE configs[3] shape: lidar discs fused into the cfg2 scene at high rate."""
    _gpu()
    from grid_ndt_b200 import TwoDmap
    base = synthetic.cfg2(400_000, scale=0.2)
    scans = list(synthetic.scans(6, 20_000, radius=4.0, cfg2_scale=0.2))
    m = TwoDmap(0.2, 0.1)
    m.setInterval(0.08)
    m.chatterCallback(base, "slope")
    for s in scans:
        m.change2DMap(s)
        ms = m.stage_ms()
        assert ms["total"] < 50.0
    _check_against_batch(m, np.concatenate([base] + scans), default_params(0.2, 0.1, 0.08))
    m.close()


def test_update_requires_a_resident_map():
    _gpu()
    from grid_ndt_b200 import GndtError, TwoDmap
    m = TwoDmap(0.2, 0.1)
    with pytest.raises(GndtError) as e:
        m.change2DMap(np.zeros((10, 4), np.float32))  # map2D.h:679: change2DMap on an empty map fails
    assert e.value.status == _abi.GNDT_ERR_STATE
    m.close()


def test_cell_center_and_origin():
    _gpu()
    from grid_ndt_b200 import TwoDmap
    cloud = synthetic.cfg1(50_000)
    m = TwoDmap(0.2, 0.1)
    m.chatterCallback(cloud, "slope")
    assert np.array_equal(m.origin(), cloud[0, :3])
    v = m.voxels
    fit = np.nonzero(v["flags"] & 1)[0][:2000]
    for i in fit[::97]:
        c = m.countPositionXYZ(int(v["sx"][i]), int(v["sy"][i]), int(v["sz"][i]))
        assert np.all(np.abs(c - v["mean"][i]) <= np.array([0.1, 0.1, 0.05]) + 1e-4)  # mean lies inside its cell
        key, z = m.transMortonXYZ(c)
        assert z == int(v["sz"][i])
    m.close()
