"""The x-strip protocol (plan -> per-strip build -> thin halo -> gather -> index offsets) on ONE
GPU: two handles play two ranks and the exchange is a device copy, so the multi-GPU data path
is covered by the plain `-m gpu` run; tests/multi_gpu_worker.py runs the same thing over NCCL."""
import ctypes as C

import numpy as np
import pytest

from grid_ndt_b200 import _abi, synthetic
from grid_ndt_b200._abi import VOXEL_DTYPE, default_params

pytestmark = pytest.mark.gpu
REC = VOXEL_DTYPE.itemsize


@pytest.mark.parametrize("n_strips", [2, 3])
def test_strips_with_halo_equal_untiled_build(n_strips):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test on a box without CUDA (no fallback exists)")
    from grid_ndt_b200 import TwoDmap, lib
    from grid_ndt_b200.builder import _check
    from oracle import oracle as O
    L = lib()
    cloud = synthetic.cfg2(700_000, scale=0.28)
    origin = [float(v) for v in cloud[0, :3]]
    dev_cloud = torch.from_numpy(cloud).cuda()
    maps = [TwoDmap(0.2, 0.1) for _ in range(n_strips)]
    for m in maps:
        m.setInterval(0.08)
        m.setCloudFirst(origin)
    cuts = maps[0].plan_tiles(dev_cloud, n_strips)
    assert cuts[0] == -32768 and cuts[-1] == 32768 and np.all(np.diff(cuts) > 0)
    cap = 8192
    halo = torch.zeros(n_strips, 2, (cap + 1) * REC, dtype=torch.uint8, device="cuda")
    for r, m in enumerate(maps):
        m.setTile(int(cuts[r]), int(cuts[r + 1]))
        m.uniformDivision(dev_cloud)
        m.create2DMap("slope")
        _check(m._h, L.gndt_halo_pack(m._h, halo[r, 0].data_ptr(), halo[r, 1].data_ptr(), cap, 0))
    for r, m in enumerate(maps):  # rank r gets the LAST row of r-1 and the FIRST row of r+1
        prev = halo[r - 1, 1].data_ptr() if r > 0 else None
        nxt = halo[r + 1, 0].data_ptr() if r + 1 < n_strips else None
        _check(m._h, L.gndt_halo_edges(m._h, prev, nxt, 0))
    counts = [m.counts() for m in maps]
    sizes = np.array([c["n_voxels"] for c in counts])
    assert sizes.min() > 0.6 * sizes.mean(), f"strips unbalanced: {sizes}"
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    col_off = np.concatenate([[0], np.cumsum([c["n_columns"] for c in counts])])[:-1].astype(np.uint32)
    slope_off = np.concatenate([[0], np.cumsum([c["n_slopes"] for c in counts])])[:-1].astype(np.uint32)
    table = torch.empty(int(offsets[-1]) * REC, dtype=torch.uint8, device="cuda")
    for r, m in enumerate(maps):
        ptr, n = m.device_voxels()
        assert n == sizes[r]
        _check(m._h, L.gndt_copy_voxels(m._h, table.data_ptr() + int(offsets[r]) * REC, n, _abi.GNDT_MEM_DEVICE, None))
    off = (C.c_uint64 * (n_strips + 1))(*[int(x) for x in offsets])
    co = (C.c_uint32 * n_strips)(*[int(x) for x in col_off])
    so = (C.c_uint32 * n_strips)(*[int(x) for x in slope_off])
    _check(maps[0]._h, L.gndt_apply_strip_offsets(maps[0]._h, table.data_ptr(), off, co, so, n_strips, 0))
    torch.cuda.synchronize()
    got = table.cpu().numpy().view(VOXEL_DTYPE)
    p = default_params(0.2, 0.1, 0.08, origin=origin, origin_is_first_point=0)
    o = O.oracle_build(cloud, p)
    assert len(got) == o.counts["n_voxels"]
    for f in ("sx", "sy", "sz", "count", "first_index", "column", "slope"):
        assert np.array_equal(got[f], o.voxels[f]), f
    assert np.array_equal(got["flags"] & 0x10F, o.voxels["flags"] & 0x10F)
    reach_bad = int((((got["flags"] ^ o.voxels["flags"]) & _abi.F_REACH_ALL) != 0).sum())
    assert reach_bad == 0, f"{reach_bad} reach-bit mismatches against the untiled build"
    # without the halo step the strip-boundary rows would miss their cross-strip bits
    boundary = np.isin(np.where(got["sx"] > 0, got["sx"] - 1, got["sx"]), np.concatenate([cuts[1:-1], cuts[1:-1] - 1]))
    assert int(((got["flags"] & (_abi.F_REACH_F | _abi.F_REACH_B))[boundary] != 0).sum()) > 0
    for m in maps:
        m.close()


def test_halo_overflow_is_reported():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test on a box without CUDA (no fallback exists)")
    from grid_ndt_b200 import GndtError, TwoDmap, lib
    from grid_ndt_b200.builder import _check
    L = lib()
    cloud = synthetic.cfg1(100_000)
    m = TwoDmap(0.2, 0.1)
    m.chatterCallback(cloud, "slope")
    cap = 4  # far too small for an x row
    bufs = torch.zeros(2, (cap + 1) * REC, dtype=torch.uint8, device="cuda")
    _check(m._h, L.gndt_halo_pack(m._h, bufs[0].data_ptr(), bufs[1].data_ptr(), cap, 0))
    _check(m._h, L.gndt_halo_edges(m._h, bufs[1].data_ptr(), bufs[0].data_ptr(), 0))
    with pytest.raises(GndtError) as e:
        m.counts()
    assert e.value.status == _abi.GNDT_ERR_CAPACITY
    m.close()
