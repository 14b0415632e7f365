"""gndt_multi_*: ONE process, several GPUs, everything inside libgndt.so (the form the reference's
single receiver process would call; VERDICT r1 missing #5).  Needs >= 2 GPUs on the box."""
import numpy as np
import pytest

from grid_ndt_b200 import synthetic
from grid_ndt_b200._abi import default_params

pytestmark = pytest.mark.gpu


def test_single_process_two_gpus_equal_the_oracle():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test on a box without CUDA (no fallback exists)")
    nd = min(torch.cuda.device_count(), 4)
    if nd < 2:
        pytest.skip("needs 2 GPUs")
    from grid_ndt_b200.multi import MultiTwoDmap
    from oracle import oracle as O
    from tests import parity
    cloud = synthetic.cfg2(1_200_000, scale=0.35)
    mm = MultiTwoDmap(0.2, 0.1, 0.08, list(range(nd)), capacity=2_000_000)
    p = default_params(0.2, 0.1, 0.08)
    for src in (cloud, torch.from_numpy(cloud).cuda(0)):  # host input, then device input on devices[0]
        mm.chatterCallback(src)
        t0 = mm.tables(0)
        t1 = mm.tables(nd - 1)
        for k in ("voxels", "slopes", "columns"):
            assert t0[k].tobytes() == t1[k].tobytes(), f"GPU 0 and GPU {nd - 1} hold different {k}"
        o32, o64 = O.oracle_build(cloud, p, "faithful32"), O.oracle_build(cloud, p, "truth64")
        rep = parity.compare_gathered(t0["voxels"], t0["columns"], t0["slopes"], o32, o64, p)
        assert rep["ok"], rep["fail"]
        assert min(t0["strip_voxels"]) > 0 and np.all(np.diff(mm.cuts()) > 0)
    # streaming through the same interface
    a, b = cloud[:800_000], cloud[800_000:]
    mm.chatterCallback(a)
    mm.change2DMap(b)
    t = mm.tables(0)
    rep = parity.compare_gathered(t["voxels"], t["columns"], t["slopes"], o32, o64, p)
    assert rep["ok"], rep["fail"]
    mm.close()
