"""CloudPipeline (overlapped H2D / build / D2H across successive clouds) must return exactly
what one-at-a-time chatterCallback builds return, in submission order."""
import numpy as np
import pytest

from grid_ndt_b200 import synthetic

pytestmark = pytest.mark.gpu


def test_pipeline_results_equal_serial_builds():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test on a box without CUDA (no fallback exists)")
    from grid_ndt_b200 import TwoDmap
    from grid_ndt_b200.pipeline import CloudPipeline
    clouds = [synthetic.cfg1(300_000), synthetic.cfg2(500_000, scale=0.25), synthetic.cfg3(400_000),
              synthetic.cfg1(50_000), synthetic.cfg2(800_000, scale=0.3)]
    pinned = [torch.from_numpy(c).pin_memory() for c in clouds]
    want = []
    for c in clouds:
        m = TwoDmap(0.2, 0.1)
        m.setInterval(0.08)
        m.chatterCallback(c, "slope")
        want.append((m.voxels.copy(), m.slopes.copy(), m.columns.copy()))
        m.close()
    pipe = CloudPipeline(0.2, 0.1, 0.08, "slope", depth=2)
    got = []
    pipe.submit(pinned[0])
    for i in range(len(clouds)):
        if i + 1 < len(clouds):
            pipe.submit(pinned[i + 1])
        r = pipe.collect()
        got.append((r["voxels"].copy(), r["slopes"].copy(), r["columns"].copy()))
    for i, (w, g) in enumerate(zip(want, got)):
        for name, a, b in zip(("voxels", "slopes", "columns"), w, g):
            assert a.tobytes() == b.tobytes(), f"cloud {i}: {name} differ between pipelined and serial build"
    with pytest.raises(RuntimeError):
        pipe.collect()
    pipe.submit(pinned[0]); pipe.submit(pinned[1])
    with pytest.raises(RuntimeError):
        pipe.submit(pinned[2])
    pipe.collect(); pipe.collect()
    pipe.close()
