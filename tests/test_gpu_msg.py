"""gndt_build_msg: the build straight from a sensor_msgs/PointCloud2 payload (SURVEY §8(f)1;
src/receiver.cpp:137-143, src/publisher.cpp:55) gives byte-identical tables to the build from the
unpacked cloud, for the usual layout (read in place), for scattered / unaligned field offsets and
for big-endian messages (repacked on the device), from pageable memory (pinned staging ring) and
from pinned memory."""
import numpy as np
import pytest

from grid_ndt_b200 import synthetic

pytestmark = pytest.mark.gpu


def test_pointcloud2_layouts_agree():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test on a box without CUDA (no fallback exists)")
    from grid_ndt_b200 import TwoDmap
    cloud = synthetic.cfg2(2_500_000, scale=0.5)  # 40 MB: several chunks of the staging ring
    n = len(cloud)
    m = TwoDmap(0.2, 0.1)
    m.setInterval(0.08)
    m.chatterCallback(cloud, "slope")
    want = (m.voxels.tobytes(), m.slopes.tobytes(), m.columns.tobytes())

    def build(data, step, offs, big=False):
        m.chatterCallbackMsg(data, n, 1, step, offs, big, "slope")
        return (m.voxels.tobytes(), m.slopes.tobytes(), m.columns.tobytes())

    # pcl::toROSMsg of PointXYZ: point_step 16, x y z at 0 4 8 (read in place)
    assert build(cloud.tobytes(), 16, (0, 4, 8)) == want
    assert build(torch.from_numpy(cloud.view(np.uint8).reshape(-1)).pin_memory(), 16, (0, 4, 8)) == want
    # a 32-byte record with the fields scattered and one of them on an odd boundary
    rec = np.zeros((n, 35), np.uint8)
    xyz = cloud[:, :3].copy()
    rec[:, 4:8] = xyz[:, 0:1].view(np.uint8)
    rec[:, 13:17] = xyz[:, 1:2].view(np.uint8)
    rec[:, 24:28] = xyz[:, 2:3].view(np.uint8)
    assert build(rec.tobytes(), 35, (4, 13, 24)) == want
    # big-endian floats
    be = cloud[:, :3].astype(">f4")
    assert build(be.tobytes(), 12, (0, 4, 8), big=True) == want
    assert m.counts()["n_input"] == n
    m.close()
