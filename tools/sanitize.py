"""compute-sanitizer driver: a 100 k-point cfg1 build, the degenerate fuzz clouds, a 2-strip build on one
GPU, an update + remove, the traversability graph and a PointCloud2 ingest — small enough for the tools' 10-100x slowdown."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from grid_ndt_b200 import TwoDmap, synthetic
m = TwoDmap(0.2, 0.1); m.setInterval(0.08)
c = synthetic.cfg1(100_000)
m.chatterCallback(c, "slope"); print("cfg1 100k", m.counts())
m.edges()
m.change2DMap(synthetic.cfg1(100_000)[50_000:70_000]); print("update", m.counts(), len(m.changed_columns))
m.del2DMap(synthetic.cfg1(100_000)[50_000:70_000]); print("remove", m.counts())
m.chatterCallbackMsg(c.tobytes(), len(c), 1, 16); print("msg", m.counts())
rng = np.random.default_rng(7)
for i in range(12):  # degenerate little clouds: duplicates, lines, planes, far origin
    k = int(rng.integers(2, 400))
    p = rng.normal(0, [3.0, 3.0, 0.3][i % 3], (k, 4)).astype(np.float32)
    if i % 4 == 0: p[:, 2] = 0.5
    if i % 4 == 1: p[:, 0] = p[0, 0]
    if i % 5 == 0: p[k // 2:] = p[0]
    m.chatterCallback(p, "true" if i % 2 else "slope"); m.counts()
# two strips on one GPU (tile filter)
cx = TwoDmap(0.2, 0.1); cx.setInterval(0.08); cx.setCloudFirst(c[0, :3])
cuts = cx.plan_tiles(c, 2)
for lo, hi in zip(cuts[:-1], cuts[1:]):
    cx.setTile(int(lo), int(hi)); cx.uniformDivision(c); cx.create2DMap("slope"); print("strip", cx.counts()["n_voxels"])
torch.cuda.synchronize(); print("sanitize driver done")
