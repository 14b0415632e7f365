"""Probe: do copy-engine peer copies running beside the builds slow them down?
One process, GPUs 0 and 1 (plain peer access).  GPU 0 builds cfg2 10 M back to back; after every build a side
stream sends REPS x 32.8 MB (a strip's slopes + columns) to GPU 1 with cudaMemcpyPeerAsync (torch copy_), and,
with BIDIR=1, GPU 1 sends the same amount back.  Prints the mean device time of a build."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from grid_ndt_b200 import TwoDmap

STRIP = 32_800_000
reps = int(os.environ.get("REPS", "7")); mode = os.environ.get("MODE", "ce"); bidir = os.environ.get("BIDIR", "0") == "1"
d0, d1 = torch.device("cuda", 0), torch.device("cuda", 1)
torch.cuda.set_device(0)
cloud = torch.from_numpy(bench.make_cloud(0)).to(d0)
origin = [float(np.float32(0.5 * bench.SCENE_W + 0.013)), float(np.float32(40.007)), 1.0]
m = TwoDmap(0.2, 0.1); m.setInterval(0.08); m.setCloudFirst(origin)
src0 = torch.empty(STRIP, dtype=torch.uint8, device=d0); dst1 = torch.empty(STRIP * 8, dtype=torch.uint8, device=d1)
src1 = torch.empty(STRIP, dtype=torch.uint8, device=d1); dst0 = torch.empty(STRIP * 8, dtype=torch.uint8, device=d0)
bs, side = torch.cuda.Stream(d0), torch.cuda.Stream(d0)
back = torch.cuda.Stream(d1)
def run(k):
    ev = []
    for i in range(k):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(bs):
            a.record(bs); m.uniformDivision(cloud); m.create2DMap("slope", stream=bs.cuda_stream); b.record(bs)
        ev.append((a, b))
        if mode == "ce":
            side.wait_event(b)
            with torch.cuda.stream(side):
                for r in range(reps): dst1[r * STRIP:(r + 1) * STRIP].copy_(src0, non_blocking=True)
            if bidir:
                back.wait_event(b)
                with torch.cuda.stream(back):
                    for r in range(reps): dst0[r * STRIP:(r + 1) * STRIP].copy_(src1, non_blocking=True)
        if i >= 3: ev[i - 3][1].synchronize()  # stay three builds ahead at most
    torch.cuda.synchronize(d0); torch.cuda.synchronize(d1)
    return ev
run(6)
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(bs); ev = run(30); t1.record(bs); torch.cuda.synchronize(d0)
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(side):  # the copies alone
    s0.record(side)
    for r in range(reps): dst1[r * STRIP:(r + 1) * STRIP].copy_(src0, non_blocking=True)
    s1.record(side)
torch.cuda.synchronize(d0)
print(json.dumps({"mode": mode, "reps": reps, "bidir": bidir, "ms_per_step": round(t0.elapsed_time(t1) / 30, 4),
                  "build_ms": round(float(np.mean([a.elapsed_time(b) for a, b in ev])), 4),
                  "copies_alone_ms": round(s0.elapsed_time(s1), 4), "copy_GBs": round(reps * STRIP / s0.elapsed_time(s1) / 1e6, 1)}), flush=True)
