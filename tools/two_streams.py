"""Throughput of back-to-back builds with 1, 2 and 3 builds in flight on one GPU (device-resident
input, separate handles and streams): does another build fill the tails and the small kernels?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from grid_ndt_b200 import TwoDmap, synthetic
cloud = torch.from_numpy(synthetic.cfg2(10_000_000)).cuda()
K = 24
for depth in (1, 2, 3):
    maps = [TwoDmap(0.2, 0.1) for _ in range(depth)]
    for m in maps: m.setInterval(0.08)
    streams = [torch.cuda.Stream() for _ in range(depth)]
    def run(k):
        for i in range(k):
            m, s = maps[i % depth], streams[i % depth]
            m._p.origin_is_first_point = 1
            m.uniformDivision(cloud)
            m.create2DMap("slope", stream=s.cuda_stream)
    run(2 * depth); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams: s.wait_stream(torch.cuda.current_stream())
    run(K)
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"builds in flight {depth}: {ms:.4f} ms per build, {10e6 / ms / 1e6:.2f} G points/s", flush=True)
    for m in maps: m.close()
