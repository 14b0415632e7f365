"""Back-to-back map builds from host clouds with the PCIe transfers overlapped.

The reference handles one `PointCloud2` message per callback (src/receiver.cpp:132-162): copy
the message into a cloud, bin it, fit it, hand the containers to the planner.  Through the C
ABI one such build is  H2D(cloud) -> kernels -> D2H(records),  and on a B200 the two copies are
five times longer than the kernels.  `CloudPipeline` keeps `depth` builders (each a TwoDmap
with its own handle, stream, device buffers and pinned result buffers) and rotates through
them, so that the upload of cloud i+1 runs while cloud i is being built and read back — PCIe
is full duplex and the copy engines run beside the SMs.  Every build is still a complete,
independent `chatterCallback`; results are byte-identical to the unpipelined call
(tests/test_gpu_pipeline.py).  Device-resident clouds profit too: a second build in flight
fills the SMs that the partial last wave of every kernel and the few tiny kernels of the
first leave idle (measured: 0.786 -> 0.690 ms per 10 M-point cloud, tools/two_streams.py).  Nothing here computes: it is stream plumbing above the ABI.
"""
from collections import deque
from typing import Optional

import torch

from .builder import TwoDmap


class CloudPipeline:
    def __init__(self, res: float = 0.5, zres: float = 0.1, interval: float = 0.08, demand: str = "slope",
                 depth: int = 2, device: Optional[int] = None):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.demand = demand
        self.maps = []
        for _ in range(depth):
            m = TwoDmap(res, zres, device=device)
            m.setInterval(interval)
            m.pin_results(True)
            self.maps.append(m)
        dev = torch.device("cuda", self.maps[0]._device)
        self.streams = [torch.cuda.Stream(dev) for _ in range(depth)]  # non-blocking streams
        self._next = 0
        self._inflight = deque()

    @property
    def depth(self) -> int:
        return len(self.maps)

    @property
    def pending(self) -> int:
        """Builds submitted and not yet collected / released."""
        return len(self._inflight)

    def submit(self, cloud, origin=None) -> int:
        """Start one build (asynchronous).  `cloud`: float32 [n, >=3], a pinned host tensor for the
        overlap to happen (pageable memory works but the copy then blocks the caller).  The
        origin is point 0 as in chatterCallback unless `origin` is given.  Blocks only if all
        `depth` builders are busy — collect() first in that case."""
        if len(self._inflight) == self.depth:
            raise RuntimeError("all builders busy: collect() before submitting another cloud")
        slot = self._next
        self._next = (slot + 1) % self.depth
        m = self.maps[slot]
        if origin is None:
            m._p.origin_is_first_point = 1
        else:
            m.setCloudFirst(origin)
        if isinstance(cloud, torch.Tensor) and cloud.is_cuda:
            # a device-resident cloud was produced on the caller's current stream
            self.streams[slot].wait_stream(torch.cuda.current_stream(cloud.device))
        m.uniformDivision(cloud)
        m.create2DMap(self.demand, stream=self.streams[slot].cuda_stream)
        self._inflight.append(slot)
        return slot

    def release(self) -> TwoDmap:
        """Drop the OLDEST outstanding build from the queue WITHOUT waiting or copying anything
        (its tables stay on the device: `map.device_voxels()`, or read them later through the
        returned TwoDmap).  Reusing the builder is safe: its next build is ordered behind this
        one on the same stream."""
        if not self._inflight:
            raise RuntimeError("nothing submitted")
        return self.maps[self._inflight.popleft()]

    def join(self):
        """Make the caller's current stream wait for everything submitted so far."""
        cur = torch.cuda.current_stream(torch.device("cuda", self.maps[0]._device))
        for s in self.streams:
            cur.wait_stream(s)

    def collect(self) -> dict:
        """Wait for the OLDEST outstanding build and return its tables (views of the builder's
        pinned buffers: valid until that builder is used again, `depth` submits later)."""
        if not self._inflight:
            raise RuntimeError("nothing submitted")
        m = self.maps[self._inflight.popleft()]
        return {"voxels": m.voxels, "slopes": m.slopes, "columns": m.columns, "counts": m.counts(), "map": m}

    def close(self):
        for m in self.maps:
            m.close()
