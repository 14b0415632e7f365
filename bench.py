#!/usr/bin/env python
"""bench.py — NDT map-build throughput (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] ...    the reference's own CPU code

A "step" is one complete map build (bounds -> partition -> fit -> labels -> edges) of one
synthetic cloud.  N = 1: BASELINE.json configs[1] (cfg2: 10 M points, multi-level bridge /
underpass scene, 0.2 m cells).  N > 1: one x strip per GPU, each strip one cfg2 scene
(weak scaling: 10 M points per GPU), thin halo rows swapped between neighbour strips and
the finished strips gathered over NCCL inside the timed region.

`value`   : points/s, device-timed with CUDA events, inputs resident in HBM, max over ranks,
            K back-to-back builds with two in flight on separate streams (a second build fills
            the SMs idle in the last wave of every kernel; at N > 1 it also hides the NVLink
            gather).  `serial_ms_per_step` is the device time of one build run alone.
`e2e`     : the same metric through the public TwoDmap call with PINNED HOST input, the
            host->device copy of the cloud and the device->host copy of the voxel / slope /
            column tables inside the timed region, every step.  At N = 1 the headline e2e
            runs two builders deep (CloudPipeline: upload of cloud i+1 overlaps build and
            read-back of cloud i), at N > 1 the two-deep strip pipeline with host input;
            the one-at-a-time figure is `serial_ms_per_step`.
`roofline`: algorithmic bytes of one build (16 B per point read once + 96 B per voxel
            record written once) / device time per build, against the
            measured HBM copy bandwidth in MEASURED_PEAKS.json.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CFG = "cfg2"
POINTS_PER_GPU = 10_000_000
GRID_LEN, Z_LEN, INTERVAL = 0.2, 0.1, 0.08
SCENE_W = 120.0
DEPTH = int(os.environ.get("GNDT_BENCH_DEPTH", "3"))  # N > 1: builds in flight (4 GPUs: 3 -> 0.99 ms, 2 -> 1.07 ms per step)
REF_STEP_POINTS = 1_000_000      # --impl reference: points per timed step (bounded sample; smaller samples flatter the CPU code)
CPU_BASELINE_POINTS = 4_000_000  # cpu_baseline leg of the default run


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json (device copy, measured)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


class ClockSampler:
    """Polls SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {}
        for k in dir(nv):
            if k.startswith("nvmlClocksEventReason") or k.startswith("nvmlClocksThrottleReason"):
                v = getattr(nv, k)
                if isinstance(v, int) and v:
                    names[v] = k.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", "")
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit and bin(bit).count("1") == 1:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        reasons = sorted(r for r in self.reasons if r not in ("GpuIdle", "None", "ApplicationsClocksSetting"))
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.samples)}


def make_cloud(rank):
    """cfg2 scene for strip `rank`, shifted by rank * 120 m in x (common frame)."""
    from grid_ndt_b200 import synthetic
    cloud = synthetic.cfg2(POINTS_PER_GPU)
    if rank:
        nz = np.abs(cloud[:, :3]).sum(axis=1) > 0  # the (0,0,0) padding stays where it is
        cloud[nz, 0] += np.float32(SCENE_W * rank)
    return cloud


def run_reference(args):
    """The reference's own CPU implementation of the path (oracle/_ref: src/receiver.cpp +
    include/map2D.h compiled against inert shims), single-threaded like its initial-build
    loop (the OpenMP pragma at src/receiver.cpp:149 is commented out)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from grid_ndt_b200 import synthetic
    from grid_ndt_b200._abi import default_params
    from oracle import oracle as O

    kind = "reference" if os.path.exists(O.REF_SO) else "port"
    fn = O.ref_build if kind == "reference" else (lambda c, p: O.oracle_build(c, p, "faithful32"))
    cloud = synthetic.cfg2(POINTS_PER_GPU)[:REF_STEP_POINTS]
    p = default_params(GRID_LEN, Z_LEN, INTERVAL)
    for _ in range(args.warmup):
        fn(cloud[: REF_STEP_POINTS // 8], p)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn(cloud, p)
    dt = time.perf_counter() - t0
    value = REF_STEP_POINTS * args.steps / dt
    sample = f"first {REF_STEP_POINTS} points of the 10M-point cfg2 cloud per step (sweep order), {kind} build"
    line = {
        "impl": "reference", "metric": "ndt_map_build_points_per_sec", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2 multi-level bridge/underpass scene, 0.2 m cells, z 0.1 m, slope_interval 0.08 (BASELINE configs[1])",
                   "points_per_step": REF_STEP_POINTS},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    args.json_out.write(json.dumps(line) + "\n")
    args.json_out.flush()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly ONE line, the JSON: everything else any library writes to fd 1
    # (NCCL's version banner, for one) is sent to stderr for the whole run
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args.json_out = json_out
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from grid_ndt_b200 import TwoDmap

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cloud = make_cloud(rank)
    n_pts = cloud.shape[0]
    host = torch.from_numpy(cloud).pin_memory()
    resident = host.to(dev)
    origin = [float(np.float32(0.5 * SCENE_W + 0.013)), float(np.float32(40.007)), 1.0]  # cfg2's designated first point

    if world > 1:
        from grid_ndt_b200.tiles import TiledTwoDmap
        # DEPTH builders deep: the NVLink gather of build i overlaps the SM work of the next builds
        tmp = TiledTwoDmap(GRID_LEN, Z_LEN, INTERVAL, rank, world, device=local, depth=DEPTH)
        tm = TiledTwoDmap(GRID_LEN, Z_LEN, INTERVAL, rank, world, device=local, depth=1)  # one at a time (e2e)
        m = tm.map

        def step(src):
            tm.build(src, "slope", origin=origin, cuts=None, filter_points=False)

        def run_steps(src, k):
            n_launch = 0
            ahead = 0
            for i in range(k):
                while ahead < min(k, i + tmp.depth):  # keep `depth` builds in flight
                    tmp.submit(src, "slope", origin=origin, cuts=None, filter_points=False)
                    ahead += 1
                tmp.collect()
                n_launch += tmp.last_map.launch_count()
            tmp.join()
            return n_launch
    else:
        from grid_ndt_b200.pipeline import CloudPipeline
        m = TwoDmap(GRID_LEN, Z_LEN, device=local)  # one build at a time: stage times, latency, e2e serial
        m.setInterval(INTERVAL)
        # throughput: two builds in flight on two streams (a second build fills the SMs left idle
        # by the partial last wave of every kernel and by the tiny kernels of the first)
        flight = CloudPipeline(GRID_LEN, Z_LEN, INTERVAL, "slope", depth=2, device=local)

        def step(src):
            m.chatterCallback(src, "slope")

        def run_steps(src, k):
            n_launch = 0
            for _ in range(k):
                if flight.pending == flight.depth:
                    n_launch += flight.release().launch_count()
                flight.submit(src)
            while flight.pending:
                n_launch += flight.release().launch_count()
            flight.join()
            return n_launch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run_steps(resident, args.warmup)
    step(resident)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        ev0.record()
        launches = run_steps(resident, args.steps)
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    total_pts = n_pts * world
    value = total_pts / (ms_per_step * 1e-3)
    # one build at a time (latency of a single cloud), same events, outside the headline region
    barrier()
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_serial = max(3, min(args.steps, 10))
    es0.record()
    for _ in range(n_serial):
        step(resident)
    es1.record()
    barrier()
    serial_ms = es0.elapsed_time(es1) / n_serial
    counts = m.counts()
    stages = m.stage_ms()

    # ---- end to end: pinned host cloud in, tables out (into pinned host buffers), every step
    m.pin_results(True)

    def e2e_step():
        step(host)
        v, s, c = m.voxels, m.slopes, m.columns
        return v.nbytes + s.nbytes + c.nbytes

    def wall(fn, n):
        barrier()
        t0 = time.perf_counter()
        fn(n)
        barrier()
        dt = (time.perf_counter() - t0) / n
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    d2h = 0
    for _ in range(2):
        d2h = e2e_step()
    n_e2e = max(3, min(args.steps, 10))
    serial_s = wall(lambda n: [e2e_step() for _ in range(n)], n_e2e)
    e2e_s, e2e_mode = serial_s, "one build at a time: H2D -> kernels (+ NCCL gather at N>1) -> D2H"
    if world == 1:
        # the same call, two builders deep: the upload of cloud i+1 overlaps the build and
        # read-back of cloud i (grid_ndt_b200.pipeline.CloudPipeline).  Every step still
        # uploads its whole cloud and reads back all three tables inside the timed region.
        from grid_ndt_b200.pipeline import CloudPipeline
        pipe = CloudPipeline(GRID_LEN, Z_LEN, INTERVAL, "slope", depth=2, device=local)

        def piped(n):
            pipe.submit(host)
            for i in range(n):
                if i + 1 < n:
                    pipe.submit(host)
                r = pipe.collect()
                assert r["voxels"].nbytes + r["slopes"].nbytes + r["columns"].nbytes == d2h
        piped(3)
        e2e_s = wall(piped, n_e2e)
        e2e_mode = "CloudPipeline depth 2: H2D of cloud i+1 overlaps kernels + D2H of cloud i"
        pipe.close()
    else:
        # N > 1: the two-deep strip pipeline with HOST input: upload of cloud i+1 overlaps the
        # gather and the read-back of this rank's strip tables of cloud i
        for sl in tmp.slots:
            sl.map.pin_results(True)

        def piped(n):
            tmp.submit(host, "slope", origin=origin, cuts=None, filter_points=False)
            for i in range(n):
                if i + 1 < n:
                    tmp.submit(host, "slope", origin=origin, cuts=None, filter_points=False)
                tmp.collect()
                mm = tmp.last_map
                got = mm.voxels.nbytes + mm.slopes.nbytes + mm.columns.nbytes
                assert got == d2h, (got, d2h)
            tmp.synchronize()
        piped(3)
        e2e_s = wall(piped, n_e2e)
        e2e_mode = "TiledTwoDmap depth 2, host input: H2D of cloud i+1 overlaps NCCL gather + D2H of cloud i"
    e2e_value = total_pts / e2e_s

    peak, peak_src = measured_peak()
    v_tab = counts["n_voxels"]
    b_alg = 16.0 * n_pts + 96.0 * v_tab            # per GPU per build
    # device time of the build itself on this rank (stage events; excludes the NCCL gather)
    t_build_ms = stages["total"]
    achieved = b_alg / (t_build_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("build_dram_bytes")
        except Exception:
            traffic = None

    layout = m.key_layout()
    n_pass = max(1, layout["passes"])
    pass_ms = stages["sort"] / n_pass
    pass_bytes = 32.0 * n_pts  # every point read once and written once per pass
    dominant = {"name": "sort_pass_kernel", "launches_per_build": n_pass, "share_of_build": stages["sort"] / stages["total"],
                "algorithmic_bytes_per_launch": pass_bytes, "avg_launch_ms": pass_ms,
                "achieved": pass_bytes / (pass_ms * 1e-3) / 1e9, "frac": pass_bytes / (pass_ms * 1e-3) / 1e9 / peak,
                "note": "the same memory pattern with no other work runs at 4.9 TB/s (tools/micro/scatter_pattern.cu, 65 us)"}

    line = {
        "metric": "ndt_map_build_points_per_sec", "value": value, "unit": "points/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "ms_per_10M_points": ms_per_step * 1e7 / total_pts, "serial_ms_per_step": serial_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2 multi-level bridge/underpass scene, 10M points per GPU, 0.2 m cells, z 0.1 m, slope_interval 0.08, demand slope (BASELINE configs[1])",
                   "points_per_gpu": n_pts, "voxels_per_gpu": v_tab, "columns_per_gpu": counts["n_columns"], "slopes_per_gpu": counts["n_slopes"],
                   "l2": "inputs (160 MB) and work buffers (320 MB) exceed the 126 MB L2; no explicit flush",
                   "parallelism": (f"x-strips x{world}: thin halo swap + NCCL gather of finished strips; " if world > 1 else "single GPU; ")
                                  + (f"{DEPTH if world > 1 else 2} builds in flight on separate streams (serial_ms_per_step = one build at a time)")},
        "stage_ms": stages,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "algorithmic_bytes": b_alg, "kernel_ms": t_build_ms,
                     "dominant_kernel": dominant,
                     "achieved_in_flight": (b_alg / (ms_per_step * 1e-3) / 1e9) if world == 1 else None,
                     "what": "whole build (all kernels of one step) on one GPU, one build at a time: (16 B x points + 96 B x voxels) / "
                             "device time of the build (stage events); achieved_in_flight = the same bytes / ms_per_step of the timed "
                             "region (two builds in flight); peak = " + peak_src},
        "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": int(n_pts * 16), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_s * 1e3, "mode": e2e_mode, "serial_ms_per_step": serial_s * 1e3},
        "gpu_launches": launches,
        "clocks": clk.summary(),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from grid_ndt_b200._abi import default_params
        from oracle import oracle as O
        kind = "reference" if os.path.exists(O.REF_SO) else "port"
        fn = O.ref_build if kind == "reference" else (lambda c, p: O.oracle_build(c, p, "faithful32"))
        sample = cloud[:CPU_BASELINE_POINTS]
        t0 = time.perf_counter()
        r = fn(sample, default_params(GRID_LEN, Z_LEN, INTERVAL))
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": CPU_BASELINE_POINTS / dt, "unit": "points/s", "cores": 1, "kind": kind,
                                "sample": f"first {CPU_BASELINE_POINTS} points of the same 10M cloud, one build, {dt:.1f} s "
                                          f"(division {r.division_s:.1f} s + calculate {r.calculate_s:.1f} s); single thread like the reference's initial-build loop",
                                "host_cores_available": os.cpu_count()}
    if rank == 0:
        args.json_out.write(json.dumps(line) + "\n")
        args.json_out.flush()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
