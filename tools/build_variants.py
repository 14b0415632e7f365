"""Compile tuning variants of libgndt.so into _variants/ (git-ignored, travels with gpurun).
usage: build_variants.py name=-DGNDT_SORT_THREADS=256,-DGNDT_SORT_ITEMS=16 ...   (A/B them with GNDT_LIB=...)"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from grid_ndt_b200._lib import NVCC_FLAGS, SOURCES
os.makedirs(os.path.join(ROOT, "_variants"), exist_ok=True)


def one(spec):
    name, _, defs = spec.partition("=")
    out = os.path.join(ROOT, "_variants", f"libgndt_{name}.so")
    cmd = ["nvcc"] + NVCC_FLAGS + ["-Xptxas", "-v"] + [d for d in defs.split(",") if d] + ["-o", out, SOURCES[0], "-I" + os.path.join(ROOT, "include")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    lines = (r.stdout + r.stderr).split("\n")
    info = []
    for i, l in enumerate(lines):
        if "Compiling entry function" in l and ("sort_pass_kernelILb0ELb1" in l or "reduce_kernelILb1" in l or "finalize" in l):
            info.append(l.split("'")[1][:40] + " | " + " ".join(x.strip() for x in lines[i + 1:i + 4] if "Used" in x or "spill" in x))
    return name, r.returncode, info, "" if r.returncode == 0 else r.stderr[-2000:]


with ThreadPoolExecutor(8) as ex:
    for name, rc, info, err in ex.map(one, sys.argv[1:]):
        print(name, "rc", rc)
        for i in info:
            print("   ", i)
        if err:
            print(err)
