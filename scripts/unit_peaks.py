#!/usr/bin/env python
"""Build and run scripts/unit_peaks.cu (measured ceilings of the tensor / shared-atomic / POPC units on this B200) and
store its JSON line as profiles/r2_unit_peaks.json — bench.py's roofline denominators.

  python scripts/unit_peaks.py --build     # compile only (no GPU needed: nvcc cross-compiles sm_100a)
  python scripts/unit_peaks.py             # on the GPU box: run and write gpurun_out/r2_unit_peaks.json + profiles/
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EXE = os.path.join(HERE, "unit_peaks")


def build(force=False):
    src = os.path.join(HERE, "unit_peaks.cu")
    if force or not os.path.exists(EXE) or os.path.getmtime(EXE) < os.path.getmtime(src):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", src, "-o", EXE])
    return EXE


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    if "--build" in sys.argv:
        sys.exit(0)
    out = subprocess.check_output([EXE], text=True, timeout=300)
    rec = json.loads(out.strip().splitlines()[-1])
    for d in (os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")):
        os.makedirs(d, exist_ok=True)
        json.dump(rec, open(os.path.join(d, "r2_unit_peaks.json"), "w"), indent=1)
    print(json.dumps(rec))
