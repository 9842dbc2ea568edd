#!/usr/bin/env python
"""Build and run scripts/epi_probe.cu (tensor / TMEM-read / max-tree unit rates, alone and mixed) -> profiles/r2_epi_probe.txt.

  python scripts/epi_probe.py --build     # compile only (nvcc cross-compiles sm_100a)
  python scripts/epi_probe.py             # on the GPU box
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EXE = os.path.join(HERE, "epi_probe")

if __name__ == "__main__":
    src = os.path.join(HERE, "epi_probe.cu")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < os.path.getmtime(src):
        subprocess.check_call([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-gencode", "arch=compute_100a,code=sm_100a",
                               "-O3", "-lineinfo", "-std=c++17", src, "-o", EXE])
    if "--build" in sys.argv:
        sys.exit(0)
    out = subprocess.check_output([EXE], text=True, timeout=90)
    for d in ("gpurun_out", "profiles"):
        os.makedirs(os.path.join(ROOT, d), exist_ok=True)
        open(os.path.join(ROOT, d, "r2_epi_probe.txt"), "w").write(out)
    print(out)
