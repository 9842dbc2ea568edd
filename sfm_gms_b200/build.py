"""Builds sfm_gms_b200/libsfmgms.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["capi.cu", "multi.cu", "hamming_popc.cu", "hamming_tc.cu", "hamming_fp4.cu", "gms.cu", "l2_dp4a.cu", "l2_tc.cu", "l2_f32.cu", "orb.cu"]
HEADERS = ["common.cuh", "capi_internal.h", "hamming_tc.cuh", "tc_ptx.cuh", "orb_pattern.inc", os.path.join("..", "..", "include", "sfmgms.h")]
LIB = os.path.join(HERE, "libsfmgms.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",            # GMS decisions must round each FP op separately
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    "-Xptxas", "-v",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: build an experimental variant (e.g. defines=["-DSFMGMS_TC_MSUB=3"], out="libsfmgms_msub3.so")."""
    global LIB
    if out:
        LIB_SAVE = LIB
        try:
            LIB = os.path.join(HERE, out)
            return _build(True, verbose, list(defines), os.path.join(HERE, "build_" + out.replace(".so", "")))
        finally:
            LIB = LIB_SAVE
    if not force and not _stale():
        return LIB
    return _build(force, verbose, [], os.path.join(HERE, "build"))


def _build(force, verbose, defines, bdir):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for s in SOURCES:
        o = os.path.join(bdir, s.replace(".cu", ".o"))
        objs.append(o)
        cmd = [nvcc] + NVCC_FLAGS + defines + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for s, p in procs:
        out, _ = p.communicate()
        log.append("== %s ==\n%s" % (s, out))
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError("nvcc failed on %s" % s)
    with open(os.path.join(bdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
