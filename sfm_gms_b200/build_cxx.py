"""Builds the C++ demos of the header-only shim (sfm_gms_b200/cxx/sfmgms.hpp) against libsfmgms.so:
demo_match (one pair, the reference's two call lines) and demo_multi (an image sequence on several GPUs of one box)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
HDR = os.path.join(HERE, "cxx", "sfmgms.hpp")
OUT = os.path.join(HERE, "cxx", "demo_match")
OUT_MULTI = os.path.join(HERE, "cxx", "demo_multi")


def _build_one(name, force):
    src, out = os.path.join(HERE, "cxx", name + ".cpp"), os.path.join(HERE, "cxx", name)
    lib = os.path.join(HERE, "libsfmgms.so")
    deps = [src, HDR, lib, os.path.join(HERE, "..", "include", "sfmgms.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-Wextra", src, "-o", out, "-L" + HERE, "-lsfmgms",
                           "-Wl,-rpath,$ORIGIN/.."])
    return out


def build(force=False):
    _build_one("demo_multi", force)
    return _build_one("demo_match", force)


if __name__ == "__main__":
    print(build(force=True))
