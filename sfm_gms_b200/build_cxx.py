"""Builds the C++ demo of the header-only shim (sfm_gms_b200/cxx/sfmgms.hpp) against libsfmgms.so."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cxx", "demo_match.cpp")
HDR = os.path.join(HERE, "cxx", "sfmgms.hpp")
OUT = os.path.join(HERE, "cxx", "demo_match")


def build(force=False):
    lib = os.path.join(HERE, "libsfmgms.so")
    deps = [SRC, HDR, lib, os.path.join(HERE, "..", "include", "sfmgms.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-Wextra", SRC, "-o", OUT, "-L" + HERE, "-lsfmgms",
                           "-Wl,-rpath,$ORIGIN/.."])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
