#!/usr/bin/env python
"""Regenerates the derived evidence files under profiles/ (runs in the build container, no GPU):

  python scripts/make_evidence.py sass                    -> profiles/r2_sass_opcodes.txt (cuobjdump -sass of the in-tree .so)
  python scripts/make_evidence.py traffic <rep> [pairs]   -> profiles/r2_traffic.json + profiles/r2_top_raw.csv from an
                                                             `ncu --set full` report (per-kernel duration, DRAM bytes, pipes)
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sfm_gms_b200", "libsfmgms.so")
OPS = ["UTCOMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "SYNCS", "ELECT", "BRA.U.ANY", "FMNMX3", "FADD2", "POPC", "ATOMS", "DFMA",
       "CCTL", "STS"]


def short(mangled):
    out = subprocess.run(["cu++filt", mangled], capture_output=True, text=True).stdout.strip() or mangled
    m = re.search(r"(\w+_kernel)(<[^>]*>)?", out)
    return (m.group(1) + (m.group(2) or "")) if m else out[:40]


def sass():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = short(m.group(1))
            counts.setdefault(cur, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    counts[cur][o] += 1
    lines = ["SASS opcode counts per kernel of sfm_gms_b200/libsfmgms.so (cuobjdump -sass, occurrences per kernel; sm_100a; "
             "scripts/make_evidence.py sass)", ""]
    lines.append("%-36s" % "kernel" + "".join("%10s" % o for o in OPS))
    tot = collections.Counter()
    for k, c in counts.items():
        if not k.endswith("kernel") and "kernel<" not in k:
            continue
        lines.append("%-36s" % k[:36] + "".join("%10d" % c[o] for o in OPS))
        tot.update(c)
    lines.append("%-36s" % "TOTAL" + "".join("%10d" % tot[o] for o in OPS))
    lines += ["",
              "UTCOMMA = tcgen05.mma kind::mxf4.block_scale (block-scaled FP4); UTCIMMA = tcgen05.mma kind::i8; LDTM / STTM = tcgen05.ld / st",
              "(tensor memory); UTMALDG = cp.async.bulk.tensor (TMA load); UTCBAR = tcgen05.commit; SYNCS = mbarrier operations; ELECT = elect.sync;",
              "BRA.U.ANY = the serialisation loop ptxas wraps around a uniform-datapath instruction issued from divergent code (0 in every",
              "tensor kernel since the converged-role-warp + elect.sync issue path); FMNMX3 / FADD2 = the 3-input max and packed fp32 add of the",
              "fp4 epilogue; POPC in hamming_fp4_kernel<0> = the fused tie resolution (warps 14-15 re-score the 8 candidates of every query row",
              "on the packed descriptors); ATOMS = shared-memory atomics (GMS histograms).  hamming_fp4_kernel<0> is the production",
              "instantiation; <1>, <3>-<7> are the timing ablations (SFMGMS_TC_DEBUG)."]
    open(os.path.join(ROOT, "profiles", "r2_sass_opcodes.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


def traffic(rep, pairs):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    open(os.path.join(ROOT, "profiles", "r2_top_raw.csv"), "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    h = rows[0]

    def col(r, name):
        return float(r[h.index(name)].replace(",", "") or 0) if name in h else None

    def scaled(r, name, units):   # ncu prints a unit row: normalise to base units
        v = col(r, name)
        u = units[h.index(name)] if name in h else ""
        mul = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1e-3, "ms": 1, "ns": 1e-6, "s": 1e3, "usecond": 1e-3, "msecond": 1,
               "nsecond": 1e-6, "second": 1e3}.get(u, 1)
        return None if v is None else v * mul

    units = rows[1]
    allk = collections.OrderedDict()
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[h.index("Kernel Name")]).split("::")[-1]
        name = re.sub(r"<.*", "", name)
        allk[name] = {
            "duration_ms": scaled(r, "gpu__time_duration.sum", units),
            "dram_bytes_read": scaled(r, "dram__bytes_read.sum", units),
            "dram_bytes_write": scaled(r, "dram__bytes_write.sum", units),
            "tensor_pct_elapsed": col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            "issue_active_pct": col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "alu_pct": col(r, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
            "regs": col(r, "launch__registers_per_thread"),
            "sm_throughput_pct": col(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        }
    k = allk.get("hamming_fp4_kernel", {})
    out = {"kernel": "hamming_fp4_kernel", "file": "r2_traffic.json", "pairs_per_launch": pairs,
           "dram_bytes_read": k.get("dram_bytes_read"), "dram_bytes_write": k.get("dram_bytes_write"),
           "source": "profiles/r2_top_raw.csv (ncu --set full, bench.py --steps 2 --warmup 3 --no-allpairs --no-cpu-baseline; "
                     "scripts/make_evidence.py traffic)", "all_kernels": allk}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sass":
        sass()
    elif len(sys.argv) > 2 and sys.argv[1] == "traffic":
        traffic(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 256)
    else:
        print(__doc__)
