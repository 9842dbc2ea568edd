import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
# find header line starting with "Address"
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')][0]
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
h = rows[0]
iS = h.index("# Samples"); iSrc = h.index("Source"); iEx = h.index("Instructions Executed")
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
data = []
for k, r in enumerate(rows[1:]):
    if len(r) <= iS or r[0].startswith("Kernel") or not r[iS].isdigit():
        continue
    data.append((int(r[iS]), k, r))
tot = sum(d[0] for d in data)
print("total samples", tot)
for s, k, r in sorted(data, reverse=True)[:top]:
    st = sorted(((int(r[i]) if r[i].isdigit() else 0, h[i]) for i in stall_cols), reverse=True)[:3]
    print("%6d %5.1f%% line%5d ex=%9s  %-60s %s" % (s, 100.0 * s / tot, k, r[iEx], r[iSrc].strip()[:60], " ".join("%s=%d" % (n[6:], v) for v, n in st if v)))
