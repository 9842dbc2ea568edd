"""Summarise an .ncu-rep: key raw metrics per kernel + top stall reasons per source line (needs ncu locally)."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__cycles_active.avg",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled"]
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    if pat and pat not in name:
        continue
    print("==", name[:100])
    for i, n in enumerate(h):
        if any(n.startswith(k) for k in KEYS) or "tensor" in n.lower() or "tmem" in n.lower():
            print("   %-95s %-10s %s" % (n, units[i], r[i]))
    for i, n in enumerate(h):
        if "warp_issue_stalled" in n and n.endswith("_per_warp_active.pct") and float(r[i] or 0) > 1.0:
            print("   STALL %-85s %s" % (n, r[i]))
