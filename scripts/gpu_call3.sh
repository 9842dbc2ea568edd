#!/bin/bash
# ncu evidence for round 2: launch list of one bench command + full captures of the top kernels
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-allpairs --no-cpu-baseline"
$B > gpurun_out/c3_plain.json 2> gpurun_out/c3_plain.err; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/c3_ncu1.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --import-source on --clock-control none -k regex:"gms_vote2|hamming_fp4_kernel|hamming_resolve|gms_assign_cnt|gms_count|unpack_fp4" --launch-skip 12 -c 6 -f -o gpurun_out/r2_top $B > gpurun_out/c3_ncu2.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
