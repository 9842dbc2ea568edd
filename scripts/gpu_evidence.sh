#!/bin/bash
# round-2 evidence batch (one B200): bench lines, ncu launch list of the bench command, full ncu capture of the top kernels
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "default rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/r2_bench_default.err; echo "reference rc=$?"
python bench.py --steps 10 --warmup 3 --workload cfg3 > gpurun_out/r2_bench_cfg3.json 2>> gpurun_out/r2_bench_default.err; echo "cfg3 rc=$?"
python bench.py --steps 10 --warmup 3 --workload cfg4 > gpurun_out/r2_bench_cfg4.json 2>> gpurun_out/r2_bench_default.err; echo "cfg4 rc=$?"
python bench.py --steps 3 --warmup 3 --workload allpairs > gpurun_out/r2_bench_allpairs_n1.json 2>> gpurun_out/r2_bench_default.err; echo "allpairs rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-allpairs --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu1.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --import-source on --clock-control none -k regex:"gms_vote2|hamming_fp4_kernel|gms_assign_cnt|gms_count|unpack_fp4" --launch-skip 10 -c 5 -f -o gpurun_out/r2_top $B > gpurun_out/r2_ncu2.log 2>&1; echo "ncu full rc=$?"
python scripts/latency.py > gpurun_out/r2_latency.log 2>&1; cat gpurun_out/r2_latency.log
for d in 0 1 3 4 6; do SFMGMS_KERNEL=fp4 SFMGMS_TC_DEBUG=$d python scripts/tc_time.py 256 2>&1 | grep -v sustained | tail -1; done > gpurun_out/r2_ablation.log 2>&1
for v in "SFMGMS_FP4_FUSED_RESOLVE=0 SFMGMS_FP4_PACKED_QUERIES=0" "SFMGMS_FP4_FUSED_RESOLVE=1 SFMGMS_FP4_PACKED_QUERIES=0" "SFMGMS_FP4_FUSED_RESOLVE=1 SFMGMS_FP4_PACKED_QUERIES=1"; do
  echo "== $v" >> gpurun_out/r2_ablation.log
  env $v SFMGMS_KERNEL=fp4 python scripts/tc_time.py 256 2>&1 | tail -3 >> gpurun_out/r2_ablation.log
done
cat gpurun_out/r2_ablation.log
