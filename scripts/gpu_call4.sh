#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gms_dll.py tests/test_gpu_compact.py tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -x -k "gms or compact or multi or match_pair or microcase or config" > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c4_pytest.log
for b in 5 10; do SFMGMS_GMS_BAND0=$b python bench.py --steps 10 --warmup 3 --no-allpairs --no-cpu-baseline > gpurun_out/c4_bench_band$b.json 2> gpurun_out/c4_bench_band$b.err; echo "band $b rc=$?"; done
python bench.py --steps 5 --warmup 3 --workload cfg3 --no-cpu-baseline > gpurun_out/c4_bench_cfg3.json 2> gpurun_out/c4_bench_cfg3.err; echo "cfg3 rc=$?"
for d in 0 1 3 4; do SFMGMS_KERNEL=fp4 SFMGMS_TC_DEBUG=$d python scripts/tc_time.py 256 2>&1 | grep -v sustained | tail -2; done > gpurun_out/c4_ablation.log 2>&1
SFMGMS_KERNEL=fp4 python scripts/tc_time.py 256 >> gpurun_out/c4_ablation.log 2>&1
cat gpurun_out/c4_ablation.log
