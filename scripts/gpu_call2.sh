#!/bin/bash
# measurement batch run on the GPU box (gpurun): unit peaks, full GPU test suite, bench A/B of the GMS kernels, cfg3 / cfg4
mkdir -p gpurun_out
python scripts/unit_peaks.py > gpurun_out/c2_peaks.log 2>&1; tail -1 gpurun_out/c2_peaks.log
python -m pytest tests -m gpu -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log; tail -15 gpurun_out/c2_pytest.log
python bench.py --steps 10 --warmup 3 --no-allpairs > gpurun_out/c2_bench_v2.json 2> gpurun_out/c2_bench_v2.err; echo "bench rc=$?"
SFMGMS_GMS_V1=1 python bench.py --steps 10 --warmup 3 --no-allpairs --no-cpu-baseline > gpurun_out/c2_bench_v1.json 2> gpurun_out/c2_bench_v1.err
python bench.py --steps 5 --warmup 3 --workload cfg3 > gpurun_out/c2_bench_cfg3.json 2> gpurun_out/c2_bench_cfg3.err; echo "cfg3 rc=$?"
python bench.py --steps 5 --warmup 3 --workload cfg4 > gpurun_out/c2_bench_cfg4.json 2> gpurun_out/c2_bench_cfg4.err; echo "cfg4 rc=$?"
tail -c 600 gpurun_out/c2_bench_cfg3.err gpurun_out/c2_bench_cfg4.err gpurun_out/c2_bench_v2.err
