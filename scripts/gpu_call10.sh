#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c10_pytest.log
timeout 120 python scripts/l2_time.py > gpurun_out/r2_l2_time.log 2>&1; cat gpurun_out/r2_l2_time.log
for k in tc fp4; do SFMGMS_KERNEL=$k timeout 120 python scripts/tc_time.py 256 2>&1 | grep -v sustained | tail -2; done | tee gpurun_out/c10_tc.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-allpairs --no-cpu-baseline > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --workload cfg3 --no-cpu-baseline > gpurun_out/c10_cfg3.json 2> gpurun_out/c10_cfg3.err
timeout 200 python scripts/stress.py 90 > gpurun_out/r2_stress.log 2>&1; tail -2 gpurun_out/r2_stress.log
python - <<'PY'
import json
for f in ('c10_bench','c10_cfg3'):
    j=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f,'value %.0f ms/step %.3f'%(j['value'], j['ms_per_step']), j['stage_ms_per_step'], {k['kernel']:round(k['ms_per_step'],4) for k in j['roofline_kernels'] if k['kernel'].startswith('gms')})
PY
