#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bf or match or config" > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c7_pytest.log
python bench.py --steps 10 --warmup 3 --no-allpairs --no-cpu-baseline > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err; echo "bench rc=$?"
for d in 0 1 3 4; do SFMGMS_KERNEL=fp4 SFMGMS_TC_DEBUG=$d python scripts/tc_time.py 256 2>&1 | grep -v sustained | tail -2; done > gpurun_out/c7_ablation.log 2>&1
SFMGMS_KERNEL=fp4 python scripts/tc_time.py 256 >> gpurun_out/c7_ablation.log 2>&1
cat gpurun_out/c7_ablation.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/c7_bench.json').read().strip().splitlines()[-1])
print('value %.0f ms/step %.3f e2e %.0f'%(j['value'], j['ms_per_step'], j['e2e']['value']), j['stage_ms_per_step'], 'roof', j['roofline']['frac'])
for k in j['roofline_kernels']: print('  %-18s %.4f'%(k['kernel'],k['ms_per_step']))
PY
