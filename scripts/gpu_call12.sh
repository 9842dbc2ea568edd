#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c12_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/c12_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --workload cfg3 --no-cpu-baseline > gpurun_out/c12_cfg3.json 2> gpurun_out/c12_cfg3.err
timeout 100 python scripts/latency.py > gpurun_out/c12_latency.log 2>&1; cat gpurun_out/c12_latency.log
python - <<'PY'
import json
for f in ('c12_cfg3',):
    j=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f,'value %.0f ms/step %.3f'%(j['value'], j['ms_per_step']), j['stage_ms_per_step'], {k['kernel']:round(k['ms_per_step'],4) for k in j['roofline_kernels'] if k['kernel'].startswith('gms')})
PY
