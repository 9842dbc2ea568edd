#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/c18_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c18_pytest.log
timeout 100 python scripts/stress.py 40 2>&1 | tail -2
for lib in libsfmgms.so libsfmgms_pm1.so; do
  echo "== $lib"
  SFMGMS_LIB=$PWD/sfm_gms_b200/$lib SFMGMS_KERNEL=fp4 timeout 120 python scripts/tc_time.py 256 2>&1 | tail -3
done
python scripts/latency.py 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c18_bench.json 2> gpurun_out/c18_bench.err; echo "bench rc=$?"
python - <<PY
import json
j=json.loads(open('gpurun_out/c18_bench.json').read().strip().splitlines()[-1])
print('value %.0f ms/step %.3f e2e %.0f'%(j['value'], j['ms_per_step'], j['e2e']['value']), j['stage_ms_per_step'])
a=j['allpairs']; print('allpairs value %.0f e2e %.0f'%(a['value'], a['e2e']['value']), a['stage_ms_per_step'])
print([(k['kernel'], round(k['ms_per_step'],4)) for k in j['roofline_kernels']])
PY
