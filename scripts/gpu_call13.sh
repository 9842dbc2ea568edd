#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c13_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c13_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err; echo "bench rc=$?"
timeout 100 python scripts/stress.py 40 2>&1 | tail -1
python - <<'PY'
import json
j=json.loads(open('gpurun_out/c13_bench.json').read().strip().splitlines()[-1])
print('value %.0f ms/step %.3f e2e %.0f'%(j['value'], j['ms_per_step'], j['e2e']['value']), j['stage_ms_per_step'])
a=j['allpairs']; print('allpairs value %.0f e2e %.0f'%(a['value'], a['e2e']['value']), a['stage_ms_per_step'])
PY
