#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/c19_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c19_pytest.log
timeout 60 python scripts/stress.py 20 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_default.err
python - <<PY
import json
j=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
print('value %.0f ms/step %.3f e2e %.0f'%(j['value'], j['ms_per_step'], j['e2e']['value']), j['stage_ms_per_step'], j['roofline']['frac'], j['roofline'].get('frac_kernel'))
a=j['allpairs']; print('allpairs value %.0f e2e %.0f e2e8 %.0f'%(a['value'], a['e2e']['value'], a['e2e_index_pairs']['value']), a['stage_ms_per_step'])
PY
