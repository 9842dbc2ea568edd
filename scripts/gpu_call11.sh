#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gms_dll.py tests/test_gpu_compact.py tests/test_gpu_parity.py -m gpu -q -x -k "gms or compact or match or microcase or config" > gpurun_out/c11_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c11_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-allpairs --no-cpu-baseline > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err; echo "bench rc=$?"
for w in cfg3 cfg4; do
  B="python bench.py --steps 2 --warmup 3 --workload $w --no-cpu-baseline"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_$w.csv $B > gpurun_out/c11_ncu_$w.log 2>&1; echo "ncu launches $w rc=$?"
  ncu --set full --import-source on --clock-control none -k regex:"gms_" --launch-skip 12 -c 8 -f -o gpurun_out/r2_gms_$w $B > gpurun_out/c11_ncu2_$w.log 2>&1; echo "ncu full $w rc=$?"
done
python - <<'PY'
import json
j=json.loads(open('gpurun_out/c11_bench.json').read().strip().splitlines()[-1])
print('value %.0f ms/step %.3f'%(j['value'], j['ms_per_step']), j['stage_ms_per_step'], {k['kernel']:round(k['ms_per_step'],4) for k in j['roofline_kernels'] if k['kernel'].startswith('gms')})
PY
