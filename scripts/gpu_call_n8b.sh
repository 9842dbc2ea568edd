#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_scale_n8.json 2> gpurun_out/r2_scale_n8.err; echo "n8 rc=$?"; tail -c 400 gpurun_out/r2_scale_n8.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2_scale_n8.json').read().strip().splitlines()[-1])
print('value %.0f'%j['value'],'e2e %.0f'%j['e2e']['value'],'e2e8 %.0f'%j['e2e_index_pairs']['value'], 'ms %.1f / %.1f / %.1f'%(j['ms_per_step'], j['e2e']['ms_per_step'], j['e2e_index_pairs']['ms_per_step']), j['parity_checked_pairs'])
PY
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2
