#!/bin/bash
# strong-scaling sweep of the all-pairs workload on one 8-GPU box (what the driver does at round end)
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 3 --warmup 3 --workload allpairs > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err; echo "n1 rc=$?"
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540+n)) bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r2_scale_n$n.json 2> gpurun_out/r2_scale_n$n.err; echo "n$n rc=$?"
done
python - <<'PY'
import json
for n in (1,2,4,8):
    j=json.loads(open('gpurun_out/r2_scale_n%d.json'%n).read().strip().splitlines()[-1])
    print(n,'value %.0f'%j['value'],'e2e %.0f'%j['e2e']['value'],'ms %.1f / %.1f'%(j['ms_per_step'], j['e2e']['ms_per_step']), 'bcast %.2f'%j['e2e']['upload_plus_broadcast_ms'], j['stage_ms_per_step'], j['parity_checked_pairs'])
PY
