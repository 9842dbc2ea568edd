#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_scale_n8.json 2> gpurun_out/r2_scale_n8.err; echo "n8 rc=$?"; tail -c 600 gpurun_out/r2_scale_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_scale_n4.json 2> gpurun_out/r2_scale_n4.err; echo "n4 rc=$?"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/c_n8_pytest.log 2>&1; tail -3 gpurun_out/c_n8_pytest.log
./sfm_gms_b200/cxx/demo_multi 0 64 10000 0 0 | tee gpurun_out/r2_demo_multi_8gpu.log
./sfm_gms_b200/cxx/demo_multi 0 64 10000 0 0 | tee -a gpurun_out/r2_demo_multi_8gpu.log
