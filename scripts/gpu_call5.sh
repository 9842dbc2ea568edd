#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py tests/test_gms_dll.py tests/test_gpu_compact.py -m gpu -q > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/c5_pytest.log
python bench.py --steps 5 --warmup 3 --workload cfg3 --no-cpu-baseline > gpurun_out/c5_bench_cfg3.json 2> gpurun_out/c5_bench_cfg3.err; echo "cfg3 rc=$?"
for th in 512 1024; do for b in 5 10; do SFMGMS_GMS_THREADS=$th SFMGMS_GMS_BAND0=$b python bench.py --steps 10 --warmup 3 --no-allpairs --no-cpu-baseline > gpurun_out/c5_bench_t${th}_b$b.json 2> gpurun_out/c5_bench_t${th}_b$b.err; echo "t$th b$b rc=$?"; done; done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/c5_bench_n2.json 2> gpurun_out/c5_bench_n2.err; echo "n2 rc=$?"; tail -c 800 gpurun_out/c5_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/c5_bench_n2_ref.json 2> gpurun_out/c5_bench_n2_ref.err; echo "n2 ref rc=$?"
./sfm_gms_b200/cxx/demo_multi 0 24 10000 0 0 | tee gpurun_out/c5_demo_multi.log
