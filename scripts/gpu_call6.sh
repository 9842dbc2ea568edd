#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/c6_pytest.log
./sfm_gms_b200/cxx/demo_multi 0 48 10000 0 0 | tee gpurun_out/c6_demo_multi.log
./sfm_gms_b200/cxx/demo_multi 1 48 10000 0 0 | tee -a gpurun_out/c6_demo_multi.log
