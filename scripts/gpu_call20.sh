#!/bin/bash
# V3 epilogue (3 TMEM slots x 160 columns, warp owns a slot): parity + A/B.  Every step under a hard timeout (a wrong barrier would hang).
mkdir -p gpurun_out
V3=$PWD/sfm_gms_b200/libsfmgms_v3.so
SFMGMS_LIB=$V3 timeout -s KILL 150 python -m pytest tests/test_gpu_parity.py tests/test_gpu_compact.py -m gpu -q -x > gpurun_out/c20_pytest_v3.log 2>&1; echo "pytest v3 rc=$?"; tail -3 gpurun_out/c20_pytest_v3.log
SFMGMS_LIB=$V3 timeout -s KILL 60 python scripts/stress.py 25 2>&1 | tail -2
for lib in libsfmgms_v3.so libsfmgms.so; do
  echo "== $lib"
  SFMGMS_LIB=$PWD/sfm_gms_b200/$lib SFMGMS_KERNEL=fp4 timeout -s KILL 60 python scripts/tc_time.py 256 2>&1 | tail -3
done
