#!/bin/bash
# update timing on fresh batches under a few tuning knobs (side-stream footprints, retire threshold builds)
cd "$(dirname "$0")/../.."
for cfg in "" "ST_DEEP_CTAS=1" "ST_DEEP_CTAS=4" "ST_RETIRE_CTAS=2" "ST_RETIRE_CTAS=6" "ST_DEEP_CTAS=1 ST_RETIRE_CTAS=2" \
           "STARR_B200_LIB=$PWD/gpurun_in_lib_r256.so" "STARR_B200_LIB=$PWD/gpurun_in_lib_r1024.so" "STARR_B200_LIB=$PWD/gpurun_in_lib_r2048.so" \
           "STARR_B200_LIB=$PWD/gpurun_in_lib_h8192.so" "STARR_B200_LIB=$PWD/gpurun_in_lib_h4096.so" "STARR_B200_LIB=$PWD/gpurun_in_lib_h2048.so"; do
  echo "== $cfg"
  env $cfg timeout 100 python tools/debug/sweep_lib.py 2>&1 | tail -1
done
