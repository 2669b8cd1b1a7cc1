#!/bin/bash
cd "$(dirname "$0")/../.."
for cfg in "" "STARR_B200_LIB=$PWD/gpurun_in_lib_hy4096.so" "ST_OVERLAP=0" "ST_OVERLAP=0 STARR_B200_LIB=$PWD/gpurun_in_lib_hy4096.so"; do
  echo "== $cfg"
  env $cfg ST_TIMELINE=1 timeout 100 python tools/debug/few_updates.py 5 2>&1 | grep timeline | tail -2
  env $cfg timeout 100 python tools/debug/search_time.py 2>&1 | tail -1
done
