#!/usr/bin/env bash
# bench.py on N GPUs of one box, launched the way the driver does.
#   /usr/local/graft/bin/gpurun --gpus N --timeout 500 -- 'bash tools/scaling.sh N'
set -u
N=${1:?usage: tools/scaling.sh N}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus "$N" --steps 20 --warmup 3 > "gpurun_out/bench_n${N}_final.json" 2> "gpurun_out/bench_n${N}_final.err"
echo "rc=$?"; tail -1 "gpurun_out/bench_n${N}_final.err"
