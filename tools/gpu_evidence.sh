#!/usr/bin/env bash
# One gpurun call that produces everything profiles/ needs for a 1-GPU round-end refresh.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_evidence.sh'
#   python tools/ncu_summarize.py rN_final gpurun_out/launches_final.csv gpurun_out/prof_final.ncu-rep
#   cp gpurun_out/bench_final.json profiles/rN_final_bench_1gpu.json
# ncu runs only after the identical command has exited 0 without it; numbers printed under ncu are never
# bench values.  Multi-GPU lines: tools/scaling.sh N (one gpurun --gpus N call per N).
set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference arm rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_a.json 2>/dev/null &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
echo "launch list rc=$?"
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_b.json 2>/dev/null &&
  ncu --set full --clock-control none --import-source on \
      -k regex:"hy_hist|hy_scan|hy_scatter|hy_segment|deep_red|retire_kernel|search_kernel|fold_huge_maps|fold_huge_apply|fold_scan|leaf_pass|build_pass|make_keys|RadixSortOnesweep" \
      -c 24 -f -o gpurun_out/prof_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
echo "full capture rc=$?"
python tools/config_bench.py > gpurun_out/configs_final.md 2> gpurun_out/configs_final.err; echo "per-config table rc=$?"
