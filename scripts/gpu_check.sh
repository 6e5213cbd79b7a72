#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + full capture of the hot kernel.
# Usage (from the repo root on the GPU box): bash scripts/gpu_check.sh [quick]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" 
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
echo "== bench"
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
if [ "${1:-}" != "quick" ]; then
echo "== bench bert_adam"
timeout 600 python bench.py --workload bert_adam --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bert.json 2> gpurun_out/bench_bert.err; echo "exit $?"; cat gpurun_out/bench_bert.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "exit $?"
echo "== ncu full on dense_fused"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_fused -s 2 -c 2 -f -o gpurun_out/prof_dense python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "exit $?"
echo "== compute-sanitizer memcheck on smoke"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/memcheck.log 2>&1; echo "memcheck exit $?"; tail -5 gpurun_out/memcheck.log
fi
echo done
