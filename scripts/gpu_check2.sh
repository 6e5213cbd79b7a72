#!/bin/bash
# 2-GPU visit: the whole -m gpu suite (incl. multi-GPU, group, allreduce fallback), rsp bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu_all.log
echo "== rsp bench (2 GPUs, 8 values)"
timeout 600 python bench_rsp.py --values 8 > gpurun_out/bench_rsp.json 2> gpurun_out/bench_rsp.err; echo "exit $?"; cat gpurun_out/bench_rsp.json; tail -3 gpurun_out/bench_rsp.err
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo done
