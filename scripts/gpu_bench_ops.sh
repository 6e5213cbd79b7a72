#!/bin/bash
# 1-GPU visit: optimizer-operator bench (LARS / LAMB / AdamW) + its ncu launch list, rsp ncu list
set -u
mkdir -p gpurun_out
timeout 600 python bench_ops.py > gpurun_out/bench_ops.json 2> gpurun_out/bench_ops.err; echo "exit $?"; cut -c1-700 gpurun_out/bench_ops.json; tail -5 gpurun_out/bench_ops.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 30 -c 150 --csv --log-file gpurun_out/ops_launches.csv python bench_ops.py --steps 3 --warmup 2 > gpurun_out/ops_ncu.log 2>&1; echo "ncu exit $?"
echo done
