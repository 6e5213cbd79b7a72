#!/bin/bash
# round 2, 1-GPU visit #1: full -m gpu suite (incl. the BASELINE-size parity tests), smoke, the new
# bench line (parity + config legs), rsp bench line, ncu launch list of the bench
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -30 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cut -c1-6000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
echo "== bench rsp"
timeout 600 python bench.py --workload rsp --steps 20 > gpurun_out/bench_rsp_n1.json 2> gpurun_out/bench_rsp_n1.err; echo "exit $?"; cut -c1-3000 gpurun_out/bench_rsp_n1.json; tail -5 gpurun_out/bench_rsp_n1.err
echo "== ncu launch list (bench, no legs)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-config-legs > gpurun_out/ncu_list.log 2>&1; echo "exit $?"
echo "== ncu launch list (rsp)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches_rsp.csv python bench.py --workload rsp --steps 3 > gpurun_out/ncu_list_rsp.log 2>&1; echo "exit $?"
echo done
