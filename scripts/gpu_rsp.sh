#!/bin/bash
# row_sparse visit (2 GPUs): parity tests, bench at 2 and 1 GPU, ncu launch list of the 1-GPU bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rowsparse_gpu.py tests/test_multigpu.py tests/test_trainer_patterns_gpu.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_rsp.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_rsp.log | cut -c1-400
timeout 600 python bench_rsp.py --values 8 > gpurun_out/bench_rsp_g2.json 2> gpurun_out/bench_rsp_g2.err; echo "exit $?"; cat gpurun_out/bench_rsp_g2.json; tail -3 gpurun_out/bench_rsp_g2.err
timeout 300 python bench_rsp.py --values 8 --gpus 1 > gpurun_out/bench_rsp_g1.json 2> gpurun_out/bench_rsp_g1.err; cat gpurun_out/bench_rsp_g1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file gpurun_out/rsp_launches.csv python bench_rsp.py --values 8 --gpus 1 --steps 4 > gpurun_out/rsp_ncu.log 2>&1; echo "ncu exit $?"
echo done
