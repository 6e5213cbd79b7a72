#!/bin/bash
# 1-GPU visit: multi-tensor optimizer operator parity, row_sparse parity + bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_rowsparse_gpu.py tests/test_kvstore_gpu.py -m gpu -q --maxfail=15 -p no:cacheprovider > gpurun_out/pytest_ops.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/pytest_ops.log | cut -c1-300
timeout 300 python bench_rsp.py --values 8 --gpus 1 > gpurun_out/bench_rsp_g1.json 2> gpurun_out/bench_rsp_g1.err; cat gpurun_out/bench_rsp_g1.json; tail -3 gpurun_out/bench_rsp_g1.err
echo done
