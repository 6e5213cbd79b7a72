#!/bin/bash
# 2-GPU visit: whole -m gpu suite, bench N=1 and N=2 (driver-style launches), reference arm, ops bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu_all.log | cut -c1-300
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "ref exit $?"; cut -c1-400 gpurun_out/bench_ref_n1.json
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 exit $?"; cut -c1-300 gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"; grep "^{" gpurun_out/bench_n2.json | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29503 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 exit $?"; grep "^{" gpurun_out/bench_ref_n2.json | cut -c1-300
timeout 300 python bench_ops.py > gpurun_out/bench_ops.json 2> gpurun_out/bench_ops.err; echo "ops exit $?"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo done
