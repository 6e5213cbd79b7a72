#!/bin/bash
# Multi-GPU visit: bash scripts/gpu_multi.sh N   (N = GPUs on the box)
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n$N.txt 2>&1
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
echo "== pytest multi-GPU"
timeout 900 python -m pytest tests/test_multigpu.py tests/test_group_gpu.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_multigpu_n$N.log 2>&1
echo "pytest exit $?"; tail -30 gpurun_out/pytest_multigpu_n$N.log
echo "== bench N=1 (corrected stream)"
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "exit $?"; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    echo "== bench N=$n (torchrun)"
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "exit $?"; cat gpurun_out/bench_n$n.json; tail -5 gpurun_out/bench_n$n.err
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --impl reference --gpus $n --steps 3 --warmup 1 > gpurun_out/bench_ref_n$n.json 2> gpurun_out/bench_ref_n$n.err; echo "ref exit $?"; cut -c1-400 gpurun_out/bench_ref_n$n.json
  fi
done
echo done
