#!/bin/bash
# 8-GPU visit: group tests at world 4, scaling bench 2/4/8 (+ reference arm), training bench.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n8.txt 2>&1
echo "== pytest multi-GPU (world 2 and 4)"
timeout 900 python -m pytest tests/test_multigpu.py tests/test_group_gpu.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_multigpu_n8.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/pytest_multigpu_n8.log
for n in 8 4 2; do
  echo "== bench N=$n (torchrun)"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "exit $?"; cut -c1-1400 gpurun_out/bench_n$n.json; grep -v "^$\|\*\*\*\|OMP_NUM" gpurun_out/bench_n$n.err | tail -4
done
echo "== reference arm N=8"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench_ref_n8.json 2> gpurun_out/bench_ref_n8.err; echo "ref exit $?"; cut -c1-700 gpurun_out/bench_ref_n8.json
echo "== bert adam N=4"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29710 bench.py --gpus 4 --workload bert_adam --steps 20 --warmup 3 > gpurun_out/bench_bert_n4.json 2> gpurun_out/bench_bert_n4.err; echo "exit $?"; cut -c1-1400 gpurun_out/bench_bert_n4.json
echo "== train bench N=8"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29720 train_bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/train_n8.json 2> gpurun_out/train_n8.err; echo "exit $?"; cat gpurun_out/train_n8.json; grep -v "^$\|\*\*\*\|OMP_NUM" gpurun_out/train_n8.err | tail -5
echo done
