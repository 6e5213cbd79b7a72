#!/bin/bash
# 8-GPU visit: multi-GPU / group parity (worlds 2 and 4, IPC + NVLS), driver-style bench at N=8 and
# N=4 (+ reference arm at N=8), row_sparse bench over 8 GPUs, ResNet-50 training bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py tests/test_group_gpu.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_multigpu_n8.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_multigpu_n8.log | cut -c1-300
for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench n$n exit $?"
  python - <<PY
import json
for l in open("gpurun_out/bench_n$n.json"):
    if l.startswith("{"):
        d=json.loads(l); print("N=$n value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "busbw/GPU", round(d["roofline"]["achieved"],1), "nvls", d["config"]["nvls_in_switch_reduce"], "e2e ms", round(d["e2e"]["ms_per_step"],3))
PY
  grep -v "^$\|\*\*\*\|OMP_NUM\|NCCL version" gpurun_out/bench_n$n.err | tail -3 | cut -c1-300
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench_ref_n8.json 2> gpurun_out/bench_ref_n8.err; echo "ref exit $?"; grep "^{" gpurun_out/bench_ref_n8.json | cut -c1-260
timeout 300 python bench_rsp.py --values 8 > gpurun_out/bench_rsp_n8.json 2> gpurun_out/bench_rsp_n8.err; echo "rsp exit $?"; cut -c1-420 gpurun_out/bench_rsp_n8.json; tail -2 gpurun_out/bench_rsp_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29720 train_bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/train_n8.json 2> gpurun_out/train_n8.err; echo "train exit $?"; cut -c1-600 gpurun_out/train_n8.json; grep -v "^$\|\*\*\*\|OMP_NUM" gpurun_out/train_n8.err | tail -3 | cut -c1-300
echo done
