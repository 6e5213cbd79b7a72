#!/bin/bash
# 8-GPU visit for the NVLS mode: group parity at world 4 (both modes), bench N=8 and N=4 with the
# in-switch reduce off / on.
set -u
mkdir -p gpurun_out
echo "== pytest group (world 2 and 4, IPC + NVLS)"
timeout 900 python -m pytest tests/test_group_gpu.py -m gpu -q -s --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_nvls_n8.log 2>&1
echo "pytest exit $?"; grep -v "^$" gpurun_out/pytest_nvls_n8.log | tail -8 | cut -c1-600
for n in 8 4; do
  for nvls in 0 1; do
    echo "== bench N=$n NVLS=$nvls"
    B200KV_NVLS=$nvls timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n+10*nvls)) bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/bench_n${n}_nvls$nvls.json 2> gpurun_out/bench_n${n}_nvls$nvls.err; echo "exit $?"
    python - <<EOF
import json
for l in open("gpurun_out/bench_n${n}_nvls$nvls.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("ms/step", d["ms_per_step"], "busbw/GPU", d["roofline"]["achieved"], "nvls", d["config"]["nvls_in_switch_reduce"], "e2e", d["e2e"]["ms_per_step"])
EOF
    grep -v "^$\|\*\*\*\|OMP_NUM\|NCCL version" gpurun_out/bench_n${n}_nvls$nvls.err | tail -4 | cut -c1-400
  done
done
echo "== bert adam N=8 NVLS=1"
B200KV_NVLS=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29710 bench.py --gpus 8 --workload bert_adam --steps 20 --warmup 3 > gpurun_out/bench_bert_n8_nvls1.json 2> gpurun_out/bench_bert_n8_nvls1.err; echo "exit $?"; cut -c1-300 gpurun_out/bench_bert_n8_nvls1.json
echo done
