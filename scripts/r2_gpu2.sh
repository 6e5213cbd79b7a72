#!/bin/bash
# round 2, 2-GPU visit: whole -m gpu suite (multi-GPU + rank-per-GPU tests included), the NVLink
# access-pattern probe, the N=2 bench line (parity + legs)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
echo "== pytest -m gpu (2 GPUs)"
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest_gpu_2gpu.log 2>&1
echo "pytest exit $?"; tail -30 gpurun_out/pytest_gpu_2gpu.log
echo "== peer probe N=2"
timeout 300 tools/bin/peer_probe 2 > gpurun_out/peer_probe_n2.txt 2>&1; echo "probe exit $?"; cat gpurun_out/peer_probe_n2.txt
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench exit $?"; cut -c1-5000 gpurun_out/bench_n2.json; tail -8 gpurun_out/bench_n2.err
echo "== bench N=1 (frontends / rsp after the bitmap union)"
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1b.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['parity']['ok'], d['e2e']['ms_per_step'])
print(json.dumps(d['frontends'])[:1500])
print(json.dumps(d['configs']['rsp'])[:1500])
PY
tail -5 gpurun_out/bench_n1b.err
echo done
