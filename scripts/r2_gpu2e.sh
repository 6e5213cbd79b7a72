#!/bin/bash
# round 2, 2-GPU visit e: suite after reverting the peer unroll; e2e staging A/B; new group features
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (2 GPUs)"
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest_gpu_2gpu_e.log 2>&1
echo "pytest exit $?"; tail -30 gpurun_out/pytest_gpu_2gpu_e.log | cut -c1-600
for mode in double single; do
  echo "== bench N=2 staging=$mode"
  if [ $mode = single ]; then export B200KV_STAGE_SINGLE=1; else unset B200KV_STAGE_SINGLE; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-config-legs > gpurun_out/bench_n2e_$mode.json 2> gpurun_out/bench_n2e_$mode.err; echo "bench exit $?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n2e_$mode.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'parity', d['parity']['ok'], 'e2e ms', d['e2e']['ms_per_step'], 'e2e parity', d['e2e']['parity']['ok'])
PY
done
unset B200KV_STAGE_SINGLE
echo done
