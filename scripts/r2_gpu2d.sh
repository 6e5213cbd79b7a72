#!/bin/bash
# round 2, 2-GPU visit d: NCCL fallback tests, unrolled dense kernel, double-buffered host staging,
# shared-memory mailbox
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (2 GPUs)"
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest_gpu_2gpu_d.log 2>&1
echo "pytest exit $?"; tail -40 gpurun_out/pytest_gpu_2gpu_d.log | cut -c1-400
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2d.json 2> gpurun_out/bench_n2d.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2d.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'parity', d['parity']['ok'], 'e2e ms', d['e2e']['ms_per_step'], 'e2e parity', d['e2e']['parity']['ok'])
for k,v in d['configs'].items(): print(k, {a:v.get(a) for a in ('ms_per_step','push_ms','pull_ms')}, v['parity']['ok'])
PY
tail -5 gpurun_out/bench_n2d.err
echo "== bench N=1"
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1d.json 2> gpurun_out/bench_n1d.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1d.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['parity']['ok'], d['e2e']['ms_per_step'])
for k,v in d['configs'].items(): print(k, {a:v.get(a) for a in ('ms_per_step','push_ms','pull_ms')}, v['parity']['ok'], v['roofline'].get('frac'))
PY
tail -5 gpurun_out/bench_n1d.err
echo done
