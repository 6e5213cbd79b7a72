#!/bin/bash
# round 2, 2-GPU visit f: N=2 after restoring the register budgets; e2e direction diagnostics
set -u
mkdir -p gpurun_out
export B200KV_E2E_DIAG=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2f.json 2> gpurun_out/bench_n2f.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2f.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'parity', d['parity']['ok'], 'e2e', json.dumps(d['e2e'])[:900])
for k,v in d['configs'].items(): print(k, {a:v.get(a) for a in ('ms_per_step','push_ms','pull_ms')}, v['parity']['ok'])
print(json.dumps(d['frontends']))
PY
tail -4 gpurun_out/bench_n2f.err
unset B200KV_E2E_DIAG
echo "== bench N=1"
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1f.json 2> gpurun_out/bench_n1f.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1f.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['parity']['ok'], d['e2e']['ms_per_step'])
for k,v in d['configs'].items(): print(k, {a:v.get(a) for a in ('ms_per_step','push_ms','pull_ms')}, v['parity']['ok'], v['roofline'].get('frac'))
PY
echo done
