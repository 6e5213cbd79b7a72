#!/bin/bash
# round 2, 2-GPU visit g: gate kernel for the start barrier, A/B; group suites
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_group_gpu.py tests/test_multigpu.py tests/test_nccl_fallback_gpu.py tests/test_group_rsp_plain_gpu.py tests/test_bucketing_gpu.py -q -p no:cacheprovider 2>&1 | tail -4 | cut -c1-300
for gate in 1 0; do
  B200KV_GATE=$gate timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29517+gate)) bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2g_gate$gate.json 2> gpurun_out/bench_n2g_gate$gate.err; echo "gate=$gate exit $?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n2g_gate$gate.json').read().strip().splitlines()[-1])
print("gate=$gate", {k:d[k] for k in ('value','ms_per_step')}, 'parity', d['parity']['ok'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['parity']['ok'])
for k,v in d['configs'].items(): print("   ", k, {a:v.get(a) for a in ('ms_per_step','push_ms','pull_ms')}, v['parity']['ok'])
print("   ", json.dumps(d['frontends']))
PY
done
echo done
