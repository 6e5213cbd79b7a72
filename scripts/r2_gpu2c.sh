#!/bin/bash
# round 2, 2-GPU visit c: full suite after the optimizer front-end split / scan + retain rewrite /
# ld-st host packing; N=2 and N=1 bench lines
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (2 GPUs)"
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest_gpu_2gpu_c.log 2>&1
echo "pytest exit $?"; tail -30 gpurun_out/pytest_gpu_2gpu_c.log
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2c.json 2> gpurun_out/bench_n2c.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2c.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'parity', d['parity']['ok'], 'e2e ms', d['e2e']['ms_per_step'], 'e2e parity', d['e2e']['parity']['ok'])
print('frontends', json.dumps(d['frontends']))
for k,v in d['configs'].items(): print(k, json.dumps(v)[:900])
PY
tail -5 gpurun_out/bench_n2c.err
echo "== bench N=2 e2e with TMA host packing (A/B)"
B200KV_PACK_MODE=tma timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 --no-config-legs > gpurun_out/bench_n2c_tma.json 2> gpurun_out/bench_n2c_tma.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2c_tma.json').read().strip().splitlines()[-1])
print('tma pack: e2e ms', d['e2e']['ms_per_step'])
PY
echo "== bench N=1"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1c.json 2> gpurun_out/bench_n1c.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1c.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['parity']['ok'], d['e2e']['ms_per_step'])
print(json.dumps(d['configs']['rsp'])[:1200])
PY
tail -5 gpurun_out/bench_n1c.err
echo "== rsp bench line + launch list"
timeout 600 python bench.py --workload rsp --steps 20 > gpurun_out/bench_rsp_n1c.json 2> gpurun_out/bench_rsp_n1c.err; cut -c1-1500 gpurun_out/bench_rsp_n1c.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/launches_rsp3.csv python bench.py --workload rsp --steps 3 > gpurun_out/ncu_list_rsp3.log 2>&1; echo "exit $?"
echo done
