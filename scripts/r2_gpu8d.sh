#!/bin/bash
set -u
mkdir -p gpurun_out
export B200KV_DEBUG_NVLS=1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29609 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_dbg_n8.json 2> gpurun_out/bench_dbg_n8.err; echo "exit $?"
grep "NVLS off" gpurun_out/bench_dbg_n8.err | head -5
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/bench_dbg_n8.json") if l.startswith("{")][-1]
print("main", d["ms_per_step"], d["impl_detail"]["nvls_in_switch_reduce"])
for k,v in d["configs"].items(): print("   leg", k, {a:v.get(a) for a in ("ms_per_step","push_ms","pull_ms","nvls_in_switch_reduce")})
PY
echo "== standalone bert"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29610 bench.py --gpus 8 --steps 10 --warmup 3 --workload bert_adam --no-config-legs > gpurun_out/bench_dbg_bert_n8.json 2> gpurun_out/bench_dbg_bert_n8.err; echo "exit $?"
grep "NVLS off" gpurun_out/bench_dbg_bert_n8.err | head -5
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/bench_dbg_bert_n8.json") if l.startswith("{")][-1]
print("bert standalone", d["ms_per_step"], d["impl_detail"]["nvls_in_switch_reduce"], d["parity"]["ok"], "e2e", d["e2e"]["ms_per_step"])
PY
echo done
