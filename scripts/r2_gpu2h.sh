#!/bin/bash
set -u
mkdir -p gpurun_out
export B200KV_DEBUG_NVLS=1 B200KV_NVLS=1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_dbg_n2.json 2> gpurun_out/bench_dbg_n2.err; echo "exit $?"
grep "NVLS off" gpurun_out/bench_dbg_n2.err | head -5
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/bench_dbg_n2.json") if l.startswith("{")][-1]
print("main", d["ms_per_step"], d["impl_detail"]["nvls_in_switch_reduce"], d["parity"]["ok"], d["e2e"]["parity"]["ok"])
for k,v in d["configs"].items(): print("   leg", k, {a:v.get(a) for a in ("ms_per_step","push_ms","pull_ms","nvls_in_switch_reduce")}, v["parity"]["ok"])
PY
