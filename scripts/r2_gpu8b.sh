#!/bin/bash
# round 2, final 8-GPU visit: the driver's bench command at N=8 and N=4 with the final tree
set -u
mkdir -p gpurun_out
for n in 8 4; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_final_n$n.json 2> gpurun_out/bench_final_n$n.err; echo "N=$n exit $?"
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/bench_final_n$n.json") if l.startswith("{")][-1]
r=d["roofline"]
print("N=$n value %.1f ms %.4f busbw/GPU %.1f frac %.3f nvls %s parity %s e2e %.3f e2e-parity %s" % (d["value"], d["ms_per_step"], r["achieved"], r["frac"], d["impl_detail"]["nvls_in_switch_reduce"], d["parity"]["ok"], d["e2e"]["ms_per_step"], (d["e2e"]["parity"] or {}).get("ok")))
for k,v in d["configs"].items():
    print("   leg", k, {a:v.get(a) for a in ("ms_per_step","push_ms","pull_ms")}, "parity", v["parity"]["ok"], "frac", v["roofline"].get("frac"))
print("   frontends", json.dumps(d["frontends"]))
PY
  grep -v "^$\|\*\*\*\|OMP_NUM\|NCCL version" gpurun_out/bench_final_n$n.err | tail -3 | cut -c1-300
done
echo done
