#!/bin/bash
# round 2, 8-GPU visit c: allocator change (lowest-address block), NVLS reporting; N=8 line
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_group_gpu.py tests/test_engine_gpu.py tests/test_kvstore_gpu.py -q -p no:cacheprovider 2>&1 | tail -3 | cut -c1-300
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_final2_n8.json 2> gpurun_out/bench_final2_n8.err; echo "N=8 exit $?"
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/bench_final2_n8.json") if l.startswith("{")][-1]
r=d["roofline"]
print("N=8 value %.1f ms %.4f busbw/GPU %.1f frac %.3f nvls %s parity %s e2e %.3f e2e-parity %s" % (d["value"], d["ms_per_step"], r["achieved"], r["frac"], d["impl_detail"]["nvls_in_switch_reduce"], d["parity"]["ok"], d["e2e"]["ms_per_step"], (d["e2e"]["parity"] or {}).get("ok")))
for k,v in d["configs"].items():
    print("   leg", k, {a:v.get(a) for a in ("ms_per_step","push_ms","pull_ms","nvls_in_switch_reduce")}, "parity", v["parity"]["ok"], "frac", v["roofline"].get("frac"))
PY
grep -v "^$\|\*\*\*\|OMP_NUM\|NCCL version" gpurun_out/bench_final2_n8.err | tail -3 | cut -c1-300
echo done
