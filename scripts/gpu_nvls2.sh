#!/bin/bash
# 2-GPU visit: NVLS (multimem) mode of the fused kernel vs peer-load mode
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_group_gpu.py tests/test_compression_gpu.py -m gpu -q -s --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_nvls.log 2>&1; echo "pytest exit $?"; grep -v "^$" gpurun_out/pytest_nvls.log | tail -25
for nvls in 0 1; do
  echo "== bench N=2 NVLS=$nvls"
  B200KV_NVLS=$nvls timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29800+nvls)) bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2_nvls$nvls.json 2> gpurun_out/bench_n2_nvls$nvls.err; echo "exit $?"; cut -c1-200 gpurun_out/bench_n2_nvls$nvls.json; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n2_nvls$nvls.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "busbw/GPU", d["roofline"]["achieved"], "nvls", d["config"].get("nvls_in_switch_reduce"))
except Exception as e: print("ERR", e)
PY
  grep -v "^$\|\*\*\*\|OMP_NUM" gpurun_out/bench_n2_nvls$nvls.err | tail -6
done
echo "== rsp ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/rsp_launches.csv python bench_rsp.py --values 8 --steps 3 > gpurun_out/rsp_ncu.log 2>&1; echo "exit $?"
echo done
