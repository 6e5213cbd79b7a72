#!/bin/bash
# 2-GPU visit: e2e (host buffers) with the copy-lane pipeline: N=1 zc vs pipe, N=2 group mode buckets
set -u
mkdir -p gpurun_out
for mode in zc pipe; do
  for mb in 8 16; do
    [ "$mode" = zc ] && [ "$mb" = 16 ] && continue
    B200KV_HOST_MODE=$mode B200KV_STAGE_BUCKET_MB=$mb timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/e2e_n1_${mode}_$mb.json 2> gpurun_out/e2e_n1_${mode}_$mb.err
    python - <<PY
import json
for l in open("gpurun_out/e2e_n1_${mode}_$mb.json"):
    if l.startswith("{"):
        d=json.loads(l); print("N=1 $mode bucket=$mb  e2e ms", round(d["e2e"]["ms_per_step"],3), " value ms", round(d["ms_per_step"],4))
PY
    tail -2 gpurun_out/e2e_n1_${mode}_$mb.err | cut -c1-300
  done
done
for mb in 16 8 1000; do
  B200KV_GROUP_BUCKET_MB=$mb timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600+mb%100)) bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/e2e_n2_b$mb.json 2> gpurun_out/e2e_n2_b$mb.err
  python - <<PY
import json
for l in open("gpurun_out/e2e_n2_b$mb.json"):
    if l.startswith("{"):
        d=json.loads(l); print("N=2 group bucket=$mb  e2e ms", round(d["e2e"]["ms_per_step"],3), " value ms", round(d["ms_per_step"],4))
PY
  grep -v "^$\|\*\*\*\|OMP_NUM\|NCCL version" gpurun_out/e2e_n2_b$mb.err | tail -3 | cut -c1-300
done
timeout 600 python -m pytest tests/test_group_gpu.py tests/test_kvstore_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
echo done
