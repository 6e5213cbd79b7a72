#!/bin/bash
# host-buffer (e2e) path variants on one GPU + new tests
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
for mode in zc staged in_dma in_tma; do
  for bucket in 8 32; do
    B200KV_HOST_MODE=$mode B200KV_STAGE_BUCKET_MB=$bucket timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/e2e_${mode}_${bucket}.json 2> gpurun_out/e2e_${mode}_${bucket}.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/e2e_${mode}_${bucket}.json").read().strip().splitlines()[-1])
print("mode=$mode bucket=${bucket}MB e2e ms/step=%.3f GB/s=%.1f  value=%.0f launches=%s" % (d["e2e"]["ms_per_step"], d["e2e"]["value"], d["value"], d["gpu_launches"]))
PY
  done
done
echo "== train bench N=1"
timeout 900 python train_bench.py --steps 8 --warmup 3 > gpurun_out/train_n1.json 2> gpurun_out/train_n1.err; echo "exit $?"; cat gpurun_out/train_n1.json; tail -5 gpurun_out/train_n1.err
echo done
