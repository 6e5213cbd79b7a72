#!/bin/bash
# round 2, 1-GPU visit 10: queue with move semantics (tests), hybrid host mode A/B
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bucketing_gpu.py tests/test_kvstore_gpu.py tests/test_trainer_patterns_gpu.py tests/test_engine_abi_gpu.py -q -p no:cacheprovider 2>&1 | tail -4
for mode in zc hybrid; do
  for mb in 8 16 32; do
    [ $mode = zc ] && [ $mb != 8 ] && continue
    B200KV_HOST_MODE=$mode B200KV_STAGE_BUCKET_MB=$mb timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config-legs > gpurun_out/e2e2_${mode}_$mb.json 2> gpurun_out/e2e2_${mode}_$mb.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/e2e2_${mode}_$mb.json").read().strip().splitlines()[-1])
    print("mode=$mode bucket=$mb MB: e2e %.3f ms/step parity=%s ; per-key %.4f, measure-pattern %.4f (grouped %.4f)" % (d["e2e"]["ms_per_step"], d["e2e"]["parity"]["ok"], d["frontends"]["per_key_pushpull_priority_minus_i"]["ms_per_step"], d["frontends"]["per_key_push_all_then_pull_all"]["ms_per_step"], d["ms_per_step"]))
except Exception as e:
    print("mode=$mode bucket=$mb failed", e); print(open("gpurun_out/e2e2_${mode}_$mb.err").read()[-600:])
PY
  done
done
B200KV_HOST_MODE=hybrid B200KV_STAGE_BUCKET_MB=16 B200KV_DMA_MIN_KB=4096 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config-legs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('hybrid 16MB min4MB: e2e %.3f' % d['e2e']['ms_per_step'])"
echo done
