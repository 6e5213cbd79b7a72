#!/bin/bash
# round 2, 2-GPU visit b: barrier timeline probe, PCIe pipeline probe, e2e host modes, rsp launch list
set -u
mkdir -p gpurun_out
echo "== engine ABI tests"
timeout 600 python -m pytest tests/test_engine_abi_gpu.py tests/test_bucketing_gpu.py tests/test_kvstore_gpu.py -q -p no:cacheprovider 2>&1 | tail -5
echo "== peer probe N=2 (with barriers)"
timeout 300 tools/bin/peer_probe 2 > gpurun_out/peer_probe_n2b.txt 2>&1; echo "probe exit $?"; head -12 gpurun_out/peer_probe_n2b.txt
echo "== pcie pipeline probe"
timeout 300 python tools/pcie_pipeline_probe.py > gpurun_out/pcie_pipeline_probe.txt 2>&1; cat gpurun_out/pcie_pipeline_probe.txt
echo "== e2e host modes N=1"
for cfg in "zc 8" "staged 4" "staged 8" "staged 16" "pipe 8" "pipe 16"; do
  set -- $cfg
  B200KV_HOST_MODE=$1 B200KV_STAGE_BUCKET_MB=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config-legs > gpurun_out/e2e_$1_$2.json 2> gpurun_out/e2e_$1_$2.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/e2e_$1_$2.json").read().strip().splitlines()[-1])
    print("mode=$1 bucket=$2 MB: e2e %.3f ms/step parity=%s" % (d["e2e"]["ms_per_step"], d["e2e"]["parity"]["ok"]))
except Exception as e:
    print("mode=$1 bucket=$2 failed", e)
PY
done
echo "== ncu launch list (rsp)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/launches_rsp2.csv python bench.py --workload rsp --steps 3 > gpurun_out/ncu_list_rsp2.log 2>&1; echo "exit $?"
echo done
