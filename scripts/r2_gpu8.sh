#!/bin/bash
# round 2, 8-GPU visit: cross-GPU parity suites, NVLink probes at N=4 / N=8, driver-style bench lines at
# N=8 (NVLS unroll A/B, bit-exact mode) and N=4 (BERT + Adam = BASELINE configs[3]), row_sparse over 8
# ranks and over 8 GPUs of one process, ResNet-50 training (configs[2]), CPU reference arm at N=8
set -u
mkdir -p gpurun_out
run_bench() {  # name nproc extra-env... -- args
  local name=$1 n=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name exit $?"
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/$name.json") if l.startswith("{")][-1]
    r=d.get("roofline",{})
    print("$name: value %.1f ms %.4f busbw/GPU %s nvls %s parity %s e2e %s" % (d["value"], d["ms_per_step"], r.get("achieved"), d.get("impl_detail",{}).get("nvls_in_switch_reduce"), (d.get("parity") or {}).get("ok"), (d.get("e2e") or {}).get("ms_per_step")))
    for k,v in (d.get("configs") or {}).items():
        print("   leg", k, {a:v.get(a) for a in ("ms_per_step","push_ms","pull_ms")}, "parity", (v.get("parity") or {}).get("ok"), "frac", (v.get("roofline") or {}).get("frac"))
    print("   frontends", json.dumps(d.get("frontends"))[:300])
except Exception as e:
    print("$name: no line", e)
PY
  grep -v "^$\|\*\*\*\|OMP_NUM\|NCCL version" gpurun_out/$name.err | tail -4 | cut -c1-400
}
echo "== cross-GPU parity suites"
timeout 1200 python -m pytest tests/test_multigpu.py tests/test_group_gpu.py tests/test_nccl_fallback_gpu.py tests/test_group_rsp_plain_gpu.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_multigpu_n8.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_multigpu_n8.log | cut -c1-400
echo "== NVLink probes"
timeout 200 tools/bin/peer_probe 4 > gpurun_out/peer_probe_n4.txt 2>&1; head -9 gpurun_out/peer_probe_n4.txt
timeout 200 tools/bin/peer_probe 8 > gpurun_out/peer_probe_n8.txt 2>&1; head -9 gpurun_out/peer_probe_n8.txt
echo "== bench lines"
run_bench bench_n8 8 --steps 20 --warmup 5
B200KV_NVLS_UNROLL=1 run_bench bench_n8_nvls_u1 8 --steps 20 --warmup 5 --no-config-legs
B200KV_NVLS_UNROLL=2 run_bench bench_n8_nvls_u2 8 --steps 20 --warmup 5 --no-config-legs
B200KV_NVLS=0 run_bench bench_n8_nvls0 8 --steps 20 --warmup 5 --no-config-legs
B200KV_DMA_MIN_KB=0 run_bench bench_n8_packonly 8 --steps 10 --warmup 3 --no-config-legs
run_bench bench_n4 4 --steps 20 --warmup 5
B200KV_NVLS=1 run_bench bench_n4_nvls1 4 --steps 20 --warmup 5 --no-config-legs
run_bench bench_rsp_n8 8 --workload rsp --steps 20
echo "== reference arm N=8"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29931 bench.py --impl reference --gpus 8 --steps 5 --warmup 1 > gpurun_out/bench_ref_n8.json 2> gpurun_out/bench_ref_n8.err; echo "ref exit $?"; grep "^{" gpurun_out/bench_ref_n8.json | cut -c1-300
echo "== row_sparse, one process driving 8 GPUs"
timeout 300 python bench_rsp.py --values 8 > gpurun_out/bench_rsp_1proc_n8.json 2> gpurun_out/bench_rsp_1proc_n8.err; echo "rsp exit $?"; cut -c1-500 gpurun_out/bench_rsp_1proc_n8.json; tail -2 gpurun_out/bench_rsp_1proc_n8.err
echo "== ResNet-50 training, 8 GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29941 train_bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/train_n8.json 2> gpurun_out/train_n8.err; echo "train exit $?"; cut -c1-700 gpurun_out/train_n8.json; grep -v "^$\|\*\*\*\|OMP_NUM" gpurun_out/train_n8.err | tail -3 | cut -c1-300
echo done
