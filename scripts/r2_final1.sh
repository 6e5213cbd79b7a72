#!/bin/bash
# round 2, final 1-GPU visit: whole -m gpu suite, smoke, default bench line, ncu launch list + full captures
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu_final.log | cut -c1-300
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_final.log
echo "== bench (driver command)"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['parity']['ok'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['roofline']['frac'])
print(json.dumps(d['frontends'])[:1200])
for k,v in d['configs'].items(): print(k, {a:v.get(a) for a in ('ms_per_step','push_ms','pull_ms')}, v['parity']['ok'], v['roofline'].get('frac'))
print('cpu', d['cpu_baseline'])
PY
tail -3 gpurun_out/bench_final_n1.err
echo "== reference arm"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final_ref_n1.json 2> gpurun_out/bench_final_ref_n1.err; cut -c1-300 gpurun_out/bench_final_ref_n1.json
echo "== ncu launch list (bench, no legs)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-config-legs > gpurun_out/ncu_list_final.log 2>&1; echo "exit $?"
echo "== ncu full: dense_fused (SGD, N=1)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_fused -s 4 -c 2 -f -o gpurun_out/r02_prof_dense python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-config-legs > gpurun_out/ncu_full_dense.log 2>&1; echo "exit $?"
echo "== ncu full: rsp_sum_bm + retain"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rsp_sum_bm|retain_kernel" -s 6 -c 2 -f -o gpurun_out/r02_prof_rsp python bench.py --workload rsp --steps 3 > gpurun_out/ncu_full_rsp.log 2>&1; echo "exit $?"
ls -la gpurun_out/*.ncu-rep
echo done
