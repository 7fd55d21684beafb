set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 600 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-streaming-compare --no-c1-pair > gpurun_out/r3a_bench_c4.json 2> gpurun_out/r3a_bench_c4.err
el "c4 rc=$?"; tail -2 gpurun_out/r3a_bench_c4.err; python -c "
import json; j=json.load(open('gpurun_out/r3a_bench_c4.json'))
print({k: j.get(k) for k in ('ms_per_step','phases_ms','sweep_mode','rls')}, j['e2e'].get('ms_per_step'))
print(json.dumps(j.get('parity'))[:700])
print(json.dumps(j['roofline'].get('other_kernel'))[:500])"
timeout 300 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r3a_bench_c5_1gpu.json 2> gpurun_out/r3a_bench_c5.err
el "c5 rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r3a_bench_c5_1gpu.json'))
print({k: j.get(k) for k in ('ms_per_step','rois_per_s')}, j['e2e'])"
timeout 600 python bench.py --workload mb --steps 1 --warmup 1 > gpurun_out/r3a_bench_mb.json 2> gpurun_out/r3a_bench_mb.err
el "mb rc=$?"; tail -2 gpurun_out/r3a_bench_mb.err; python -c "
import json; j=json.load(open('gpurun_out/r3a_bench_mb.json'))
print({k: j.get(k) for k in ('ms_per_step','ms_per_refit_and_scoring','gpu_launches')}, j['cpu_baseline'])"
