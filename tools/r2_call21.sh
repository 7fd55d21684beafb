set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "panel16 or resident or hi_only" > gpurun_out/r2x_panel_tests.log 2>&1
el "panel tests rc=$?"; tail -25 gpurun_out/r2x_panel_tests.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest_gpu.log 2>&1
el "gpu suite rc=$?"; tail -12 gpurun_out/r2x_pytest_gpu.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-c1-pair --no-cpu-baseline --no-streaming-compare > gpurun_out/r2x_bench_c2.json 2> gpurun_out/r2x_bench_c2.err
el "bench rc=$?"; tail -3 gpurun_out/r2x_bench_c2.err; python -c "
import json; j=json.load(open('gpurun_out/r2x_bench_c2.json'))
print({k: j[k] for k in ('ms_per_step','phases_ms','gpu_launches')}, j['e2e']['ms_per_step'])
print(json.dumps(j['parity']))
print(json.dumps(j['roofline']['per_kernel']))"
