set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "triangular" > gpurun_out/r3f_tri_test.log 2>&1
el "tri test rc=$?"; tail -8 gpurun_out/r3f_tri_test.log
timeout 200 python tools/apply_time.py 2>&1 | grep "T=1 \|M=10000 T=30"
timeout 300 python tools/small_fit_probe.py 2>&1 | head -3
el "probe done"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3f_pytest_gpu.log 2>&1
el "gpu suite rc=$?"; tail -4 gpurun_out/r3f_pytest_gpu.log
timeout 600 python bench.py --workload mb --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r3f_bench_mb.json 2> gpurun_out/r3f_bench_mb.err
el "mb rc=$?"; tail -2 gpurun_out/r3f_bench_mb.err; python -c "
import json; j=json.load(open('gpurun_out/r3f_bench_mb.json'))
print({k: j.get(k) for k in ('ms_per_step','ms_per_refit_and_scoring','gpu_launches')})"
