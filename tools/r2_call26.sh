set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "cuda_graph" > gpurun_out/r3c_graph_test.log 2>&1
el "graph test rc=$?"; tail -12 gpurun_out/r3c_graph_test.log
for g in 0 1; do ODF_CUDA_GRAPHS=$g timeout 300 python tools/small_fit_probe.py 2>&1 | sed "s/^/graphs=$g: /" | head -4; done
el "probe done"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3c_pytest_gpu.log 2>&1
el "gpu suite rc=$?"; tail -6 gpurun_out/r3c_pytest_gpu.log
timeout 600 python bench.py --workload mb --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r3c_bench_mb.json 2> gpurun_out/r3c_bench_mb.err
el "mb rc=$?"; tail -2 gpurun_out/r3c_bench_mb.err; python -c "
import json; j=json.load(open('gpurun_out/r3c_bench_mb.json'))
print({k: j.get(k) for k in ('ms_per_step','ms_per_refit_and_scoring','gpu_launches')})"
