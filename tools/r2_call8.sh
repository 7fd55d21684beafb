set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest_gpu.log 2>&1
el "gpu suite rc=$?"; tail -25 gpurun_out/r2h_pytest_gpu.log
timeout 300 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/r2h_bench_c5_1gpu.json 2> gpurun_out/r2h_bench_c5.err
el "c5 rc=$?"; tail -c 1500 gpurun_out/r2h_bench_c5_1gpu.json; tail -3 gpurun_out/r2h_bench_c5.err
timeout 600 python bench.py --workload mb --steps 1 --warmup 1 > gpurun_out/r2h_bench_mb.json 2> gpurun_out/r2h_bench_mb.err
el "mb rc=$?"; tail -c 2500 gpurun_out/r2h_bench_mb.json; tail -5 gpurun_out/r2h_bench_mb.err
timeout 600 python bench.py --workload c4 --steps 2 --warmup 3 --no-c1-pair > gpurun_out/r2h_bench_c4.json 2> gpurun_out/r2h_bench_c4.err
el "c4 rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r2h_bench_c4.json'))
print({k: j[k] for k in ('ms_per_step','phases_ms','sweep_mode','rls','parity')})"; tail -3 gpurun_out/r2h_bench_c4.err
