set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "native_cg" > gpurun_out/r3d_native_test.log 2>&1
el "native test rc=$?"; tail -12 gpurun_out/r3d_native_test.log
for g in 0 1; do ODF_NATIVE_CG=$g timeout 300 python tools/small_fit_probe.py 2>&1 | sed "s/^/native=$g: /" | head -3; done
el "probe done"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3d_pytest_gpu.log 2>&1
el "gpu suite rc=$?"; tail -6 gpurun_out/r3d_pytest_gpu.log
timeout 600 python bench.py --workload mb --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r3d_bench_mb.json 2> gpurun_out/r3d_bench_mb.err
el "mb rc=$?"; tail -2 gpurun_out/r3d_bench_mb.err; python -c "
import json; j=json.load(open('gpurun_out/r3d_bench_mb.json'))
print({k: j.get(k) for k in ('ms_per_step','ms_per_refit_and_scoring','gpu_launches')})"
timeout 600 python bench.py --steps 3 --warmup 3 --no-parity --no-c1-pair --no-cpu-baseline --no-streaming-compare > gpurun_out/r3d_bench_c2.json 2> gpurun_out/r3d_bench_c2.err
el "c2 rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r3d_bench_c2.json'))
print({k: j[k] for k in ('ms_per_step','phases_ms','gpu_launches')}, j['e2e']['ms_per_step'])"
