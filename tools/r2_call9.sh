set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest_gpu.log 2>&1
el "gpu suite rc=$?"; tail -12 gpurun_out/r2i_pytest_gpu.log
timeout 400 python tools/accuracy_probe.py > gpurun_out/r2i_accuracy_c2like.log 2>&1
el "probe c2-like rc=$?"; cat gpurun_out/r2i_accuracy_c2like.log
ODF_PRECOND_APPLY=cublas timeout 400 python tools/accuracy_probe.py 2>&1 | grep "default\|CPU fp32" > gpurun_out/r2i_accuracy_c2like_cublas.log
el "probe c2-like cublas"; cat gpurun_out/r2i_accuracy_c2like_cublas.log
timeout 400 python tools/accuracy_probe.py 200000 4000 256 15 50 1e-3 > gpurun_out/r2i_accuracy_c4like.log 2>&1
el "probe c4-like rc=$?"; cat gpurun_out/r2i_accuracy_c4like.log
ODF_PRECOND_APPLY=cublas timeout 400 python tools/accuracy_probe.py 200000 4000 256 15 50 1e-3 2>&1 | grep "default\|CPU fp32" > gpurun_out/r2i_accuracy_c4like_cublas.log
cat gpurun_out/r2i_accuracy_c4like_cublas.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-parity --no-c1-pair --no-cpu-baseline --no-streaming-compare > gpurun_out/r2i_bench_c2.json 2> gpurun_out/r2i_bench_c2.err
el "bench rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r2i_bench_c2.json'))
print({k: j[k] for k in ('ms_per_step','phases_ms','gpu_launches')}, j['e2e']['ms_per_step'])"
