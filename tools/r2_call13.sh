set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k triangular > gpurun_out/r2o_tri_test.log 2>&1
el "tri test rc=$?"; tail -15 gpurun_out/r2o_tri_test.log
timeout 200 python tools/apply_time.py > gpurun_out/r2o_apply_time.log 2>&1; cat gpurun_out/r2o_apply_time.log
el "apply timing done"
timeout 400 python bench.py --steps 3 --warmup 3 --no-parity --no-c1-pair --no-cpu-baseline --no-streaming-compare > gpurun_out/r2o_bench_c2.json 2> gpurun_out/r2o_bench_c2.err
el "bench rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r2o_bench_c2.json'))
print({k: j[k] for k in ('ms_per_step','phases_ms','gpu_launches')}, j['e2e']['ms_per_step'])"
