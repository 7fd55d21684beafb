set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -k "chunked_upload or host" > gpurun_out/r2z_upload_tests.log 2>&1
el "upload tests rc=$?"; tail -15 gpurun_out/r2z_upload_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-parity --no-c1-pair --no-cpu-baseline --no-streaming-compare > gpurun_out/r2z_bench_c2.json 2> gpurun_out/r2z_bench_c2.err
el "bench rc=$?"; tail -3 gpurun_out/r2z_bench_c2.err; python -c "
import json; j=json.load(open('gpurun_out/r2z_bench_c2.json'))
print({k: j[k] for k in ('ms_per_step','phases_ms','gpu_launches')}, j['e2e'])"
