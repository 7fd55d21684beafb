set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_pytest_gpu.log 2>&1
el "gpu suite rc=$?"; tail -4 gpurun_out/r2y_pytest_gpu.log
ODF_N=524288 ODF_MODE=resident ODF_REPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"panel16|gauss_tile2" -c 4 -f -o gpurun_out/r2y_prof python tests/ncu_target.py > gpurun_out/r2y_ncu.log 2>&1
el "ncu rc=$?"; tail -2 gpurun_out/r2y_ncu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2y_bench_c2.json 2> gpurun_out/r2y_bench_c2.err
el "bench rc=$?"; tail -3 gpurun_out/r2y_bench_c2.err; python -c "
import json; j=json.load(open('gpurun_out/r2y_bench_c2.json'))
print({k: j[k] for k in ('ms_per_step','phases_ms','gpu_launches')}, j['e2e']['ms_per_step'])
print(json.dumps(j['parity']))
print(json.dumps(j['roofline']['per_kernel']))
print(json.dumps(j['c1_pair'])[:600])"
