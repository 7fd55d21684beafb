set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3i_smoke.log 2>&1
el "smoke rc=$?"; tail -3 gpurun_out/r3i_smoke.log
ODF_N=524288 ODF_MODE=resident ODF_REPS=1 timeout 900 ncu --set full --clock-control none -k regex:"rownorm|split_kernel|finish_rows|split_w16|tri_apply|cg_elem|colreduce" -c 14 -f -o gpurun_out/r3i_vec python tests/ncu_target.py > gpurun_out/r3i_ncu.log 2>&1
el "ncu vec rc=$?"; tail -2 gpurun_out/r3i_ncu.log
python bench.py > gpurun_out/r3i_bench_default.json 2> gpurun_out/r3i_bench_default.err
import json; j=json.load(open('gpurun_out/r3i_bench_default.json'))
print({k: j[k] for k in ('ms_per_step','phases_ms','gpu_launches','streaming_fit_s')}, j['e2e']['ms_per_step'])
print(json.dumps(j['parity']['sub_fit'])[:600])
print(json.dumps(j['cpu_baseline'])[:300])
print(json.dumps(j['c1_pair']['batched'])[:300])"
