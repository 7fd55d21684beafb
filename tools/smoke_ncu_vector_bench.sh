# One gpurun call: smoke(), ncu --set full of the vector kernels (profiles/r2_ncu_vector_kernels.md), the DEFAULT bench run timed by the wall clock.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG:-r3j}_smoke.log 2>&1
el "smoke rc=$?"; tail -3 gpurun_out/${TAG:-r3j}_smoke.log
if [ "${DO_NCU:-1}" = "1" ]; then
ODF_N=524288 ODF_MODE=resident ODF_REPS=1 timeout 900 ncu --set full --clock-control none -k regex:"rownorm|split_kernel|finish_rows|split_w16|tri_apply|cg_elem|colreduce" -c 14 -f -o gpurun_out/${TAG:-r3j}_vec python tests/ncu_target.py > gpurun_out/${TAG:-r3j}_ncu.log 2>&1
el "ncu vec rc=$?"; tail -2 gpurun_out/${TAG:-r3j}_ncu.log
fi
python bench.py > gpurun_out/${TAG:-r3j}_bench_default.json 2> gpurun_out/${TAG:-r3j}_bench_default.err
el "default bench rc=$? (wall clock of the plain 'python bench.py' = the difference to the previous mark)"
python - <<PY
import json
j = json.load(open('gpurun_out/${TAG:-r3j}_bench_default.json'))
print({k: j[k] for k in ('ms_per_step', 'phases_ms', 'gpu_launches', 'streaming_fit_s')}, j['e2e']['ms_per_step'])
print(json.dumps(j['parity']['sub_fit'])[:700])
print(json.dumps(j['cpu_baseline'])[:300])
print(json.dumps(j['c1_pair']['batched'])[:400])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG:-r3j}_bench_reference.json 2> gpurun_out/${TAG:-r3j}_bench_reference.err
el "reference arm rc=$?"; head -c 600 gpurun_out/${TAG:-r3j}_bench_reference.json
