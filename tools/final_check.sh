# Final state of the round on one B200: build check, smoke(), pytest -m gpu, the plain bench run and the reference arm.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG:-fin}_smoke.log 2>&1
el "smoke rc=$?"; tail -1 gpurun_out/${TAG:-fin}_smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG:-fin}_pytest_gpu.log 2>&1
el "pytest -m gpu rc=$?"; tail -3 gpurun_out/${TAG:-fin}_pytest_gpu.log
python bench.py > gpurun_out/${TAG:-fin}_bench_default.json 2> gpurun_out/${TAG:-fin}_bench_default.err
el "python bench.py rc=$?"
python - <<PY
import json
j = json.load(open('gpurun_out/${TAG:-fin}_bench_default.json'))
print({k: j[k] for k in ('ms_per_step', 'phases_ms', 'gpu_launches', 'streaming_fit_s', 'executed_tensor_tflops')}, 'e2e', j['e2e']['ms_per_step'], j['clocks'])
print({k: j['roofline'][k] for k in ('kernel', 'achieved', 'peak', 'frac', 'traffic', 'share_of_step', 'algorithmic_bytes_per_launch')})
p = j['parity']; print({k: p[k] for k in ('rel_score_err', 'rel_score_err_rms', 'argmax_flips_outside_band')}, {k: p['sub_fit'][k] for k in ('rel_score_err', 'rel_score_err_rms', 'residual_gpu', 'residual_oracle')}, p['sub_fit']['cpu_fp32_port_vs_fp64_oracle'])
print(json.dumps(j['c1_pair']['batched'])[:300])
PY
timeout 300 python bench.py --impl reference > gpurun_out/${TAG:-fin}_bench_reference.json 2> gpurun_out/${TAG:-fin}_bench_reference.err
el "reference arm rc=$?"; head -c 300 gpurun_out/${TAG:-fin}_bench_reference.json; echo
