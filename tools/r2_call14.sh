set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k one_pass > gpurun_out/r2q_sweep_test.log 2>&1
el "sweep test rc=$?"; tail -25 gpurun_out/r2q_sweep_test.log
timeout 300 python tools/sweep_time.py > gpurun_out/r2q_sweep_time.log 2>&1; cat gpurun_out/r2q_sweep_time.log | tail -5
el "sweep timing done"
for lag in 1 3; do ODF_SWEEP_ONLY=1 ODF_SWEEP_LAG=$lag timeout 200 python tools/sweep_time.py 2>&1 | tail -1 | sed "s/^/lag=$lag: /"; done
for pa in 0 2; do ODF_SWEEP_ONLY=1 ODF_SWEEP_POLICY_A=$pa timeout 200 python tools/sweep_time.py 2>&1 | tail -1 | sed "s/^/policy_a=$pa: /"; done
for pc in 0 1; do ODF_SWEEP_ONLY=1 ODF_SWEEP_POLICY_C=$pc timeout 200 python tools/sweep_time.py 2>&1 | tail -1 | sed "s/^/policy_c=$pc: /"; done
el "variants done"
