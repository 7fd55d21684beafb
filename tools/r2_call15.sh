set -u
mkdir -p gpurun_out
for ncs in 1 4; do for lag in 2 4; do ODF_SWEEP_ONLY=1 ODF_SWEEP_DEBUG_NCS=$ncs ODF_SWEEP_LAG=$lag timeout 200 python tools/sweep_time.py 2>&1 | tail -1 | sed "s/^/dbg_ncs=$ncs lag=$lag: /"; done; done
ODF_SWEEP_ONLY=1 ODF_SWEEP_DEBUG_NCS=1 ODF_SWEEP_LAG=3 ODF_SWEEP_POLICY_A=2 timeout 200 python tools/sweep_time.py 2>&1 | tail -1 | sed "s/^/dbg_ncs=1 lag=3 pa=2: /"
ODF_SWEEP_ONLY=1 ODF_SWEEP_DEBUG_NCS=1 ODF_SWEEP_LAG=3 ODF_SWEEP_POLICY_A=0 ODF_SWEEP_POLICY_C=0 timeout 200 python tools/sweep_time.py 2>&1 | tail -1 | sed "s/^/dbg_ncs=1 lag=3 pa=0 pc=0: /"
