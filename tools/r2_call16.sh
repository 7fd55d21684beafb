set -u
mkdir -p gpurun_out
ODF_SWEEP_ONLY=1 timeout 200 python tools/sweep_time.py 2>&1 | tail -1
ODF_SWEEP_ONLY=1 ODF_SWEEP_DEBUG_NCS=1 timeout 200 python tools/sweep_time.py 2>&1 | tail -1 | sed "s/^/dbg_ncs=1: /"
ODF_SWEEP_ONLY=1 ODF_SWEEP_DEBUG_NCS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:panel16_sweep -s 4 -c 1 -f -o gpurun_out/r2r_sweep_dbg python tools/sweep_time.py > gpurun_out/r2r_ncu.log 2>&1
tail -2 gpurun_out/r2r_ncu.log
