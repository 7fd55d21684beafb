set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 600 python tools/accuracy_probe.py > gpurun_out/r2g_accuracy_probe.log 2>&1
el "accuracy probe rc=$?"; cat gpurun_out/r2g_accuracy_probe.log
