set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2l_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-parity --no-c1-pair --no-cpu-baseline --no-streaming-compare > gpurun_out/r2l_launches_bench.log 2>&1
el "launch list rc=$?"
python tools/ncu_launch_summary.py gpurun_out/r2l_launches.csv > gpurun_out/r2l_launches_summary.md; head -40 gpurun_out/r2l_launches_summary.md
