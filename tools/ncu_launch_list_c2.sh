# ncu launch list (gpu__time_duration) of the bench command: the per-kernel shares of a C2 fit (profiles/r2_ncu_launches_c2_final*.{csv,md})
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${TAG:-r3k}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-parity --no-c1-pair --no-cpu-baseline --no-streaming-compare --no-e2e > gpurun_out/${TAG:-r3k}_launches_bench.log 2>&1
echo "rc=$?"
python tools/ncu_launch_summary.py gpurun_out/${TAG:-r3k}_launches.csv > gpurun_out/${TAG:-r3k}_launches_summary.md; head -32 gpurun_out/${TAG:-r3k}_launches_summary.md
