set -u
mkdir -p gpurun_out
for lag in 0 1 2 3 5; do
ODF_SWEEP_ONLY=1 ODF_SWEEP_DEBUG_NCS=1 ODF_SWEEP_LAG=$lag timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:panel16_sweep -s 4 -c 1 --csv --log-file gpurun_out/r2s_lag$lag.csv python tools/sweep_time.py 524288 > /dev/null 2>&1
echo "lag=$lag: $(grep -v '^==' gpurun_out/r2s_lag$lag.csv | tail -4 | awk -F'","' '{print $(NF-2), $(NF)}' | tr '\n' ' ')"
done
