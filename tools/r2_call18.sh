set -u
mkdir -p gpurun_out
for n in 524160 500000; do for lag in 0 2; do
ODF_SWEEP_ONLY=1 ODF_SWEEP_DEBUG_NCS=1 ODF_SWEEP_LAG=$lag timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:panel16_sweep -s 4 -c 1 --csv --log-file gpurun_out/r2t_${n}_lag${lag}.csv python tools/sweep_time.py $n > /dev/null 2>&1
echo "n=$n lag=$lag: $(grep -v '^==' gpurun_out/r2t_${n}_lag${lag}.csv | tail -4 | awk -F'","' '{print $(NF-2), $(NF)}' | tr '\n' ' ')"
done; done
ODF_SWEEP_ONLY=1 ODF_SWEEP_DEBUG_NCS=1 timeout 200 python tools/sweep_time.py 524160 2>&1 | tail -1
ODF_SWEEP_ONLY=1 timeout 200 python tools/sweep_time.py 524160 2>&1 | tail -1
