set -u
mkdir -p gpurun_out
for pol in "0 0" "0 2" "1 0" "2 0" "0 1"; do set -- $pol; for lag in 1 2; do
ODF_SWEEP_ONLY=1 ODF_SWEEP_DEBUG_NCS=1 ODF_SWEEP_LAG=$lag ODF_SWEEP_POLICY_A=$1 ODF_SWEEP_POLICY_C=$2 timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:panel16_sweep -s 4 -c 1 --csv --log-file gpurun_out/r2v.csv python tools/sweep_time.py 500000 > /dev/null 2>&1
echo "policy_a=$1 policy_c=$2 lag=$lag: $(grep -v '^==' gpurun_out/r2v.csv | tail -2 | awk -F'","' '{print $(NF-2), $(NF)}' | tr '\n' ' ')"
done; done
