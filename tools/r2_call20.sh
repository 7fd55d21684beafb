set -u
mkdir -p gpurun_out
for ov in 0 1; do
ODF_OVERLAP_RHS=$ov timeout 400 python bench.py --steps 3 --warmup 3 --no-parity --no-c1-pair --no-cpu-baseline --no-streaming-compare > gpurun_out/r2w_bench_ov$ov.json 2> gpurun_out/r2w_bench_ov$ov.err
echo "overlap=$ov rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r2w_bench_ov$ov.json'))
print({k: j[k] for k in ('ms_per_step','phases_ms','gpu_launches')}, j['e2e']['ms_per_step'])"
done
