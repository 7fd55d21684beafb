set -u
N=${NGPU:-2}
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
if [ "$N" = "2" ]; then
timeout 600 python -m pytest tests -m gpu -x -q -k "nccl or two_rank or mgpu or ranks" > gpurun_out/r3b_mgpu_tests.log 2>&1
el "mgpu tests rc=$?"; tail -4 gpurun_out/r3b_mgpu_tests.log
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-c1-pair --no-streaming-compare > gpurun_out/r3b_bench_c2_${N}gpu.json 2> gpurun_out/r3b_bench_c2_${N}gpu.err
el "c2 x$N rc=$?"; tail -2 gpurun_out/r3b_bench_c2_${N}gpu.err; python -c "
import json; j=json.load(open('gpurun_out/r3b_bench_c2_${N}gpu.json'))
print({k: j.get(k) for k in ('n_gpus','ms_per_step','phases_ms')}, j['e2e'].get('ms_per_step'))
print(json.dumps(j.get('parity'))[:900])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r3b_bench_c5_${N}gpu.json 2> gpurun_out/r3b_bench_c5_${N}gpu.err
el "c5 x$N rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r3b_bench_c5_${N}gpu.json'))
print({k: j.get(k) for k in ('n_gpus','ms_per_step','rois_per_s')}, j['e2e'])"
