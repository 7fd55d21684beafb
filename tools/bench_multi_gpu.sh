set -u
N=${NGPU:-8}
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
run() { # name port args...
  name=$1; port=$2; shift 2
  timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" > gpurun_out/r3h_${name}_${N}gpu.out 2> gpurun_out/r3h_${name}_${N}gpu.err
  rc=$?
  grep '^{' gpurun_out/r3h_${name}_${N}gpu.out | tail -1 > gpurun_out/r3h_${name}_${N}gpu.json
  el "$name x$N rc=$rc"
  python -c "
import json; j=json.load(open('gpurun_out/r3h_${name}_${N}gpu.json'))
print({k: j.get(k) for k in ('n_gpus','ms_per_step','phases_ms','rois_per_s','sweep_mode')}, (j.get('e2e') or {}).get('ms_per_step'))
p=j.get('parity')
if p: print({k: p.get(k) for k in ('alpha_max_abs_diff_across_ranks','rel_score_err','argmax_flips_outside_band')}, {k: (p.get('sub_fit') or {}).get(k) for k in ('rel_score_err','residual_gpu','residual_oracle')})
" 2>&1 | tail -3
}
run c2 29521 --steps 5 --warmup 3 --no-cpu-baseline --no-c1-pair --no-streaming-compare
run c5 29522 --workload c5 --steps 5 --warmup 3 --no-cpu-baseline
if [ "$N" = "8" ] && [ "${DO_C3:-1}" = "1" ]; then run c3 29523 --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-c1-pair --no-streaming-compare --no-parity; fi
