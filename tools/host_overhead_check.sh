set -u
mkdir -p gpurun_out
timeout 300 python tools/small_fit_probe.py 2>&1 | head -4
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --workload mb --steps 1 --warmup 1 > gpurun_out/r3p_bench_mb.json 2> gpurun_out/r3p_bench_mb.err; python -c "
import json; j=json.load(open('gpurun_out/r3p_bench_mb.json'))
print({k: j.get(k) for k in ('ms_per_step','ms_per_refit_and_scoring','gpu_launches')}, j['cpu_baseline']['surviving_sets_identical'])"
timeout 400 python bench.py --steps 3 --warmup 3 --no-parity --no-c1-pair --no-cpu-baseline --no-streaming-compare 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: j[k] for k in ('ms_per_step','phases_ms')}, j['e2e']['ms_per_step'])"
