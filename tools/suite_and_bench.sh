timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py --steps 5 --warmup 3 --no-parity --no-c1-pair --no-cpu-baseline --no-streaming-compare 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: j[k] for k in ('ms_per_step','phases_ms')}, j['e2e']['ms_per_step'], j['clocks'], {k: round(v['avg_launch_ms'],3) for k,v in j['roofline']['per_kernel'].items()})"
