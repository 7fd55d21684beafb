# cProfile of the minibootstrap workload (4 classes): where the HOST time of the reference's regime goes
set -u
mkdir -p gpurun_out
timeout 600 python -m cProfile -o gpurun_out/mb.prof bench.py --workload mb --n 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/mb_prof.json 2> gpurun_out/mb_prof.err
python - <<'PY'
import pstats
p = pstats.Stats('gpurun_out/mb.prof')
p.sort_stats('cumulative').print_stats(45)
p.sort_stats('tottime').print_stats(30)
PY
