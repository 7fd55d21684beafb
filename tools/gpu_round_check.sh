#!/bin/bash
# One gpurun call: resident-mode parity tests -> bench (both sweep modes) -> full GPU suite -> ncu launch list.
# Every step is bounded by its own timeout; logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv,noheader
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k resident > gpurun_out/v7_resident_tests.log 2>&1
RES=$?
tail -5 gpurun_out/v7_resident_tests.log
el "resident tests rc=$RES"
if [ $RES -ne 0 ]; then
  export ODF_SWEEP_MODE=panel16
  DESEL="--deselect tests/test_gpu_parity.py::test_resident_sweeper_matches_oracle --deselect tests/test_gpu_parity.py::test_resident_fit_matches_streaming_fit_and_oracle"
else
  DESEL=""
fi
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/v7_bench_c2.json 2> gpurun_out/v7_bench_c2.err
el "bench rc=$?"
tail -c 3000 gpurun_out/v7_bench_c2.json
tail -3 gpurun_out/v7_bench_c2.err
timeout ${SUITE_TIMEOUT:-480} python -m pytest tests -m gpu -x -q $DESEL > gpurun_out/v7_pytest_gpu.log 2>&1
el "gpu suite rc=$?"
tail -6 gpurun_out/v7_pytest_gpu.log
if [ "${DO_NCU:-1}" = "1" ]; then
  timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/v7_launches.csv \
    python bench.py --steps 1 --warmup 3 --n 131072 --no-e2e --no-cpu-baseline --no-streaming-compare > gpurun_out/v7_launches_bench.log 2>&1
  el "ncu launch list rc=$?"
  tail -c 600 gpurun_out/v7_launches_bench.log
fi
el done
