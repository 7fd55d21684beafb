#!/bin/bash
# One gpurun call: resident-mode parity tests -> bench (both sweep modes) -> full GPU suite -> ncu launch list.
# Knobs: TAG (file prefix), DO_NCU=0, DO_NCU_FULL=0, DO_BRINGUP=0, SUITE_TIMEOUT.
# Every step is bounded by its own timeout; logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv,noheader
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "resident or panel16_mmv" > gpurun_out/${TAG:-v9}_resident_tests.log 2>&1
RES=$?
tail -5 gpurun_out/${TAG:-v9}_resident_tests.log
el "resident tests rc=$RES"
if [ $RES -ne 0 ]; then
  export ODF_SWEEP_MODE=panel16
  DESEL="--deselect tests/test_gpu_parity.py::test_resident_sweeper_matches_oracle --deselect tests/test_gpu_parity.py::test_resident_fit_matches_streaming_fit_and_oracle --deselect tests/test_gpu_parity.py::test_panel16_mmv_kernel_matches_fp64_product"
else
  DESEL=""
fi
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG:-v9}_bench_c2.json 2> gpurun_out/${TAG:-v9}_bench_c2.err
el "bench rc=$?"
tail -c 3000 gpurun_out/${TAG:-v9}_bench_c2.json
tail -3 gpurun_out/${TAG:-v9}_bench_c2.err
timeout ${SUITE_TIMEOUT:-480} python -m pytest tests -m gpu -x -q $DESEL > gpurun_out/${TAG:-v9}_pytest_gpu.log 2>&1
el "gpu suite rc=$?"
tail -6 gpurun_out/${TAG:-v9}_pytest_gpu.log
if [ "${DO_NCU:-1}" = "1" ]; then
  timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG:-v9}_launches.csv \
    python bench.py --steps 1 --warmup 3 --n 131072 --no-e2e --no-cpu-baseline --no-streaming-compare > gpurun_out/${TAG:-v9}_launches_bench.log 2>&1
  el "ncu launch list rc=$?"
  tail -c 600 gpurun_out/${TAG:-v9}_launches_bench.log
fi
if [ "${DO_NCU_FULL:-1}" = "1" ]; then
  ODF_MODE=resident ODF_REPS=2 timeout 240 ncu --set full --clock-control none --import-source on -k regex:panel16 -c 3 \
    -f -o gpurun_out/${TAG:-v9}_prof python tests/ncu_target.py > gpurun_out/${TAG:-v9}_prof.log 2>&1
  el "ncu --set full rc=$?"
  tail -3 gpurun_out/${TAG:-v9}_prof.log
  ls -la gpurun_out/${TAG:-v9}_prof.ncu-rep
fi
if [ "${DO_BRINGUP:-1}" = "1" ]; then
  timeout 200 python tools/p16v_bringup.py > gpurun_out/${TAG:-v9}_p16v_bringup.log 2>&1
  el "p16v bring-up rc=$?"
  tail -8 gpurun_out/${TAG:-v9}_p16v_bringup.log
fi
el done
