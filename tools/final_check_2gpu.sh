set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/fin_mgpu_test.log 2>&1; echo "mgpu test rc=$?"; tail -3 gpurun_out/fin_mgpu_test.log
NGPU=2 bash tools/bench_multi_gpu.sh
