set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
MOCAP_TEST_PHASED=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests.log
for mode in fused split phased; do
  MOCAP_PIPELINE=$mode timeout 300 python bench.py --workload c8m16 --profile --steps 30 --warmup 5 > gpurun_out/r2_c8m16_$mode.log 2>&1
  MOCAP_PIPELINE=$mode timeout 300 python bench.py --workload c4m4 --profile --steps 30 --warmup 5 > gpurun_out/r2_c4m4_$mode.log 2>&1
done
timeout 300 python bench.py --workload c8m16 --profile --steps 30 --warmup 5 > gpurun_out/r2_c8m16_auto.log 2>&1
cat gpurun_out/r2_tests.log; tail -2 gpurun_out/r2_c*.log
