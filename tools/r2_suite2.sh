set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2f_suite_tests.log
cat gpurun_out/r2f_suite_tests.log
timeout 300 python tests/stage_bench.py colour > gpurun_out/r2f_colour_fused.json 2>&1
MOCAP_PIPELINE=split timeout 300 python tests/stage_bench.py colour > gpurun_out/r2f_colour_split.json 2>&1
timeout 600 python tools/ba_time.py > gpurun_out/r2f_ba_time.json 2> gpurun_out/r2f_ba_time.err
cat gpurun_out/r2f_colour_fused.json gpurun_out/r2f_colour_split.json | tail -4
