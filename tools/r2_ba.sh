set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "ba_ or tracks or config3" 2>&1 | tail -25 > gpurun_out/r2_ba_tests.log
cat gpurun_out/r2_ba_tests.log
timeout 600 python tools/ba_time.py > gpurun_out/r2_ba_time.json 2> gpurun_out/r2_ba_time.err
tail -5 gpurun_out/r2_ba_time.err
cat gpurun_out/r2_ba_time.json | head -150
