set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2g_suite_tests.log
cat gpurun_out/r2g_suite_tests.log
timeout 600 python bench.py --profile --steps 30 --warmup 5 > gpurun_out/r2g_c4m4.log 2>&1
timeout 600 python bench.py --workload c8m16 --profile --steps 20 --warmup 5 > gpurun_out/r2g_c8m16_ba.log 2>&1
timeout 600 python tools/ba_time.py > gpurun_out/r2g_ba_time.json 2> gpurun_out/r2g_ba_time.err
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2g_smoke.log 2>&1
tail -n 1 gpurun_out/r2g_c4m4.log gpurun_out/r2g_c8m16_ba.log gpurun_out/r2g_smoke.log | cut -c1-900
