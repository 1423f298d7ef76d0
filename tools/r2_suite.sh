set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_suite_tests.log
cat gpurun_out/r2_suite_tests.log
MOCAP_PIPELINE=split ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2b_launches_c8m16_split.csv python bench.py --workload c8m16 --no-ba --profile --steps 2 --warmup 1 > gpurun_out/r2b_l1.log 2>&1
MOCAP_PIPELINE=split ncu --set full --clock-control none --import-source on -k regex:"k_blob_reduce_warp|k_match_triangulate" -s 6 -c 2 -o gpurun_out/r2b_c8m16_sparse python bench.py --workload c8m16 --no-ba --profile --steps 2 --warmup 1 > gpurun_out/r2b_l2.log 2>&1
timeout 600 python tools/ba_time.py > gpurun_out/r2b_ba_time.json 2> gpurun_out/r2b_ba_time.err
tail -3 gpurun_out/r2b_l1.log gpurun_out/r2b_l2.log gpurun_out/r2b_ba_time.err
