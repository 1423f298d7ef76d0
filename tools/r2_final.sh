set -x
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_c4m4.json 2> $O/bench_c4m4.err
timeout 900 python bench.py --workload c8m16 --steps 20 --warmup 5 > $O/bench_c8m16_config3.json 2> $O/bench_c8m16_config3.err
timeout 600 python bench.py --workload c8m16 --no-ba --profile --steps 20 --warmup 5 > $O/c8m16_s1s3.json 2>&1
timeout 900 python tests/stage_bench.py > $O/stage_bench.json 2> $O/stage_bench.err
ncu --set full --clock-control none --import-source on -k regex:k_pipeline_fused -s 4 -c 1 -o $O/c4m4_fused python bench.py --profile --steps 2 --warmup 1 > $O/n1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_ba_solve|k_threshold_segments|k_blob_reduce_warp|k_match_triangulate" -s 12 -c 4 -o $O/c8m16_config3 python bench.py --workload c8m16 --profile --steps 2 --warmup 1 > $O/n2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|^void k_" -s 27 -c 40 --csv --log-file $O/launches_c8m16_config3.csv python bench.py --workload c8m16 --profile --steps 2 --warmup 1 > $O/n3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|^void k_" -s 12 -c 12 --csv --log-file $O/launches_c4m4.csv python bench.py --profile --steps 2 --warmup 1 > $O/n4.log 2>&1
ls -la $O | tail -20
tail -n 2 $O/*.err | cut -c1-300
