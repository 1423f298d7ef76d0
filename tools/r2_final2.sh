set -x
mkdir -p gpurun_out/final2
O=gpurun_out/final2
# 1. full captures of the dominant kernels of both workloads, traffic stamps from them (the benches below read the stamps)
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_pipeline_fused -s 4 -c 1 -o $O/c4m4_fused python bench.py --profile --steps 2 --warmup 1 > $O/n1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_threshold_segments|k_blob_reduce_warp|k_match_triangulate|k_match_chunks" -s 8 -c 4 -o $O/c8m16_s1s3 python bench.py --workload c8m16 --no-ba --profile --steps 2 --warmup 1 > $O/n2.log 2>&1
ncu -i $O/c4m4_fused.ncu-rep --page raw --csv > $O/ncu_full_r02b_c4m4_fused.csv
ncu -i $O/c8m16_s1s3.ncu-rep --page raw --csv > $O/ncu_full_r02b_c8m16_s1s3.csv
python tools/ncu_traffic.py $O/c4m4_fused.ncu-rep k_pipeline_fused c4m4 profiles/ncu_full_r02b_c4m4_fused.csv
python tools/ncu_traffic.py $O/c8m16_s1s3.ncu-rep k_threshold_segments c8m16 profiles/ncu_full_r02b_c8m16_s1s3.csv
cp profiles/traffic_c4m4.json profiles/traffic_c8m16.json $O/
# 2. the GPU test suite
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest_gpu.txt
cat $O/pytest_gpu.txt
# 3. bench lines
timeout 300 python bench.py --workload c8m16 --steps 20 --warmup 5 > $O/bench_c8m16_config3.json 2> $O/bench_c8m16_config3.err
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_c4m4.json 2> $O/bench_c4m4.err
timeout 300 python bench.py --workload c8m16 --no-ba --profile --steps 20 --warmup 5 > $O/c8m16_s1s3.json 2>&1
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
# 4. launch list of the config-3 step
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|^void k_" -s 30 -c 44 --csv --log-file $O/launches_c8m16_config3.csv python bench.py --workload c8m16 --profile --steps 2 --warmup 1 > $O/n3.log 2>&1
rm -f $O/c4m4_fused.ncu-rep
ls -la $O | tail -24
tail -n 2 $O/*.err | cut -c1-300
tail -c 600 $O/c8m16_s1s3.json
