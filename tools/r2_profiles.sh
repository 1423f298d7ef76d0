set -x
mkdir -p gpurun_out
timeout 600 python tools/ba_time.py > gpurun_out/r2e_ba_time.json 2> gpurun_out/r2e_ba_time.err
timeout 900 python bench.py --workload c8m16 --steps 20 --warmup 5 > gpurun_out/r2e_bench_c8m16.json 2> gpurun_out/r2e_bench_c8m16.err
# ncu: fused kernel on config 2 (traffic), BA kernel, stream kernel on config 3
ncu --set full --clock-control none --import-source on -k regex:k_pipeline_fused -s 4 -c 1 -o gpurun_out/r2e_c4m4_fused python bench.py --profile --steps 2 --warmup 1 > gpurun_out/r2e_n1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_ba_solve|k_threshold_segments" -s 9 -c 2 -o gpurun_out/r2e_c8m16_ba python bench.py --workload c8m16 --profile --steps 2 --warmup 1 > gpurun_out/r2e_n2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|^void k_" -s 27 -c 40 --csv --log-file gpurun_out/r2e_launches_c8m16_ba.csv python bench.py --workload c8m16 --profile --steps 2 --warmup 1 > gpurun_out/r2e_n3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|^void k_" -s 12 -c 12 --csv --log-file gpurun_out/r2e_launches_c4m4.csv python bench.py --profile --steps 2 --warmup 1 > gpurun_out/r2e_n4.log 2>&1
ls -la gpurun_out/ | tail -12
tail -n 2 gpurun_out/r2e_n1.log gpurun_out/r2e_n2.log | cut -c1-300
