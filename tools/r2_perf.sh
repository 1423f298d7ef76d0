set -x
mkdir -p gpurun_out
timeout 600 python tools/ba_time.py > gpurun_out/r2c_ba_time.json 2> gpurun_out/r2c_ba_time.err
timeout 600 python bench.py --profile --steps 30 --warmup 5 > gpurun_out/r2c_c4m4.log 2>&1
MOCAP_PIPELINE=fused timeout 600 python bench.py --workload c8m16 --no-ba --profile --steps 20 --warmup 5 > gpurun_out/r2c_c8m16_fused.log 2>&1
MOCAP_PIPELINE=split timeout 600 python bench.py --workload c8m16 --no-ba --profile --steps 20 --warmup 5 > gpurun_out/r2c_c8m16_split.log 2>&1
timeout 600 python bench.py --workload c8m16 --profile --steps 20 --warmup 5 > gpurun_out/r2c_c8m16_ba.log 2>&1
MOCAP_PIPELINE=split ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|^void k_" -c 40 --csv --log-file gpurun_out/r2c_launches_c8m16_split.csv python bench.py --workload c8m16 --no-ba --profile --steps 2 --warmup 1 > gpurun_out/r2c_l1.log 2>&1
for f in gpurun_out/r2c_c4m4.log gpurun_out/r2c_c8m16_fused.log gpurun_out/r2c_c8m16_split.log gpurun_out/r2c_c8m16_ba.log; do tail -n 1 $f | cut -c1-1500; done
