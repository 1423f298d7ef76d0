set -x
mkdir -p gpurun_out
export MOCAP_PIPELINE=split
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_c8m16_split.csv python bench.py --workload c8m16 --profile --steps 2 --warmup 1 > gpurun_out/r2_l1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_blob_reduce_warp|k_match_triangulate" -s 6 -c 2 -o gpurun_out/r2_c8m16_sparse python bench.py --workload c8m16 --profile --steps 2 --warmup 1 > gpurun_out/r2_l2.log 2>&1
ls -la gpurun_out/
