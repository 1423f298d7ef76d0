mkdir -p gpurun_out/tiny
timeout 40 python -m pytest tests/test_parity_gpu.py -q -x -k "ba_grid_setter" 2>&1 | tail -4 | cut -c1-300
timeout 45 python bench.py --workload c8m16 --profile --steps 20 --warmup 5 > gpurun_out/tiny/profile_default.json 2> gpurun_out/tiny/profile_default.err
cut -c1-260 gpurun_out/tiny/profile_default.json; tail -n 2 gpurun_out/tiny/profile_default.err | cut -c1-300
