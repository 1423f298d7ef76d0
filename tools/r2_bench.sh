set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c4m4.json 2> gpurun_out/r2_bench_c4m4.err; tail -3 gpurun_out/r2_bench_c4m4.err
timeout 900 python bench.py --workload c8m16 --steps 20 --warmup 5 > gpurun_out/r2_bench_c8m16.json 2> gpurun_out/r2_bench_c8m16.err; tail -3 gpurun_out/r2_bench_c8m16.err
timeout 600 python bench.py --workload c8m16 --no-ba --profile --steps 20 --warmup 5 > gpurun_out/r2_bench_c8m16_noba.json 2>&1
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
cat gpurun_out/r2_bench_c4m4.json gpurun_out/r2_bench_c8m16.json gpurun_out/r2_bench_c8m16_noba.json gpurun_out/r2_bench_ref.json | cut -c1-3000
