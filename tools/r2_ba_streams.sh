mkdir -p gpurun_out/bas
O=gpurun_out/bas
timeout 45 python bench.py --workload c8m16 --ba-streams 4 --profile --steps 20 --warmup 5 > $O/profile_k4.json 2> $O/profile_k4.err
cut -c1-900 $O/profile_k4.json; tail -n 3 $O/profile_k4.err | cut -c1-300
timeout 80 python bench.py --workload c8m16 --ba-streams 4 --steps 20 --warmup 5 > $O/bench_c8m16_config3_k4.json 2> $O/bench_k4.err
cut -c1-200 $O/bench_c8m16_config3_k4.json; tail -n 3 $O/bench_k4.err | cut -c1-300
timeout 45 python bench.py --workload c8m16 --ba-streams 2 --profile --steps 20 --warmup 5 > $O/profile_k2.json 2> $O/profile_k2.err
cut -c1-300 $O/profile_k2.json
