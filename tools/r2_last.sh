set -x
mkdir -p gpurun_out/last
O=gpurun_out/last
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest_gpu.txt
cat $O/pytest_gpu.txt | cut -c1-400
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
timeout 200 python bench.py --workload c8m16 --no-ba --steps 20 --warmup 5 > $O/bench_c8m16_s1s3.json 2> $O/bench_c8m16_s1s3.err
tail -c 400 $O/bench_c8m16_s1s3.err
python tools/ncu_summary.py gpurun_out/none 2>/dev/null | head -1
