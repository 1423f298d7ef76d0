set -x
mkdir -p gpurun_out
timeout 300 python tools/matcher_probe.py 2>&1 | tail -20
MOCAP_MATCH_CHUNK=64 timeout 400 python -m pytest tests/test_parity_gpu.py -q -x -k "match or pipeline or golden or fuzz" 2>&1 | tail -5
