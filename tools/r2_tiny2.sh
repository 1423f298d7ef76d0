timeout 55 python -m pytest tests/test_parity_gpu.py -q -x -k "ba_ or config3 or tracks or locate or error" 2>&1 | tail -3 | cut -c1-300
timeout 25 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
