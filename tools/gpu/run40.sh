python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or reproducible or shards" 2>&1 | tail -2
