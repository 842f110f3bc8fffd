(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_gpu" 2>&1 | tail -6)
timeout 300 python scripts/mg_bench.py lap7 128 2 2 0 2>&1 | tail -1
