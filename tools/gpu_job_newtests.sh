timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 500 2>&1 | tail -15
timeout 100 python tools/quick_perf.py 592 10000 double auto 2 2>/dev/null | tail -1
