timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_reference_vectors_gpu.py -m gpu -x -q --timeout 200 -k "c2 or near or edge or random or golden or C2" 2>&1 | tail -3
for tw in 2 4 8; do SRB_FORCE_TW=$tw MODE=near GRID=128,256,32 LSCREEN=1e5 timeout 100 python tools/quick_perf.py 24 3329 double auto 1 2>/dev/null | tail -1; done
MODE=near GRID=128,256,32 LSCREEN=1e5 timeout 100 python tools/quick_perf.py 24 3329 double direct 1 2>/dev/null | tail -1
for tw in 4 8; do SRB_FORCE_TW=$tw timeout 100 python tools/quick_perf.py 148 10000 double drec 1 2>/dev/null | tail -1; done
