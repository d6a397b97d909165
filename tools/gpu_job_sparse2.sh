mkdir -p gpurun_out
rm -f gpurun_out/sparse_ab.log
SYNCHRAD_B200_LIB= timeout 200 python tools/c34_perf.py c3 c4 c4d 2>&1 | grep "integrate_ms" >> gpurun_out/sparse_ab.log
timeout 100 python tools/quick_perf.py 592 10000 double recur 2 2>/dev/null | tail -1 >> gpurun_out/sparse_ab.log
timeout 100 python tools/quick_perf.py 592 10000 float recur 2 2>/dev/null | tail -1 >> gpurun_out/sparse_ab.log
timeout 100 python tools/quick_perf.py 592 10000 double direct 1 2>/dev/null | tail -1 >> gpurun_out/sparse_ab.log
MODE=near GRID=128,256,32 LSCREEN=1e5 timeout 100 python tools/quick_perf.py 24 3329 double auto 1 2>/dev/null | tail -1 >> gpurun_out/sparse_ab.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_reference_vectors_gpu.py -m gpu -x -q --timeout 300 2>&1 | tail -8 > gpurun_out/sparse_tests.log
