mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "time_axis" 2>&1 | tail -40 > gpurun_out/tsplit_tests.log
timeout 200 python tools/time_split_perf.py 2>&1 | grep -v "Running on\|GPU device\|Platform\|Compiler" > gpurun_out/tsplit_perf.log
