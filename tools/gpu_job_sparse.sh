mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_reference_vectors_gpu.py -m gpu -x -q --timeout 300 2>&1 | tail -8 > gpurun_out/sparse_tests.log
timeout 300 python tools/c3_kernels.py 2>&1 | grep -v "Running on\|GPU device\|Platform\|Compiler" > gpurun_out/sparse_c3.log
timeout 400 python tools/config_times.py 2>&1 | grep -v "Running on\|GPU device\|Platform\|Compiler" > gpurun_out/sparse_config_times.log
