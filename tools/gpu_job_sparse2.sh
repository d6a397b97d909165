mkdir -p gpurun_out
rm -f gpurun_out/sparse_ab.log
for lib in ""; do
    SYNCHRAD_B200_LIB=$lib timeout 200 python tools/c34_perf.py c3 c4 c4d 2>&1 | grep "integrate_ms" >> gpurun_out/sparse_ab.log
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_reference_vectors_gpu.py -m gpu -x -q --timeout 300 2>&1 | tail -8 > gpurun_out/sparse_tests.log
