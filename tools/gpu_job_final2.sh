mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x > gpurun_out/final2_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final2_pytest_gpu.txt
