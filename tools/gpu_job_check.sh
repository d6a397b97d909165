mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x > gpurun_out/check_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/check_pytest_gpu.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
