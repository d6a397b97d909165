mkdir -p gpurun_out
python tools/quick_perf.py 592 10000 double spread 2 2>/dev/null | tee -a gpurun_out/spread_perf.txt
timeout 100 python -m pytest tests/test_gpu_parity.py -q -x -k "gridding" 2>&1 | tail -3
