mkdir -p gpurun_out
for ph in spread; do python tools/quick_perf.py 592 10000 double $ph 2 2>/dev/null | tee -a gpurun_out/spread_perf.txt; done
