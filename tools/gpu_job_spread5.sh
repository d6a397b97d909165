mkdir -p gpurun_out
SYNCHRAD_B200_LIB=$PWD/variants/libspread_nomain.so python tools/quick_perf.py 592 10000 double spread 2 2>/dev/null | tee -a gpurun_out/spread_perf2.txt
SRB_FORCE_TW=8 python tools/quick_perf.py 592 10000 double direct 1 2>/dev/null | tee -a gpurun_out/spread_perf2.txt
