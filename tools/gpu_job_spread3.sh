mkdir -p gpurun_out
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_integrate -c 1 -f -o gpurun_out/spread_prof python tools/quick_perf.py 148 2000 double spread 1 > gpurun_out/spread_ncu.log 2>&1
tail -3 gpurun_out/spread_ncu.log; ls -la gpurun_out/spread_prof.ncu-rep
