mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 1 -c 1 -f -o gpurun_out/prof_r02_c3d python tools/c34_perf.py c3 > gpurun_out/ncu_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 1 -c 1 -f -o gpurun_out/prof_r02_c4b python tools/c34_perf.py c4 > gpurun_out/ncu_c4.log 2>&1
