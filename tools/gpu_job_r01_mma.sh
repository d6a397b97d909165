set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
# full-set capture, small workload
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 1 -c 1 -o gpurun_out/prof_r01_v6_mma python tools/quick_perf.py 592 2000 > gpurun_out/ncu_v6.log 2>&1
# full-size launch metrics (few metrics -> few replays)
timeout 900 ncu --clock-control none -k regex:k_integrate -s 1 -c 1 --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_tensor.sum --log-file gpurun_out/full_size_mma.csv python tools/quick_perf.py 12500 10000 double auto 1 > gpurun_out/ncu_full.log 2>&1
# launch list of the bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v6.csv python bench.py --steps 1 --warmup 1 > gpurun_out/b_under_ncu.log 2>&1
python bench.py > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err
tail -c 3000 gpurun_out/bench_v7.json
