# Round-end evidence run (one B200): full GPU test suite, the default bench line, the ncu passes the bench line cites.
mkdir -p gpurun_out
T=${1:-a}
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/freeze_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/freeze_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/bench_r02_freeze_$T.json 2> gpurun_out/bench_r02_freeze_$T.err; echo "bench rc=$?"
python -c "import bench; print(bench.csrc_hash())" > gpurun_out/headline_csrc_sha.txt
M=$(python tools/ncu_headline.py --metrics)
timeout 900 ncu --clock-control none -k regex:k_integrate_ws -s 1 -c 1 --csv --metrics $M --log-file gpurun_out/headline_metrics.csv python tools/quick_perf.py 12500 10000 double auto 1 > gpurun_out/ncu_headline.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate_ws -s 1 -c 1 -f -o gpurun_out/prof_r02_ws_final python tools/quick_perf.py 592 2000 double auto 1 > gpurun_out/ncu_ws_final.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --api-particles 0 > gpurun_out/b_under_ncu.log 2>&1
timeout 400 python tools/config_times.py 2>&1 | grep -v "Running on\|GPU device\|Platform\|Compiler\|WARNING\|^$\|NOTE" > gpurun_out/config_times_r02b.txt
timeout 200 python tools/time_split_perf.py 2>&1 | grep -v "Running on\|GPU device\|Platform\|Compiler" > gpurun_out/tsplit_perf.log
