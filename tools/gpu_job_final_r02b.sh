# evidence refresh with the final sources of round 2 (one B200): smoke, GPU tests, ncu metrics pass keyed by the source hash,
# default bench line, reference arm, launch list of one bench step
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/final_b_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_b_pytest_gpu.txt
python -c "import bench; print(bench.csrc_hash())" > gpurun_out/headline_csrc_sha.txt
M=$(python tools/ncu_headline.py --metrics)
timeout 300 ncu --clock-control none -k regex:k_integrate_ws -s 1 -c 1 --csv --metrics $M --log-file gpurun_out/headline_metrics.csv python tools/quick_perf.py 12500 10000 double auto 1 > gpurun_out/ncu_headline.log 2>&1; echo "ncu rc=$?"
( time timeout 600 python bench.py > gpurun_out/bench_r02_final_b.json 2> gpurun_out/bench_r02_final_b.err ) 2>&1 | grep real
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference_arm_b.json 2> gpurun_out/bench_r02_reference_arm_b.err; echo "reference arm rc=$?"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r02b.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --api-particles 0 --no-fp32 > gpurun_out/b_under_ncu_b.log 2>&1; echo "launch list rc=$?"
