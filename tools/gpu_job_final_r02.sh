# last evidence refresh of round 2 (one B200): smoke, reference arm, default bench line, ncu metrics pass keyed by the source hash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err ) 2>&1 | grep real
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference_arm.json 2> gpurun_out/bench_r02_reference_arm.err; echo "reference arm rc=$?"
python -c "import bench; print(bench.csrc_hash())" > gpurun_out/headline_csrc_sha.txt
M=$(python tools/ncu_headline.py --metrics)
timeout 900 ncu --clock-control none -k regex:k_integrate_ws -s 1 -c 1 --csv --metrics $M --log-file gpurun_out/headline_metrics.csv python tools/quick_perf.py 12500 10000 double auto 1 > gpurun_out/ncu_headline.log 2>&1
timeout 300 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -2
