mkdir -p gpurun_out
timeout 300 python tools/quick_perf.py 12500 10000 double auto 1 2>/dev/null | tail -1
timeout 300 python tools/quick_perf.py 12500 10000 float auto 1 2>/dev/null | tail -1
python -c "import bench; print(bench.csrc_hash())" > gpurun_out/headline_csrc_sha.txt
M=$(python tools/ncu_headline.py --metrics)
timeout 900 ncu --clock-control none -k regex:k_integrate_ws -s 1 -c 1 --csv --metrics $M --log-file gpurun_out/headline_metrics.csv python tools/quick_perf.py 12500 10000 double auto 1 > gpurun_out/ncu_headline.log 2>&1
grep "dram__bytes\|gpu__time" gpurun_out/headline_metrics.csv | cut -d, -f13-15
