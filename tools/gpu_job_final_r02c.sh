# last run of round 2 (one B200): GPU tests incl. the converter / pipelined-batch tests, default bench line with the
# ncu evidence of the same sources (profiles/r02_ncu_headline.json, csrc a5b44043c61ea81b)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/final_c_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_c_pytest_gpu.txt
( time timeout 600 python bench.py > gpurun_out/bench_r02_final_c.json 2> gpurun_out/bench_r02_final_c.err ) 2>&1 | grep real
