# one-shot GPU validation of the reference-vector parity tests (round 1, last GPU minutes)
mkdir -p gpurun_out
timeout 330 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pin_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pin_pytest_gpu.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/pin_smoke.txt 2>&1; tail -3 gpurun_out/pin_smoke.txt
timeout 200 python bench.py --steps 1 --warmup 3 --e2e-steps 1 > gpurun_out/pin_bench_f64.json 2> gpurun_out/pin_bench_f64.err; tail -c 1500 gpurun_out/pin_bench_f64.json
timeout 100 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/pin_bench_reference.json 2> gpurun_out/pin_bench_reference.err; tail -c 600 gpurun_out/pin_bench_reference.json
