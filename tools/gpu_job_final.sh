python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python tools/config_times.py > gpurun_out/config_times_final.txt 2>&1; tail -7 gpurun_out/config_times_final.txt
python bench.py > gpurun_out/bench_final_f64.json 2> gpurun_out/bench_final_f64.err
python bench.py --dtype float --steps 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/bench_final_f32.json 2> gpurun_out/bench_final_f32.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_final_f64.json','gpurun_out/bench_final_f32.json'):
    try:
        d=json.loads(open(f).read())
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
