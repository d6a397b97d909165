# Strong scaling of the fixed C3 / C4 BASELINE configs over 1, 2, 4 GPUs of one box + the N > 1 parity tests (needs --gpus 4)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x --timeout 500 > gpurun_out/strong_pytest_mgpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/strong_pytest_mgpu.txt
rm -f gpurun_out/strong_scaling_r02.jsonl
for w in c3 c4; do
  timeout 300 python bench.py --scaling strong --workload $w --steps 3 --warmup 2 >> gpurun_out/strong_scaling_r02.jsonl 2>> gpurun_out/strong.err
  for n in 2 4; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --scaling strong --workload $w --steps 3 --warmup 2 >> gpurun_out/strong_scaling_r02.jsonl 2>> gpurun_out/strong.err
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/strong_scaling_r02.jsonl'):
    try: d = json.loads(l)
    except Exception: continue
    print(d['config']['workload'][:28], 'N=%d' % d['n_gpus'], 'ms_per_step=%.2f' % d['ms_per_step'], 'kernel_ms=%.2f' % d['kernel_ms_max_over_ranks'], 'updates/s=%.3e' % d['value'])
PY
