"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed
keys; under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                          '--steps', '1', '--warmup', '0', '--track-steps', '200'],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return [ln for ln in out.stdout.splitlines() if ln.strip()]


def test_reference_arm_json_line():
    lines = _run({'RANK': '0', 'WORLD_SIZE': '2'})
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'updates/s' and d['higher_is_better'] is True
    assert d['n_gpus'] == 2 and d['steps'] == 1 and d['warmup'] == 0 and d['scaling'] == 'weak'
    assert d['value'] > 1e6 and d['vs_baseline'] is None and d['data'] == 'synthetic'
    sys.path.insert(0, ROOT)
    from oracle import ref_kernels
    # 'reference' = the reference's own kernels compiled for the host (oracle/_ref), 'port' only when absent
    assert d['cpu_baseline']['kind'] == ('reference' if ref_kernels.available('fast') else 'port')
    assert d['cpu_baseline']['cores'] == os.cpu_count()
    assert d['cpu_baseline']['value'] == d['value'] == d['e2e']['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config'] and d['gpu_launches'] == 0


def test_reference_arm_other_ranks_are_silent():
    assert _run({'RANK': '1', 'WORLD_SIZE': '2'}) == []
