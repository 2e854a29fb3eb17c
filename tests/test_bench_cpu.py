"""CPU: the reference arm of bench.py (`--impl reference`, the oracle port on the host cores) prints ONE JSON line with the
contract's keys -- on the `tiny` workload so that it takes seconds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'tiny', '--steps', '3',
                        '--warmup', '1'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'impl', 'cpu_baseline', 'e2e', 'gpu_launches'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['gpu_launches'] == 0 and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and 'median' in d['cpu_baseline']['sample']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['value'] == d['value']
    assert d['config']['workload'] == 'tiny' and d['config']['color_dims'] == 3
