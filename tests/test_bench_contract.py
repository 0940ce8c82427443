"""CPU tier: bench.py's contract where it can be exercised without a GPU -- the `--impl reference`
arm (the reference's own CPU path through oracle/_ref, or the numpy port when oracle/_ref did not
travel) prints ONE JSON line with the agreed keys, ranks other than 0 stay silent, and the product
arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
SMALL = 'u1_16x16_nb128_nlf8_f32'        # BASELINE cfg 1, the reference's own CPU-runnable case


def run(*args, env=None):
    e = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    e.update(env or {})
    return subprocess.run([sys.executable, str(ROOT / 'bench.py'), *args], capture_output=True, text=True, env=e,
                          timeout=600, cwd=str(ROOT))


def test_reference_arm_prints_one_contract_line():
    r = run('--impl', 'reference', '--workload', SMALL, '--steps', '2', '--warmup', '1')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'gpu_launches'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['steps'] == 2 and d['warmup'] == 1 and d['gpu_launches'] == 0
    assert d['unit'] == 'link-updates/s' and d['higher_is_better'] is True and d['vs_baseline'] is None
    assert d['config']['workload'] == SMALL and 'model' not in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1 and cb['sample'] and cb['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    # value = link-updates of the sample / time: 128 chains * 16*16*2 links * 8 leapfrog steps per trajectory
    assert d['value'] == pytest.approx(128 * 512 * 8 / (d['ms_per_step'] * 1e-3), rel=1e-6)


def test_reference_arm_other_ranks_do_no_work():
    r = run('--impl', 'reference', '--workload', SMALL, '--steps', '1', '--warmup', '0', '--gpus', '2',
            env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert r.returncode == 0 and r.stdout.strip() == ''


@pytest.mark.skipif(torch.cuda.is_available(), reason='GPU present')
def test_product_arm_fails_loudly_without_gpu():
    r = run('--workload', SMALL, '--steps', '1', '--warmup', '0')
    assert r.returncode != 0
    assert 'no CPU fallback' in r.stderr
    assert not r.stdout.strip()
