"""bench.py --impl reference (the arm the driver runs first, on the host cores): every workload prints ONE JSON line with the
contract's keys.  Small --batch values keep the bounded CPU samples to a few seconds here."""
import json
import os
import subprocess
import sys

import pytest

from oracle import ref_ecos, ref_osqp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {'impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
        'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'gpu_launches'}


@pytest.mark.parametrize('workload,batch', [('mpc', 400), ('mpc_ltv', 96), ('portfolio_socp', 48)])
def test_reference_arm_prints_the_contract_line(workload, batch):
    if workload == 'portfolio_socp' and not ref_ecos.available() or workload != 'portfolio_socp' and not ref_osqp.available():
        pytest.skip('oracle/_ref not built')
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', workload,
                          '--steps', '1', '--warmup', '1', '--batch', str(batch)], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert KEYS <= set(line) and line['impl'] == 'reference' and line['value'] > 0 and line['gpu_launches'] == 0
    assert line['cpu_baseline']['kind'] == 'reference' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e'] == {'value': line['value'], 'unit': line['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in line['config'] and line['unit'] == 'instances/s' and line['dtype'] == 'f64'


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '1'],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ''
