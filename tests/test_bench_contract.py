"""CPU: the reference arm of bench.py (`--impl reference`: the unmodified reference's step on the host cores when its
sources are present (/root/reference or the mirror oracle/_ref), else the CPU port; the one place besides the tests where
bench.py executes `oracle/`) prints exactly one JSON line with the keys the measurement contract names.
The product arm needs a GPU; its line is checked by the driver."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference'
    assert d['metric'] == 'vlmo_base_pretrain_samples_per_sec' and d['unit'] == 'samples/s'
    assert d['value'] > 0 and d['ms_per_step'] > 0
    assert d['n_gpus'] == 1 and d['steps'] == 1 and d['warmup'] == 1
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert d['data'] == 'synthetic' and 'workload' in d['config'] and 'model' not in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1 and cb['unit'] == 'samples/s' and cb['value'] == d['value'] and cb['sample']
    e = d['e2e']
    assert e['value'] == d['value'] and e['unit'] == 'samples/s' and e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0


def test_reference_arm_under_torchrun_prints_once():
    """N > 1: rank 0 alone runs the CPU port and prints the line, the other ranks exit 0 without work."""
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', '29643', os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
                        '--warmup', '1'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['n_gpus'] == 2 and d['value'] > 0


def test_reference_arm_other_workloads():
    """configs[3] (VQA at 480^2) and configs[4] (ITC only) have their own reference-arm lines."""
    for wl, metric in (('itc4096', 'vlmo_base_itc_samples_per_sec'),):
        r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', wl, '--steps', '1',
                            '--warmup', '0', '--cpu-batch', '2'], capture_output=True, text=True, timeout=600, cwd=ROOT)
        assert r.returncode == 0, r.stderr[-2000:]
        d = json.loads([l for l in r.stdout.splitlines() if l.startswith('{')][0])
        assert d['metric'] == metric and d['value'] > 0 and d['impl'] == 'reference'
