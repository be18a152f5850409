"""bench.py prints ONE JSON line with the keys the driver reads (reference arm on CPU here; our arm needs a B200)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
             'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'cpu_baseline'}


def _run(*args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    """--impl reference: the reference's operator sequence (oracle port) on the host cores, same metric / config."""
    d = _run('--impl', 'reference', '--steps', '1', '--warmup', '0', '--ref-views', '1')
    assert BASE_KEYS <= set(d) and d['impl'] == 'reference'
    assert d['metric'] == 'aggregation_frames_per_s' and d['unit'] == 'frames/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['config']['workload'] == 'MultiviewC-shaped aggregation forward'      # the GPU arm's workload
    assert d['config']['batch_per_gpu'] == 4 and d['config']['views'] == 7 and d['config']['channels'] == 256
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']


@pytest.mark.gpu
def test_our_arm_line():
    d = _run('--steps', '3', '--warmup', '3', '--cpu-views', '1')
    assert BASE_KEYS | {'clocks', 'gpu_launches', 'roofline'} <= set(d)
    assert d['metric'] == 'aggregation_frames_per_s' and d['n_gpus'] == 1 and d['steps'] == 3 and d['value'] > 0
    assert d['vs_baseline'] is None and d['data'] == 'synthetic' and d['scaling'] == 'weak'
    assert d['config']['kernel_path'] == 'fside_tf32x3' and 'workload' in d['config']
    assert d["gpu_launches"] == 10 * 3
    e = d['e2e']
    assert e['value'] > 0 and e['h2d_bytes_per_step'] == 4 * 7 * 256 * (90 * 160 + 45 * 80 + 23 * 40) * 4
    assert e['d2h_bytes_per_step'] == 4 * 256 * 156 * 156 * 4 and e['value'] < d['value']
    assert set(e['variants']) == {'channels_last_f32', 'channels_last_bf16'}
    assert e['variants']['channels_last_bf16']['h2d_bytes_per_step'] * 2 == e['h2d_bytes_per_step']
    r = d['roofline']
    assert r['bound'] in ('hbm', 'tensor') and r['unit'] in ('GB/s', 'TFLOP/s')
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9 and 0 < r['frac'] < 1
    assert [k['kernel'] for k in r['kernels']] == ['pool_tile_kernel', 'ygemm_compact_kernel']
    assert r['kernel'] in ('pool_tile_kernel', 'ygemm_compact_kernel')
    # both terms printed, frac reproducible from the line
    st = r['step']
    assert abs(st['frac'] - max(st['t_bytes_ms'], st['t_flops_ms']) / st['t_measured_ms']) < 1e-9
    assert abs(st['t_bytes_ms'] - st['algorithmic_bytes'] / (st['hbm_peak_gbs'] * 1e9) * 1e3) < 1e-9
    assert abs(st['t_measured_ms'] - d['ms_per_step']) < 1e-9
    pk = r['kernels'][0]
    assert pk['algorithmic_bytes_per_launch'] == st['algorithmic_bytes'] and pk['formulation_bytes'] > pk['algorithmic_bytes_per_launch']
    # BASELINE configs 2-5 ride on the same line
    assert [c['workload'].split('-')[0] for c in d['configs']] == ['MultiviewX', 'Wildtrack']
    assert all(c['value'] > 0 and c['batch'] == 1 for c in d['configs'])
    assert d['config4']['frames_per_rank'] == 64 and d['config4']['value'] > 0 and d['strong'] is None
    c5 = d['config5']
    assert c5['value'] > 0 and c5['batch_per_gpu'] == 1 and c5['kernel_path'] == 'fside_tf32x3' and c5['e2e']['value'] > 0
    v = d['variants']
    assert v['static_cameras']['value'] > d['value'] and v['bf16_mma']['kernel_path'] == 'fside_bf16mma' and v['bf16_mma']['tolerance']
    assert {'sm_mhz', 'sm_max_mhz', 'reasons'} <= set(d['clocks'])
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['value'] > 0 and cb['cores'] >= 1


@pytest.mark.gpu
def test_train_mode_line():
    """--mode train (BASELINE config 5): one optimiser step of the whole detector; same line format, plus the split of the
    step into network forward / aggregation forward / aggregation forward + backward."""
    d = _run('--mode', 'train', '--steps', '3', '--warmup', '3')
    assert BASE_KEYS | {'clocks', 'gpu_launches', 'breakdown_ms'} <= set(d)
    assert d['metric'] == 'training_frames_per_s' and d['unit'] == 'frames/s' and d['n_gpus'] == 1 and d['value'] > 0
    assert d['config']['batch_per_gpu'] == 1 and d['config']['views'] == 7 and d['config']['kernel_path'] == 'fside_tf32x3'
    assert d['config']['parameters'] == 15597632                  # ResNet-18 trunk + laterals + 3 collapse layers + heads
    b = d['breakdown_ms']
    assert 0 < b['aggregation_forward'] < b['aggregation_forward_backward'] < b['train_step']
    assert b['network_forward_eval'] < b['train_step']
    e = d['e2e']
    assert e['h2d_bytes_per_step'] == 7 * 3 * 720 * 1280 * 4 and e['d2h_bytes_per_step'] == 4 and 0 < e['value'] <= d['value'] * 1.05
    assert d['final_loss'] == d['final_loss'] and d['final_loss'] > 0          # finite
