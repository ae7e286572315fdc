"""bench.py's control flow and JSON contract with the device stubbed out (no GPU): every key the driver reads
is present and the bookkeeping around the timed region runs — a typo in a rarely taken branch of bench.py must
not wait for a GPU box to be found.  No number produced here means anything."""
import json
import os
import sys

import numpy as np
import pytest

from conftest import ROOT
from mdpy_b200 import _native

sys.path.insert(0, ROOT)


class BenchStubDevice:
    def __init__(self, device=None):
        self.n = 0
        self.launches = 0
        self.rebuilds = 0

    def set_atoms(self, q, m):
        self.n = int(np.asarray(q).reshape(-1).size)

    def __getattr__(self, name):
        if name.startswith('set_') or name in ('reset_integrator', 'flush_l2', 'upload_positions', 'upload_velocities', 'close'):
            return lambda *a, **k: None
        raise AttributeError(name)

    def pinned_empty(self, shape, dtype=np.float32):
        return np.zeros(shape, dtype=dtype)

    def _energies(self):
        e = np.zeros(_native.NUM_ENERGIES)
        e[_native.E_LJ], e[_native.E_KINETIC] = -5.0, 2.0
        return e

    def step_langevin_host(self, x_in, v_in, x_out, v_out, dt, kT, gamma, seed, nsteps, terms):
        x_out[...] = x_in; v_out[...] = v_in
        self.launches += 15 * nsteps
        return self._energies()

    def step_langevin(self, dt, kT, gamma, seed, nsteps, terms):
        self.launches += 13 * nsteps
        self.rebuilds += nsteps // 9
        self._last_steps = nsteps

    def timing(self):
        k = getattr(self, '_last_steps', 1)
        return dict(nlist_ms=0.02 * k, pair_ms=0.05 * k, spread_ms=0.02 * k, fft_ms=0.03 * k, gather_ms=0.01 * k, bonded_ms=0.02 * k,
                    integrate_ms=0.01 * k, bare_ms=0.0, total_ms=0.1 * k, comm_ms=0.0, launches=float(self.launches),
                    rebuilds=float(self.rebuilds), pair_launches=float(k), work_units=1000.0, j_chunks=2000.0, masked_chunks=300.0,
                    seg_chunks=2.0, i_blocks=float((self.n + 31) // 32), shift_ok=1.0)

    def download_positions(self, unwrapped=False):
        return np.zeros((self.n, 3), dtype=np.float64 if unwrapped else np.float32)

    def download_velocities(self):
        return np.zeros((self.n, 3), dtype=np.float32)

    def pair_count(self):
        return 3_500_000

    def last_energies(self):
        return self._energies()


def _run_bench(monkeypatch, capsys, argv):
    import bench
    monkeypatch.setattr(_native, 'Device', BenchStubDevice)
    monkeypatch.setattr(bench, 'reference_step_seconds', lambda system, cfg, threads, budget_s=10.0: (5.0, 'stub sample'))
    monkeypatch.setattr(bench.time, 'sleep', lambda s: None)
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK', 'MDK_OPTS', 'MDK_TERMS_MASK'):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setattr(sys, 'argv', ['bench.py'] + argv)
    bench.main()
    return json.loads(capsys.readouterr().out.strip().splitlines()[-1])


def test_bench_line_has_every_key_of_the_contract(monkeypatch, capsys):
    line = _run_bench(monkeypatch, capsys, ['--steps', '20', '--warmup', '3', '--relax', '0.01', '--config', 'water_23k', '--no-sub'])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
                'dtype', 'data', 'config', 'clocks', 'gpu_launches', 'e2e', 'roofline', 'cpu_baseline', 'reps_ms'):
        assert key in line, key
    assert line['metric'] == 'atom_steps_per_s' and line['unit'] == 'atom-steps/s' and line['n_gpus'] == 1
    assert line['config']['workload'] == 'water_23k' and line['config']['atoms'] == 23556 and line['vs_baseline'] is None
    assert set(line['e2e']) >= {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'}
    assert line['e2e']['h2d_bytes_per_step'] == 24 * 23556 and line['e2e']['unit'] == line['unit']
    assert set(line['roofline']) >= {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'}
    assert line['roofline']['frac'] == pytest.approx(line['roofline']['achieved'] / line['roofline']['peak'])
    assert set(line['cpu_baseline']) >= {'value', 'unit', 'cores', 'kind', 'sample'} and line['cpu_baseline']['kind'] == 'port'
    assert len(line['reps_ms']) == 5                       # five timed repetitions of exactly K steps, the median is the value
    assert line['gpu_launches'] == 13 * 20                 # launches of ONE repetition of K steps
    assert line['value'] == pytest.approx(23556 * 20 / (0.1 * 20 * 1e-3))
    assert 'sub_records' not in line


def test_side_records_of_the_23k_and_92k_boxes_ride_in_the_single_gpu_line(monkeypatch, capsys):
    line = _run_bench(monkeypatch, capsys, ['--steps', '20', '--warmup', '3', '--relax', '0.01', '--config', 'water_23k', '--no-numba'])
    sub = line['sub_records']
    assert set(sub) == {'protein_92k'} and sub['protein_92k']['atoms'] == 92224
    rec = sub['protein_92k']
    for key in ('value', 'ms_per_step', 'ns_per_day', 'e2e', 'roofline', 'roofline_pme', 'phases_ms_per_step'):
        assert key in rec, key
    assert rec['roofline']['frac'] == pytest.approx(rec['roofline']['achieved'] / rec['roofline']['peak'])
    assert line['reference_cpu_recorded']['seconds_per_step'] == pytest.approx(419.5, rel=0.01)   # the reference itself, 1 core


def test_the_workload_is_the_same_at_every_gpu_count(monkeypatch):
    import bench
    seen = {}
    monkeypatch.setattr(bench, 'run_b200', lambda args, cfg: seen.update(config=args.config, grid=cfg['grid'], steps=args.steps))
    for argv in (['bench.py'], ['bench.py', '--gpus', '2'], ['bench.py', '--gpus', '8']):
        monkeypatch.setattr(sys, 'argv', argv)
        bench.main()
        assert seen['config'] == 'protein_1m' and tuple(seen['grid']) == (216, 216, 216) and seen['steps'] == 200
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--gpus', '2', '--config', 'protein_92k'])
    bench.main()
    assert seen['config'] == 'protein_92k'
