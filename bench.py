#!/usr/bin/env python
"""bench.py — headline benchmark of the mdpy nonbonded hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config NAME]

A "step" is one MD step of the hot path on one synthetic box: neighbour-list upkeep, CHARMM LJ +
erfc direct space over the tile list, PME reciprocal (spread, cuFFT, influence function, gather),
bonded terms and the Langevin position/velocity update.  Default workload = BASELINE.json configs[1]:
23 556-atom TIP3P box, 9 A cutoff, PME 64^3 order 4, Langevin 300 K, 2 fs.

Prints ONE JSON line (rank 0).  metric = atom-steps/s (BASELINE.json names "ns/day and atom-steps/s";
atom-steps/s is the one that stays comparable when the box changes with the GPU count: N = 1 runs
configs[1], the 23k water box; N > 1 runs configs[3], the 1.07 M-atom box the 1/2/4/8-GPU numbers are
quoted on); ns_per_day rides along.  value = state resident on the device (CUDA events around K
steps); e2e = the same metric through the drop-in integrator call with HOST state: every step is one
LangevinIntegrator.integrate(ensemble, 1), i.e. positions + velocities H2D from page-locked memory,
one step, positions + velocities + energies D2H (mdk_step_langevin_host).  roofline = the pair kernel against the FP32 CUDA-core
peak (70 flop per in-cutoff pair), roofline_pme = spread/FFT/gather against measured
HBM bandwidth.  cpu_baseline / --impl reference = the reference's per-step work (27-cell-list LJ +
all-pairs Coulomb, no PME) restated in C (oracle/), on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (generator key, cutoff, switch, pme grid, dt fs)
    'water_23k': dict(gen='water_23k', cutoff=9.0, switch=None, grid=(64, 64, 64), dt=2.0),
    'protein_92k': dict(gen='protein_92k', cutoff=12.0, switch=10.0, grid=(108, 108, 80), dt=2.0),
    'protein_1m': dict(gen='protein_1m', cutoff=12.0, switch=10.0, grid=(216, 216, 216), dt=2.0),
    # BASELINE configs[4]: 10 000 002-atom water box, 1 A skin (rebuild stress), PME 480^3.  Generates in ~25 s
    # on the host; NOT run on a GPU in round 1 (the multi-GPU budget went into the 1M box).
    'water_10m': dict(gen='water_10m', cutoff=9.0, switch=None, grid=(480, 480, 480), dt=2.0, skin=1.0),
}
FLOP_PER_PAIR = 70.0
TEMPERATURE, GAMMA = 300.0, 1e-3   # K, 1/fs (= 1/ps)


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=float(d['hbm_gbs']), sm_max_mhz=float(d.get('sm_max_mhz', 1965.0)), source='measured')
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, device=0):
        self.device, self.proc, self.lines = device, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            p = [x.strip() for x in ln.split(',')]
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(p) > 5 + k and p[5 + k].lower().startswith('active'):
                    reasons.add(nm)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['unavailable'])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


def ns_per_day(steps, seconds, dt_fs):
    return steps / seconds * dt_fs * 86400.0 * 1e-6


def build_system(cfg):
    from mdpy_b200 import synthetic
    return synthetic.CONFIGS[cfg['gen']]()


# ---------------------------------------------------------------------------------------------
class ReferenceSampler:
    """One step of the reference's CPU path (LJ over its 27-cell list + all-pairs Coulomb, both restated
    in C in oracle/), timed on a bounded slice of the outer atom loop and scaled to all N atoms."""

    def __init__(self, system, cfg, threads):
        from oracle import cpu_oracle as ora
        self.ora, self.cfg, self.threads = ora, cfg, threads
        self.n = system.num_particles
        self.topo = system.topology()
        self.pos = system.positions.astype(np.float32)
        self.pbc = np.diag(system.box).astype(np.float32)
        self.table = system.lj_table()
        self.k = 4 * np.pi * float(np.float32(0.5727653))
        self.per_atom = None

    def _run(self, m):
        """(t_lj, t_coulomb) for atoms [0, m) of the outer loops."""
        o, t = self.ora, self.topo
        t0 = time.perf_counter()
        o.lj_cell(self.pos, self.table, self.pbc, self.cfg['cutoff'], t.bonded_particles, t.scaling_particles,
                  cell_cutoff=12.0, i_range=(0, m), threads=self.threads)
        t1 = time.perf_counter()
        o.coulomb_allpairs(self.pos, t.charges, self.pbc, t.bonded_particles, self.k, i_range=(0, m), threads=self.threads)
        return t1 - t0, time.perf_counter() - t1

    def step(self, budget_s):
        """Seconds per full step estimated from a sample sized to budget_s, and the sample description."""
        n = self.n
        floor = max(self.threads * 8, min(n, 64))
        if self.per_atom is None:   # calibrate once on a small slice
            probe = max(floor, min(n, 256))
            self.per_atom = sum(self._run(probe)) / probe
        m = int(min(n, max(floor, budget_s / max(self.per_atom, 1e-9))))
        t_lj, t_c = self._run(m)
        # LJ is linear in the slice; the all-pairs loop is triangular: slice [0, m) covers
        # m n - m(m+1)/2 of the n(n-1)/2 pairs
        frac_c = (m * n - m * (m + 1) / 2.0) / (n * (n - 1) / 2.0)
        full = t_lj * n / m + t_c / frac_c
        sample = ('atoms [0,%d) of %d of the outer loops of LJ (27-cell list, rc %.0f A) and all-pairs Coulomb, fp32, '
                  'scaled to N (LJ linear, Coulomb by pair count); reference semantics: no PME' % (m, n, self.cfg['cutoff']))
        return full, sample


def reference_step_seconds(system, cfg, threads, budget_s=12.0):
    return ReferenceSampler(system, cfg, threads).step(budget_s)


def run_reference(args, cfg):
    """--impl reference: the reference's own CPU algorithm for the path on the host cores; every step is a
    bounded sample of the workload, sized so that the whole run stays within ~150 s of CPU work."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    system = build_system(cfg)
    threads = os.cpu_count() or 1
    sampler = ReferenceSampler(system, cfg, threads)
    per_step = min(12.0, 150.0 / max(1, args.steps + args.warmup))
    times = []
    sample = ''
    for it in range(args.warmup + args.steps):
        t, sample = sampler.step(per_step)
        if it >= args.warmup:
            times.append(t)
    sec = float(np.mean(times))
    value = system.num_particles / sec
    line = dict(metric='atom_steps_per_s', value=value, unit='atom-steps/s', impl='reference', n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=sec * 1e3, higher_is_better=True, scaling='strong',
                vs_baseline=None, dtype='f32', data='synthetic', ns_per_day=ns_per_day(1, sec, cfg['dt']),
                config=dict(workload=args.config, atoms=system.num_particles, cutoff_A=cfg['cutoff'], dt_fs=cfg['dt'],
                            note='reference path = plain-cutoff LJ + bare all-pairs Coulomb (no PME in the reference tree)'),
                cpu_baseline=dict(value=value, unit='atom-steps/s', cores=threads, kind='port', sample=sample),
                e2e=dict(value=value, unit='atom-steps/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
def run_b200(args, cfg):
    import mdpy_b200 as md  # noqa: F401
    from mdpy_b200 import _native, multigpu
    from mdpy_b200.integrator import LangevinIntegrator
    from mdpy_b200.unit import KB, Quantity, default_energy_unit, kelvin

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    dist = torch = None
    if world > 1:
        # NCCL / torchrun may print banners on stdout: park fd 1 on stderr and keep the real stdout for the JSON line
        sys.stdout.flush()
        real_stdout = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)
        sys.stdout = real_stdout
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    os.environ['MDPY_B200_DEVICE'] = str(local)

    system = build_system(cfg)
    n = system.num_particles
    ens = system.ensemble(cutoff=cfg['cutoff'], switch=cfg['switch'], pme=True, ewald_error=1e-6, grid=cfg['grid'],
                          order=4, bonded=True)
    ctx = _native.context_of(ens)
    dev = ctx.dev
    kT = float((Quantity(TEMPERATURE, kelvin) * KB).convert_to(default_energy_unit).value)
    dt = cfg['dt']
    skin = float(cfg.get('skin', 2.0))
    if skin != 2.0:
        dev.set_nlist(skin)
    if args.no_graph:
        dev.set_option('graph', 0)
    for kv in filter(None, os.environ.get('MDK_OPTS', '').split(',')):   # experiments: MDK_OPTS=concurrent=0,graph=1
        k, v = kv.split('=')
        dev.set_option(k, float(v))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return float(x)
        t = torch.tensor([x], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return float(x)
        t = torch.tensor([x], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # relax the lattice start (untimed; every rank runs it redundantly and deterministically, so all
    # ranks hold bit-identical state): short, strongly damped steps, then the production step
    for rdt, gamma, steps in ((0.1, 0.2, 200), (0.5, 0.05, 200), (1.0, 0.01, 300)):
        LangevinIntegrator(rdt, TEMPERATURE, gamma, seed=1).integrate(ens, max(1, int(steps * args.relax)))
    integ = LangevinIntegrator(dt, TEMPERATURE, GAMMA, seed=1)
    integ.integrate(ens, 20)
    terms = 0
    for c in ens.constraints:
        terms |= c.terms
    terms &= int(os.environ.get('MDK_TERMS_MASK', '0xffff'), 0)   # experiments: drop force terms from the direct step calls

    # ---- per-phase profile (separate untimed pass; per-phase events add syncs) ----
    prof_steps = 50 if n > 200000 else 200
    dev.set_profiling(2)
    dev.step_langevin(dt, kT, GAMMA, 1, prof_steps, terms)
    ph = dev.timing()
    dev.set_profiling(0)
    pair_ms = ph['pair_ms'] / prof_steps
    pme_ms = (ph['spread_ms'] + ph['fft_ms'] + ph['gather_ms']) / prof_steps
    bonded_ms = ph['bonded_ms'] / prof_steps
    lj = ens.constraints[0]
    ens.state._positions = dev.download_positions()
    ctx._pos_rev = None
    lj._configure(); ctx.sync_positions()
    n_pairs = dev.pair_count()               # in-cutoff pair count of the current configuration (flop model)
    slots_single = dev.timing()['j_chunks'] * 1024.0

    # ---- multi-GPU: join the communicator, deal i-blocks to ranks weighted by the extra roles ----
    weights = None
    if world > 1:
        weights = multigpu.role_weights(world, pair_ms + ph['nlist_ms'] / prof_steps, pme_ms, 0.0)   # bonded terms are split evenly
        weights = multigpu.broadcast_array(dist, weights, rank)   # one set of shard ranges for all ranks
        multigpu.attach(ctx, dist, rank, world, weights)
    integ.integrate(ens, max(args.warmup, 3))

    # ---- timed region: K steps, state resident on the device, CUDA events on the ctx stream ----
    dev.set_profiling(1)
    t_before = dev.timing()
    barrier()
    with ClockSampler(local) as clocks:
        w0 = time.perf_counter()
        dev.step_langevin(dt, kT, GAMMA, 1, args.steps, terms)
        wall = time.perf_counter() - w0
    barrier()
    t_after = dev.timing()
    dev_ms = max_over_ranks(t_after['total_ms'])
    launches = int(sum_over_ranks(t_after['launches'] - t_before['launches']))
    rebuilds = int(t_after['rebuilds'] - t_before['rebuilds'])
    value = ns_per_day(args.steps, dev_ms * 1e-3, dt)

    if args.skip_extras:
        if rank == 0:
            print(json.dumps(dict(metric='atom_steps_per_s', value=n * args.steps / (dev_ms * 1e-3), unit='atom-steps/s',
                                  ns_per_day=value, steps=args.steps, warmup=args.warmup,
                                  n_gpus=world, ms_per_step=dev_ms / args.steps, gpu_launches=launches, rebuilds=rebuilds,
                                  note='skip-extras (profiling run)')))
        if dist is not None:
            dist.destroy_process_group()
        return

    # per-phase profile of the (possibly sharded) step, every rank in lockstep
    dev.set_profiling(2)
    dev.step_langevin(dt, kT, GAMMA, 1, prof_steps, terms)
    ph_n = dev.timing()
    dev.set_profiling(1)
    # L2-flushed variant: one step at a time with a 256 MB memset in between (outside the events)
    fl = []
    for _ in range(min(50, args.steps)):
        dev.flush_l2()
        dev.step_langevin(dt, kT, GAMMA, 1, 1, terms)
        fl.append(dev.timing()['total_ms'])
    dev.set_profiling(0)

    # ---- e2e: the drop-in integrator call with host state, one call per step (all ranks in lockstep) ----
    # Every call uploads ensemble.state (positions + velocities, float32, page-locked), runs one step and
    # downloads the new state and the energies; between calls the state lives in host memory only as far
    # as the API is concerned (the device continues its float64 trajectory when the host copy is unchanged).
    e2e_steps = min(args.steps, 1000 if n < 200000 else 100)
    ens.state._positions = dev.download_positions()      # the timed region above stepped the device directly
    ens.state._velocities = dev.download_velocities()
    ens.state.revision += 1
    e2e_integ = LangevinIntegrator(dt, TEMPERATURE, GAMMA, seed=2)
    for _ in range(5):   # warm-up of the host path (graph capture of the single-step variant, pinned buffers)
        e2e_integ.integrate(ens, 1)
    l_before = dev.timing()['launches']
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_integ.integrate(ens, 1)
    e2e_sec = max_over_ranks(time.perf_counter() - t0)
    e2e_launches = (dev.timing()['launches'] - l_before) / e2e_steps
    e2e_value = n * e2e_steps / e2e_sec
    e2e_energy = float(ens.potential_energy)

    if rank != 0:
        dist.destroy_process_group()
        return

    peaks = measured_peaks()
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'pair_kernel_traffic.json'))).get(args.config)
    except (OSError, ValueError):
        pass
    fp32_peak = 148 * 128 * 2 * peaks['sm_max_mhz'] * 1e6 / 1e12
    pair_ms_n = ph_n['pair_ms'] / prof_steps
    # this rank evaluates its share of the pair slots; at N=1 that is everything
    share = (ph_n['j_chunks'] * 1024.0) / max(1.0, slots_single)
    achieved = FLOP_PER_PAIR * n_pairs * share / (pair_ms_n * 1e-3) / 1e12
    K = int(np.prod(cfg['grid']))
    pme_bytes = 44.0 * n + 34.0 * K
    pme_gbs = pme_bytes / (pme_ms * 1e-3) / 1e9

    # ---- CPU baseline (bounded sample, rank 0, N = 1 only) ----
    cpu = None
    if world == 1:
        cpu_sec, sample = reference_step_seconds(system, cfg, 1, budget_s=10.0)
        cpu = dict(value=n / cpu_sec, unit='atom-steps/s', cores=1, kind='port', sample=sample,
                   host_cores=os.cpu_count(), seconds_per_step=cpu_sec, ns_per_day=ns_per_day(1, cpu_sec, dt))

    phase_keys = ('nlist_ms', 'pair_ms', 'spread_ms', 'fft_ms', 'gather_ms', 'bonded_ms', 'integrate_ms', 'comm_ms')
    line = dict(
        metric='atom_steps_per_s', value=n * args.steps / (dev_ms * 1e-3), unit='atom-steps/s', n_gpus=world,
        steps=args.steps, warmup=args.warmup, ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling='strong',
        vs_baseline=None, dtype='f32', data='synthetic', ns_per_day=value, wall_ms_per_step=wall * 1e3 / args.steps,
        scaling_note='default workloads follow BASELINE.json: --gpus 1 = configs[1] (23 556-atom water box), --gpus > 1 = configs[3] '
                     '(1 066 628-atom box, strong scaling of that box); atom-steps/s is the size-normalised metric that makes the '
                     'two comparable (one GPU: 2.25e8 at 23k, 2.20e8 at 92k atoms)',
        config=dict(workload=args.config, atoms=n, cutoff_A=cfg['cutoff'], switch_A=cfg['switch'], pme_grid=list(cfg['grid']),
                    pme_order=4, ewald_error=1e-6, dt_fs=dt, integrator='langevin_gjf_300K_1ps', skin_A=skin,
                    terms='lj+erfc_direct+pme_recip+bond+angle+dihedral+improper', nlist_rebuilds_in_timed=rebuilds,
                    parallelism='single GPU' if world == 1 else
                    'replicated positions, i-block sharded pair forces (weights %s), bonded terms split evenly, PME on last rank, int64 all-reduce per step'
                    % np.round(weights, 3).tolist(),
                    l2='steady-state MD trajectory: every step consumes the previous step\'s output, nothing is re-timed '
                       'on a repeated input; working set %.1f MB; l2_flushed_ms_per_step gives the same step with a '
                       '256 MB L2 flush before it' % ((32.0 * n + 12.0 * K) / 1e6)),
        l2_flushed_ms_per_step=float(np.median(fl)),
        clocks=clocks.summary(), gpu_launches=launches,
        e2e=dict(value=e2e_value, unit='atom-steps/s', h2d_bytes_per_step=24 * n, d2h_bytes_per_step=24 * n + 128,
                 steps=e2e_steps, ms_per_step=e2e_sec * 1e3 / e2e_steps, ns_per_day=ns_per_day(e2e_steps, e2e_sec, dt),
                 gpu_launches_per_step=e2e_launches, potential_energy_last_step=e2e_energy,
                 path='LangevinIntegrator.integrate(ensemble, 1) per step: host State (float32 positions + velocities, '
                      'page-locked) -> mdk_step_langevin_host -> host State + energies; wall clock, max over ranks'),
        roofline=dict(bound='fp32', achieved=achieved, peak=fp32_peak, unit='TFLOP/s', frac=achieved / fp32_peak,
                      traffic=None if traffic is None or world > 1 else traffic['bytes'],
                      traffic_source=None if traffic is None or world > 1 else traffic['source'] + ' (ncu dram bytes read + written per launch)',
                      kernel='k_pair<LJ,COUL>', flop_per_pair=FLOP_PER_PAIR, pairs_in_cutoff=n_pairs, kernel_ms=pair_ms_n,
                      peak_source='148 SM x 128 lanes x 2 flop x sm_max_mhz (%s)' % peaks['source'],
                      note='rank 0 share of the pair work at N > 1' if world > 1 else 'whole pair kernel'),
        roofline_pme=dict(bound='hbm', achieved=pme_gbs, peak=peaks['hbm_gbs'], unit='GB/s', frac=pme_gbs / peaks['hbm_gbs'],
                          bytes_per_step=pme_bytes, kernels_ms=pme_ms, peak_source=peaks['source'],
                          note='spread + convert + cuFFT R2C/C2R + convolve + gather (single-GPU pass); mesh is L2 resident'),
        phases_ms_per_step={k: ph_n[k] / prof_steps for k in phase_keys},
        phases_ms_per_step_single_gpu={k: ph[k] / prof_steps for k in phase_keys},
        nlist=dict(work_units=int(ph_n['work_units']), j_chunks=int(ph_n['j_chunks']), masked_chunks=int(ph_n['masked_chunks']),
                   seg_chunks=int(ph_n['seg_chunks']), pair_slots=int(slots_single),
                   slot_efficiency=n_pairs / max(1.0, slots_single)),
    )
    if cpu is not None:
        line['cpu_baseline'] = cpu
    try:   # recorded, not live: the reference's own numba.cuda path on one B200 (baseline/ref_numba_cuda.py)
        rec = json.load(open(os.path.join(ROOT, 'profiles', 'r01_reference_numba_cuda.json'))).get(args.config)
        if rec:
            line['reference_numba_cuda_recorded'] = dict(
                atom_steps_per_s=rec['atom_steps_per_s'], ns_per_day=rec['ns_per_day_at_2fs'], seconds_per_step=rec['seconds_per_step'],
                source='profiles/r01_reference_numba_cuda.json (unmodified reference, one B200, plain-cutoff LJ + bare Coulomb, no PME)')
    except (OSError, ValueError):
        pass
    if args.config == 'water_23k':   # recorded, not live: the unmodified reference's CPU path on this very box (1 core, DOUBLE mode)
        try:
            g = np.load(os.path.join(ROOT, 'tests', 'golden', 'config2_full_f64.npz'))
            sec = float(np.sum(g['ref_seconds']))
            line['reference_cpu_recorded'] = dict(
                seconds_per_step=sec, atom_steps_per_s=n / sec, ns_per_day=ns_per_day(1, sec, dt),
                source='tests/golden/config2_full_f64.npz:ref_seconds (oracle/make_golden.py --only config2_full: one LJ + one '
                       'Coulomb evaluation of the unmodified reference, numba CPU kernels, 1 core of the build container)')
        except (OSError, KeyError, ValueError):
            pass
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=200)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default=None, choices=sorted(CONFIGS),
                    help='default: water_23k (BASELINE configs[1]) at --gpus 1, protein_1m (configs[3]) at --gpus > 1')
    ap.add_argument('--relax', type=float, default=1.0, help='scale of the untimed lattice-relaxation phase')
    ap.add_argument('--skip-extras', action='store_true', help='only the timed region (for ncu runs)')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from the host (no CUDA-graph steps)')
    args = ap.parse_args()
    if args.config is None:
        args.config = 'water_23k' if args.gpus <= 1 else 'protein_1m'
    cfg = CONFIGS[args.config]
    if args.impl == 'reference':
        run_reference(args, cfg)
    else:
        run_b200(args, cfg)


if __name__ == '__main__':
    main()
