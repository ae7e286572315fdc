#!/usr/bin/env python
"""bench.py — headline benchmark of the mdpy nonbonded hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config NAME]

A "step" is one MD step of the hot path on one synthetic box: neighbour-list upkeep, CHARMM LJ +
erfc direct space over the tile list, PME reciprocal (spread, FFT, influence function, gather),
bonded terms and the Langevin position/velocity update.

Workload: the SAME box at every GPU count — BASELINE.json configs[3], the 1 066 628-atom solvated box
(12/10 A switch, PME 216^3, Langevin 300 K, 2 fs; it fits one GPU), so the driver's 1 -> 8 scaling curve is
the strong scaling of one system.  At N = 1 the line also carries `sub_records` for configs[1] (water_23k)
and configs[2] (protein_92k, the config the >= 10x target is quoted on), each with value / e2e / roofline
and the reference's own numba.cuda path run live on the same GPU from baseline/_ref.

Prints ONE JSON line (rank 0).  metric = atom-steps/s (BASELINE.json names "ns/day and atom-steps/s");
ns_per_day rides along.  value = state resident on the device: W warm-up steps, then REPS = 5 repetitions of
exactly K steps, each bracketed by a barrier and timed with CUDA events on the context stream (max over
ranks); the MEDIAN repetition is the value, all of them are listed in reps_ms.  The nvidia-smi clock sampler
starts >= 1 s before the first repetition while the GPU runs the same steps untimed, so the clocks are
sampled under the timed load whatever K is.  e2e = the same metric through the drop-in integrator call
with HOST state: every step is one LangevinIntegrator.integrate(ensemble, 1), i.e. positions + velocities
H2D, one step, positions + velocities + energies D2H (mdk_step_langevin_host).  roofline = the pair kernel
against the FP32 CUDA-core peak (70 flop per in-cutoff pair), roofline_pme = spread / FFT / gather
against measured HBM bandwidth.  cpu_baseline / --impl reference = the reference's per-step work
(27-cell-list LJ + all-pairs Coulomb, no PME) restated in C (oracle/), on the host cores.
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
    # skin_multi: list skin of decomposed runs.  A rebuild of a decomposed job costs every rank the global sort and a state
    # all-gather, and the skin shell is cheap for the pair kernel (far-class list order), so fewer, larger lists win there
    # (N = 8: 1.114 -> 1.068 ms/step); on one GPU 2.0 A stays best (3.83 vs 3.98 ms).  Physics is the same either way.
    'protein_1m': dict(gen='protein_1m', cutoff=12.0, switch=10.0, grid=(216, 216, 216), dt=2.0, skin_multi=3.0),
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
                                          '--format=csv,noheader,nounits', '-lms', '150'],
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


class NoSampler:
    def __enter__(self): return self
    def __exit__(self, *a): pass
    def summary(self): return dict(sm_mhz=None, sm_max_mhz=None, reasons=['not sampled on this rank'])


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
RELAX = ((0.1, 0.2, 200), (0.5, 0.05, 200), (1.0, 0.01, 300))   # (dt fs, gamma 1/fs, steps): untimed lattice relaxation
REPS = 5                                                         # timed repetitions of K steps; the median is reported


class Workload:
    """One synthetic box on this rank's GPU, relaxed and warmed up, ready to be timed."""

    def __init__(self, name, args):
        from mdpy_b200 import _native
        from mdpy_b200.integrator import LangevinIntegrator
        from mdpy_b200.unit import KB, Quantity, default_energy_unit, kelvin
        self.name, self.cfg = name, CONFIGS[name]
        cfg = self.cfg
        self.system = build_system(cfg)
        self.n = self.system.num_particles
        self.ens = self.system.ensemble(cutoff=cfg['cutoff'], switch=cfg['switch'], pme=True, ewald_error=1e-6,
                                        grid=cfg['grid'], order=4, bonded=True)
        self.ctx = _native.context_of(self.ens)
        self.dev = self.ctx.dev
        self.kT = float((Quantity(TEMPERATURE, kelvin) * KB).convert_to(default_energy_unit).value)
        self.dt = cfg['dt']
        multi = int(os.environ.get('WORLD_SIZE', '1')) > 1 and args.impl != 'reference'
        self.skin = float(os.environ.get('MDK_SKIN', cfg.get('skin_multi' if multi and 'skin_multi' in cfg else 'skin', 2.0)))   # experiments: MDK_SKIN=2.5
        if self.skin != 2.0:
            self.dev.set_nlist(self.skin)
        if args.no_graph:
            self.dev.set_option('graph', 0)
        for kv in filter(None, os.environ.get('MDK_OPTS', '').split(',')):   # experiments: MDK_OPTS=concurrent=0,graph=1
            k, v = kv.split('=')
            self.dev.set_option(k, float(v))
        # relax the lattice start (untimed; every rank runs it redundantly and deterministically, so all
        # ranks hold bit-identical state): short, strongly damped steps, then the production step
        for rdt, gamma, steps in RELAX:
            LangevinIntegrator(rdt, TEMPERATURE, gamma, seed=1).integrate(self.ens, max(1, int(steps * args.relax)))
        self.integ = LangevinIntegrator(self.dt, TEMPERATURE, GAMMA, seed=1)
        self.integ.integrate(self.ens, 20)
        self.terms = 0
        for c in self.ens.constraints:
            self.terms |= c.terms
        self.terms &= int(os.environ.get('MDK_TERMS_MASK', '0xffff'), 0)   # experiments: drop force terms from the direct step calls

    def step(self, nsteps):
        self.dev.step_langevin(self.dt, self.kT, GAMMA, 1, nsteps, self.terms)

    def phase_profile(self, steps):
        """Per-phase device times (ms per step): a separate untimed pass, per-phase events add syncs."""
        self.dev.set_profiling(2)
        self.step(steps)
        ph = self.dev.timing()
        self.dev.set_profiling(0)
        keys = ('nlist_ms', 'pair_ms', 'spread_ms', 'fft_ms', 'gather_ms', 'bonded_ms', 'integrate_ms', 'comm_ms')
        out = {k: ph[k] / steps for k in keys}
        out.update({k: ph[k] for k in ('work_units', 'j_chunks', 'masked_chunks', 'seg_chunks')})
        return out

    def pair_count(self):
        """In-cutoff pair count of the current configuration (the flop model's unit count)."""
        lj = self.ens.constraints[0]
        self.ens.state._positions = self.dev.download_positions()
        self.ctx._pos_rev = None
        lj._configure(); self.ctx.sync_positions()
        return self.dev.pair_count()

    def timed_reps(self, steps, reps, reduce_max, barrier):
        """reps x (K steps, device resident, CUDA events on the ctx stream around the whole call).
        Returns (per-rep ms [max over ranks], launches of this rank over all reps, rebuilds)."""
        dev = self.dev
        dev.set_profiling(1)
        before = dev.timing()
        ms = []
        for _ in range(reps):
            barrier()
            self.step(steps)
            ms.append(reduce_max(dev.timing()['total_ms']))
        barrier()
        after = dev.timing()
        dev.set_profiling(0)
        return ms, after['launches'] - before['launches'], int(after['rebuilds'] - before['rebuilds'])

    def e2e(self, steps, reduce_max, barrier):
        """The drop-in integrator call with HOST state, one call per step: every call uploads ensemble.state
        (positions + velocities, float32), runs one step and downloads the new state and the energies."""
        from mdpy_b200.integrator import LangevinIntegrator
        dev, ens = self.dev, self.ens
        ens.state._positions = dev.download_positions()      # the timed region stepped the device directly
        ens.state._velocities = dev.download_velocities()
        if hasattr(ens.state, 'revision'):
            ens.state.revision += 1
        integ = LangevinIntegrator(self.dt, TEMPERATURE, GAMMA, seed=2)
        for _ in range(5):   # warm-up of the host path (graph capture of the single-step variant, pinned buffers)
            integ.integrate(ens, 1)
        l0 = dev.timing()['launches']
        secs = []
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                integ.integrate(ens, 1)
            secs.append(reduce_max(time.perf_counter() - t0))
        sec = float(np.median(secs))
        return dict(value=self.n * steps / sec, unit='atom-steps/s', h2d_bytes_per_step=24 * self.n,
                    d2h_bytes_per_step=24 * self.n + 224, steps=steps, reps=3, ms_per_step=sec * 1e3 / steps,
                    ns_per_day=ns_per_day(steps, sec, self.dt),
                    gpu_launches_per_step=(dev.timing()['launches'] - l0) / (3 * steps),
                    potential_energy_last_step=float(ens.potential_energy),
                    path='LangevinIntegrator.integrate(ensemble, 1) per step: host State (float32 positions + velocities) '
                         '-> mdk_step_langevin_host -> host State + energies; wall clock, median of 3 x %d calls, max over ranks' % steps)


def roofline_records(w, ph, n_pairs, slots_single, world, peaks):
    cfg = w.cfg
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'pair_kernel_traffic.json'))).get(w.name)
    except (OSError, ValueError):
        pass
    fp32_peak = 148 * 128 * 2 * peaks['sm_max_mhz'] * 1e6 / 1e12
    share = (ph['j_chunks'] * 1024.0) / max(1.0, slots_single)      # this rank's share of the pair slots (1 at N = 1)
    achieved = FLOP_PER_PAIR * n_pairs * share / (ph['pair_ms'] * 1e-3) / 1e12
    K = int(np.prod(cfg['grid']))
    pme_ms = ph['spread_ms'] + ph['fft_ms'] + ph['gather_ms']
    pme_bytes = 44.0 * w.n + 34.0 * K
    pme_gbs = pme_bytes / max(pme_ms, 1e-9) / 1e6
    roof = dict(bound='fp32', achieved=achieved, peak=fp32_peak, unit='TFLOP/s', frac=achieved / fp32_peak,
                traffic=None if traffic is None or world > 1 else traffic['bytes'],
                traffic_source=None if traffic is None or world > 1 else traffic['source'] + ' (ncu dram bytes read + written per launch)',
                kernel='k_pair<LJ,COUL>', flop_per_pair=FLOP_PER_PAIR, pairs_in_cutoff=n_pairs, kernel_ms=ph['pair_ms'],
                peak_source='148 SM x 128 lanes x 2 flop x sm_max_mhz (%s)' % peaks['source'],
                note='rank 0 share of the pair work at N > 1' if world > 1 else 'whole pair kernel; kernel_ms = CUDA events around the launch, per-phase pass')
    roof_pme = dict(bound='hbm', achieved=pme_gbs, peak=peaks['hbm_gbs'], unit='GB/s', frac=pme_gbs / peaks['hbm_gbs'],
                    bytes_per_step=pme_bytes, kernels_ms=pme_ms, peak_source=peaks['source'],
                    note='spread + mesh FFT / influence function + gather (44 N + 34 K algorithmic bytes, SURVEY 8d); the mesh is L2 resident')
    return roof, roof_pme


def numba_cuda_reference(name, budget_s=170):
    """The reference's own numba.cuda path on this GPU, live (baseline/ref_numba_cuda.py in a subprocess: the
    unmodified package from baseline/_ref).  Falls back to the value recorded in round 1 when the subprocess fails."""
    rec = None
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, 'baseline', 'ref_numba_cuda.py'), '--config', name, '--evals', '2'],
                             capture_output=True, text=True, timeout=budget_s)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith('{'):
                d = json.loads(ln)
                if 'unavailable' not in d:
                    rec = dict(atom_steps_per_s=d['atom_steps_per_s'], ns_per_day=d['ns_per_day_at_2fs'],
                               seconds_per_step=d['seconds_per_step'], live=True,
                               source='baseline/ref_numba_cuda.py run now on this GPU (unmodified reference from baseline/_ref: '
                                      'plain-cutoff LJ + bare 27-cell Coulomb, no PME, no bonded terms, no integrator)')
                else:
                    rec = dict(unavailable=d['unavailable'], live=True)
                break
    except (subprocess.TimeoutExpired, OSError, ValueError, KeyError) as e:
        rec = dict(unavailable='live run failed: %r' % (e,), live=True)
    if rec is None or 'unavailable' in rec:
        try:
            old = json.load(open(os.path.join(ROOT, 'profiles', 'r01_reference_numba_cuda.json'))).get(name)
            if old:
                rec = dict(atom_steps_per_s=old['atom_steps_per_s'], ns_per_day=old['ns_per_day_at_2fs'], seconds_per_step=old['seconds_per_step'],
                           live=False, live_error=(rec or {}).get('unavailable'),
                           source='profiles/r01_reference_numba_cuda.json (recorded in round 1 on one B200)')
        except (OSError, ValueError):
            pass
    return rec


def sub_record(name, args, peaks, identity, nobarrier):
    """A single-GPU side record (water_23k, protein_92k) inside the N = 1 line: the same measurements as the
    headline workload, so that the 92k-atom target config is in every driver-run line."""
    w = Workload(name, args)
    steps = max(args.steps, 200 if w.n > 50000 else 500)
    w.step(max(args.warmup, 20))
    ms, _, rebuilds = w.timed_reps(steps, REPS, identity, nobarrier)
    med = float(np.median(ms))
    ph = w.phase_profile(100)
    n_pairs = w.pair_count()
    roof, roof_pme = roofline_records(w, ph, n_pairs, ph['j_chunks'] * 1024.0, 1, peaks)
    e2e = w.e2e(200, identity, nobarrier)
    rec = dict(workload=name, atoms=w.n, value=w.n * steps / (med * 1e-3), unit='atom-steps/s', steps=steps, reps_ms=ms,
               ms_per_step=med / steps, ns_per_day=ns_per_day(steps, med * 1e-3, w.dt), nlist_rebuilds_per_rep=rebuilds / REPS,
               e2e=e2e, roofline=roof, roofline_pme=roof_pme,
               phases_ms_per_step={k: ph[k] for k in ('nlist_ms', 'pair_ms', 'spread_ms', 'fft_ms', 'gather_ms', 'bonded_ms', 'integrate_ms')},
               slot_efficiency=n_pairs / max(1.0, ph['j_chunks'] * 1024.0))
    if not args.no_numba:
        ref = numba_cuda_reference(name)
        if ref:
            rec['reference_numba_cuda'] = ref
            if 'ns_per_day' in ref:
                rec['speedup_vs_reference_numba_cuda'] = dict(device_resident=rec['ns_per_day'] / ref['ns_per_day'],
                                                              e2e=e2e['ns_per_day'] / ref['ns_per_day'])
    w.dev.close()
    return rec


def run_b200(args, cfg):
    import mdpy_b200 as md  # noqa: F401
    from mdpy_b200 import multigpu

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

    w = Workload(args.config, args)
    n, dev, dt = w.n, w.dev, w.dt
    prof_steps = 30 if n > 200000 else 100
    ph1 = w.phase_profile(prof_steps)                      # single-GPU profile (every rank, before the job is sharded)
    n_pairs = w.pair_count()
    slots_single = ph1['j_chunks'] * 1024.0

    # ---- multi-GPU: join the communicator, shard the work ----
    grid = (1, 1, 1)
    if world > 1:
        # one spatial domain per rank, halo exchange per step; the mesh rank's domain is smaller by what its mesh chain costs
        # (single-GPU profile: FFT + influence function, plus the sub-mesh conversions it does for the other ranks)
        weights = multigpu.domain_weights(world, ph1['pair_ms'], 1.8 * ph1['fft_ms'])
        grid = multigpu.attach(w.ctx, dist, rank, world, weights=weights)
    w.integ.integrate(w.ens, max(args.warmup, 3))

    # ---- timed region.  The clock sampler runs from >= 1 s before the first timed repetition to after the last
    # one, and the GPU does the same steps (untimed) during the pre-roll, so every clock sample is taken under
    # the load that is being timed — however short K steps are.
    # (one sampler for the job — rank 0's GPU: eight nvidia-smi pollers take the driver lock often enough to slow the
    # host-launched decomposed step)
    with (ClockSampler(local) if rank == 0 else NoSampler()) as clocks:
        t_pre = time.perf_counter()
        while time.perf_counter() - t_pre < 1.2:
            w.step(max(args.steps, 20))
        ms, launches_rank, rebuilds = w.timed_reps(args.steps, REPS, max_over_ranks, barrier)
        t_post = time.perf_counter()
        while time.perf_counter() - t_post < 0.25:
            w.step(max(args.steps, 20))
    dev_ms = float(np.median(ms))
    launches = int(sum_over_ranks(launches_rank) / REPS)

    if args.skip_extras:
        if rank == 0:
            print(json.dumps(dict(metric='atom_steps_per_s', value=n * args.steps / (dev_ms * 1e-3), unit='atom-steps/s',
                                  ns_per_day=ns_per_day(args.steps, dev_ms * 1e-3, dt), steps=args.steps, warmup=args.warmup,
                                  n_gpus=world, ms_per_step=dev_ms / args.steps, reps_ms=ms, gpu_launches=launches,
                                  rebuilds_per_rep=rebuilds / REPS, note='skip-extras (profiling run)')))
        if dist is not None:
            dist.destroy_process_group()
        return

    ph_n = w.phase_profile(prof_steps)                     # per-phase profile of the (possibly sharded) step, ranks in lockstep
    # L2-flushed variant: one step at a time with a 256 MB memset in between (outside the events)
    dev.set_profiling(1)
    fl = []
    for _ in range(min(30, max(args.steps, 10))):
        dev.flush_l2()
        w.step(1)
        fl.append(dev.timing()['total_ms'])
    dev.set_profiling(0)
    e2e = w.e2e(min(max(args.steps, 20), 500 if n < 200000 else 50), max_over_ranks, barrier)

    if rank != 0:
        dist.destroy_process_group()
        return

    peaks = measured_peaks()
    roof, roof_pme = roofline_records(w, ph_n, n_pairs, slots_single, world, peaks)
    # ---- CPU baseline (bounded sample, rank 0, N = 1 only) ----
    cpu = None
    if world == 1:
        cpu_sec, sample = reference_step_seconds(w.system, cfg, 1, budget_s=10.0)
        cpu = dict(value=n / cpu_sec, unit='atom-steps/s', cores=1, kind='port', sample=sample,
                   host_cores=os.cpu_count(), seconds_per_step=cpu_sec, ns_per_day=ns_per_day(1, cpu_sec, dt))

    phase_keys = ('nlist_ms', 'pair_ms', 'spread_ms', 'fft_ms', 'gather_ms', 'bonded_ms', 'integrate_ms', 'comm_ms')
    K = int(np.prod(cfg['grid']))
    line = dict(
        metric='atom_steps_per_s', value=n * args.steps / (dev_ms * 1e-3), unit='atom-steps/s', n_gpus=world,
        steps=args.steps, warmup=args.warmup, ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling='strong',
        vs_baseline=None, dtype='f32', data='synthetic', ns_per_day=ns_per_day(args.steps, dev_ms * 1e-3, dt),
        timing='median of %d repetitions of K = %d steps, each bracketed by a barrier and timed with CUDA events on the '
               'context stream, max over ranks per repetition; reps_ms lists them all' % (REPS, args.steps),
        reps_ms=ms,
        scaling_note='the same workload (%s) at every GPU count: strong scaling of one box; single-GPU side records of '
                     'BASELINE configs[1] (water_23k) and configs[2] (protein_92k, the >=10x target config) ride along at N = 1' % args.config,
        config=dict(workload=args.config, atoms=n, cutoff_A=cfg['cutoff'], switch_A=cfg['switch'], pme_grid=list(cfg['grid']),
                    pme_order=4, ewald_error=1e-6, dt_fs=dt, integrator='langevin_gjf_300K_1ps', skin_A=w.skin,
                    terms='lj+erfc_direct+pme_recip+bond+angle+dihedral+improper', nlist_rebuilds_per_rep=rebuilds / REPS,
                    parallelism='single GPU' if world == 1 else
                    'spatial domain decomposition %d x %d x %d, one domain per rank: halo positions out / halo forces back by grouped '
                    'ncclSend / ncclRecv every step, owner-only integration, PME sub-meshes to / from the mesh rank, state all-gather at '
                    'list rebuilds' % grid,
                    domain_decomposition=None if world == 1 else dict(grid=list(grid), weights=np.round(weights, 3).tolist(), **dev.dd_stats()),
                    l2='steady-state MD trajectory: every step consumes the previous step\'s output, nothing is re-timed '
                       'on a repeated input; working set %.1f MB (> L2 for this box: %s); l2_flushed_ms_per_step gives the same '
                       'step with a 256 MB L2 flush before it' % ((190.0 * n + 20.0 * K) / 1e6, (190.0 * n + 20.0 * K) > 126e6)),
        l2_flushed_ms_per_step=float(np.median(fl)),
        clocks=clocks.summary(), gpu_launches=launches,
        e2e=e2e, roofline=roof, roofline_pme=roof_pme,
        phases_ms_per_step={k: ph_n[k] for k in phase_keys},
        phases_ms_per_step_single_gpu={k: ph1[k] for k in phase_keys},
        nlist=dict(work_units=int(ph_n['work_units']), j_chunks=int(ph_n['j_chunks']), masked_chunks=int(ph_n['masked_chunks']),
                   seg_chunks=int(ph_n['seg_chunks']), pair_slots=int(slots_single),
                   slot_efficiency=n_pairs / max(1.0, slots_single)),
    )
    if cpu is not None:
        line['cpu_baseline'] = cpu
    if world == 1 and not args.no_sub:
        dev.close()
        identity, nobarrier = (lambda x: float(x)), (lambda: None)
        line['sub_records'] = {}
        for name in ('water_23k', 'protein_92k'):
            if name != args.config:
                line['sub_records'][name] = sub_record(name, args, peaks, identity, nobarrier)
        try:   # recorded, not live: the unmodified reference's CPU path on the 23k box (1 core, DOUBLE mode)
            g = np.load(os.path.join(ROOT, 'tests', 'golden', 'config2_full_f64.npz'))
            sec = float(np.sum(g['ref_seconds']))
            line['reference_cpu_recorded'] = dict(
                workload='water_23k', seconds_per_step=sec, atom_steps_per_s=23556 / sec, ns_per_day=ns_per_day(1, sec, 2.0),
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
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='protein_1m', choices=sorted(CONFIGS),
                    help='headline workload, the same at every GPU count (default: protein_1m, BASELINE configs[3])')
    ap.add_argument('--no-sub', action='store_true', help='N = 1: skip the water_23k / protein_92k side records')
    ap.add_argument('--no-numba', action='store_true', help='side records: skip the live run of the reference numba.cuda path')
    ap.add_argument('--relax', type=float, default=1.0, help='scale of the untimed lattice-relaxation phase')
    ap.add_argument('--skip-extras', action='store_true', help='only the timed region (for ncu runs)')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from the host (no CUDA-graph steps)')
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == 'reference':
        run_reference(args, cfg)
    else:
        run_b200(args, cfg)


if __name__ == '__main__':
    main()
