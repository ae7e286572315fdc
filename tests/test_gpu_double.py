"""env.set_precision('DOUBLE') (mdpy/environment.py:23-42 switches the reference's arithmetic type): the device then evaluates
LJ, erfc direct space, the bonded terms, the excluded-pair correction and the all-pairs Coulomb sum in float64 on float64
positions and parameters (mdk_set_precision / mdk_set_params_f64); what remains of the difference to the float64 oracle and to
the reference's own DOUBLE-mode goldens is the 2^-40 fixed-point resolution of the accumulators.  The PME mesh stays float32."""
import numpy as np
import pytest

import mdpy_b200 as md
from conftest import load_golden, rel_rms
from mdpy_b200 import _native
from mdpy_b200.constraint import (CharmmAngleConstraint, CharmmBondConstraint, CharmmImproperConstraint,
                                  CharmmNonbondedConstraint, CharmmVDWConstraint, ElectrostaticConstraint, ElectrostaticPMEConstraint)
from mdpy_b200.core import Topology
from mdpy_b200.unit import coulomb_constant
from oracle import cpu_oracle as ora

pytestmark = pytest.mark.gpu


def ensemble_f64(g, **terms):
    n = g['positions'].shape[0]
    topo = Topology.from_tables(['X'] * n, g['masses'], g['charges'], g['bonded'], g['scaling'], **terms)
    ens = md.Ensemble(topo, np.diag(g['box']))
    ens.state.set_positions(np.asarray(g['positions'], dtype=np.float64))
    return ens


@pytest.mark.parametrize('name,lj_key,el_key', [('mix_small_f64', 'lj', 'coul'),
                                                ('config1_f64', 'CharmmNonbondedConstraint', 'ElectrostaticConstraint')])
def test_double_precision_matches_the_reference_double_mode_goldens(name, lj_key, el_key):
    md.env.set_precision('DOUBLE')
    g = load_golden(name)
    ens = ensemble_f64(g)
    assert ens.state.positions.dtype == np.float64
    lj = CharmmNonbondedConstraint(g['lj_table'], cutoff_radius=float(g['rc']))
    el = ElectrostaticConstraint()
    ens.add_constraints(lj, el)
    # sum of |pair energies|: the scale a cancelling LJ total is judged on (2^-40 fixed point x ~1e4 work units: ~1e-10 of it)
    scale = ora.nonbonded_bruteforce(g['positions'], g['box'], g['lj_table'], g['charges'], g['bonded'], g['scaling'],
                                     rc_lj=float(g['rc']), threads=8)['e_lj_abs']
    lj.update()
    assert lj.forces.dtype == np.float64
    # 2^-40 fixed-point accumulators: ~1e-12 absolute per work unit; on the relaxed small box (forces ~1e-4) that is 1.4e-9
    assert rel_rms(lj.forces, g[lj_key + '_forces']) < 5e-9
    assert abs(lj.potential_energy - float(g[lj_key + '_energy'])) < 1e-9 * scale
    el.update()
    # float64 inputs: the exact L/2 ties of the PDB coordinates (Q13) fall on the reference's side now
    assert rel_rms(el.forces, g[el_key + '_forces']) < 1e-9
    assert el.potential_energy == pytest.approx(float(g[el_key + '_energy']), rel=1e-10)


def test_double_precision_bonded_terms_match_reference_goldens_directly():
    md.env.set_precision('DOUBLE')
    g = load_golden('config1_f64')
    ens = ensemble_f64(g, bonds=g['CharmmBondConstraint_idx'], angles=g['CharmmAngleConstraint_idx'],
                       impropers=g['CharmmImproperConstraint_idx'])
    cs = dict(CharmmBondConstraint=CharmmBondConstraint(g['CharmmBondConstraint_par']),
              CharmmAngleConstraint=CharmmAngleConstraint(g['CharmmAngleConstraint_par']),
              CharmmImproperConstraint=CharmmImproperConstraint(g['CharmmImproperConstraint_par']))
    ens.add_constraints(*cs.values())
    for name, c in cs.items():
        c.update()
        # bonded parameters travel as float32 tables (mdk_set_bonded): 6e-8 relative
        assert rel_rms(c.forces, g[name + '_forces']) < 1e-6, name
        assert c.potential_energy == pytest.approx(float(g[name + '_energy']), rel=1e-6), name


def test_double_precision_switch_and_erfc_direct_space_match_float64_oracle():
    md.env.set_precision('DOUBLE')
    g = load_golden('mix_small_f64')
    ens = ensemble_f64(g)
    lj = CharmmVDWConstraint(g['lj_table'], cutoff_radius=12.0, switch_radius=10.0)
    pme = ElectrostaticPMEConstraint(cutoff_radius=12.0, alpha=0.30, grid=(32, 32, 32), order=4)
    ens.add_constraints(lj, pme)
    t = ora.nonbonded_bruteforce(g['positions'], g['box'], g['lj_table'], g['charges'], g['bonded'], g['scaling'],
                                 rc_lj=12.0, r_on=10.0, coul_mode=1, k_e=coulomb_constant(), alpha=0.30, rc_coul=12.0, threads=8)
    lj.update()
    assert rel_rms(lj.forces, t['f_lj']) < 5e-9          # fixed-point resolution (1.8e-9 measured; SINGLE: 3e-6)
    assert abs(lj.potential_energy - t['e_lj']) < 1e-9 * t['e_lj_abs']
    pme.update()
    e = _native.context_of(ens).dev.last_energies()
    assert abs(e[_native.E_COUL_DIRECT] - t['e_coul']) < 1e-9 * t['e_coul_abs']
    assert e[_native.E_PME_EXCL] == pytest.approx(t['e_excl'], rel=1e-9)
    # SINGLE on the same system for scale: the float32 pair kernel sits at ~3e-6
    md.env.set_precision('SINGLE')
    ens32 = md.Ensemble(Topology.from_tables(['X'] * len(g['positions']), g['masses'], g['charges'], g['bonded'], g['scaling']), np.diag(g['box']))
    ens32.state.set_positions(g['positions'].astype(np.float32))
    lj32 = CharmmVDWConstraint(g['lj_table'], cutoff_radius=12.0, switch_radius=10.0)
    ens32.add_constraints(lj32)
    lj32.update()
    assert 1e-8 < rel_rms(lj32.forces, t['f_lj']) < 1e-5
