"""Import shims that let the UNMODIFIED reference (mdpy v0.2.x at /root/reference)
load in this container.  TEST INFRASTRUCTURE ONLY (oracle/): used by
oracle/make_golden.py to generate tests/golden/*.npz; never imported by the product.

Put this directory first on PYTHONPATH (python picks up sitecustomize automatically):
    PYTHONPATH=oracle/refshim:/root/reference python oracle/make_golden.py

Shims:
  1. empty stub modules for MDAnalysis / h5py / matplotlib, which mdpy.io imports
     at module scope (io/psf_parser.py:12, io/pdb_parser.py:12-13, io/dcd_parser.py:11,
     io/hdf5_parser.py:10, io/hdf5_writer.py:10) but the hot path never calls;
  2. np.round_ (removed in NumPy 2) used inside an njit at utils/pbc.py:42.
"""
import sys
import types

for _name in ('MDAnalysis', 'MDAnalysis.topology', 'MDAnalysis.topology.guessers',
              'h5py', 'matplotlib', 'matplotlib.pyplot'):
    sys.modules.setdefault(_name, types.ModuleType(_name))
sys.modules['MDAnalysis.topology.guessers'].guess_atom_type = lambda x: x

import numpy as _np
if not hasattr(_np, 'round_'):
    _np.round_ = _np.round
