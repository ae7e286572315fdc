"""Asynchronous frame capture (mdk_set_frame_capture / mdpy_b200.dumper.FrameDumper): every stride-th step of a device-resident
step call leaves its frame in page-locked host memory, copied out while the next steps run; the frames are the trajectory."""
import numpy as np
import pytest

from mdpy_b200 import synthetic
from mdpy_b200.dumper import FrameDumper, read_frames
from mdpy_b200.integrator import LangevinIntegrator

pytestmark = pytest.mark.gpu


def small():
    s = synthetic.water_box(2000, 5, box=np.full(3, 39.2))
    ens = s.ensemble(cutoff=9.0, pme=True, grid=(40, 40, 40))
    LangevinIntegrator(0.25, 300, 0.05, seed=3).integrate(ens, 300)
    return s, ens


def test_captured_frames_are_the_trajectory(tmp_path):
    s, ens_a = small()
    _, ens_b = small()
    ia, ib = LangevinIntegrator(1.0, 300, 0.001, seed=7), LangevinIntegrator(1.0, 300, 0.001, seed=7)
    dumper = FrameDumper(str(tmp_path / 'traj.bin'), stride=10)
    frames = dumper.integrate(ia, ens_a, 40)
    assert frames.shape == (4, 6000, 3) and frames.dtype == np.float32
    assert np.array_equal(frames[-1], ens_a.state.positions)            # the last frame is the State the call hands back
    box = s.box
    for k in range(4):                                                  # the same run in four calls of ten steps
        ib.integrate(ens_b, 10)
        d = frames[k].astype(np.float64) - ens_b.state.positions
        d -= box * np.round(d / box)
        assert np.abs(d).max() < 1e-4, k
    frames2 = dumper.integrate(ia, ens_a, 25)                           # 25 steps: frames at 10 and 20
    assert frames2.shape[0] == 2 and dumper.num_frames == 6
    got, box_f, stride = read_frames(str(tmp_path / 'traj.bin'))
    assert got.shape == (6, 6000, 3) and stride == 10 and np.allclose(box_f, box)
    assert np.array_equal(got[:4], frames) and np.array_equal(got[4:], frames2)
    # capture off again: a plain call leaves no frames behind
    ia.integrate(ens_a, 10)
    from mdpy_b200 import _native
    assert len(_native.context_of(ens_a).dev.get_frames(4)) == 0
