"""Trajectory frames from the device's asynchronous frame capture (the on-disk side of the step loop).

The reference's dumpers (mdpy/dumper/*.py) read `ensemble.state.positions` from the host once per dump period, which on a
device-resident integrator would mean one call — and one full State round trip — per frame.  Here the step call itself
leaves every stride-th frame in page-locked host memory (mdk_set_frame_capture: a copy stream moves the frames out while
the next steps run) and this class appends them to a flat binary file:

    bytes 0-7    magic b'MDPYB2TR'
    int32        number of atoms n, int32 stride
    float64[3]   box edge lengths
    then         float32 [n,3] wrapped positions per frame
"""
import struct

import numpy as np

from . import _native

MAGIC = b'MDPYB2TR'


class FrameDumper:
    def __init__(self, path, stride, max_frames_per_call=1024):
        self.path, self.stride, self.max_frames = path, int(stride), int(max_frames_per_call)
        self.num_frames = 0
        self._header_written = False

    def integrate(self, integrator, ensemble, num_steps):
        """integrator.integrate(ensemble, num_steps) with every stride-th frame of the call appended to the file."""
        ctx = _native.context_of(ensemble)
        want = min(self.max_frames, num_steps // self.stride)
        if want < num_steps // self.stride:
            raise ValueError('%d frames per call exceed max_frames_per_call = %d' % (num_steps // self.stride, self.max_frames))
        ctx.dev.set_frame_capture(self.stride, max(want, 1))
        try:
            integrator.integrate(ensemble, num_steps)
            frames = ctx.dev.get_frames(max(want, 1))
        finally:
            ctx.dev.set_frame_capture(0, 0)
        with open(self.path, 'ab' if self._header_written else 'wb') as f:
            if not self._header_written:
                box = np.asarray(ensemble.state.pbc_matrix, dtype=np.float64).diagonal()
                f.write(MAGIC + struct.pack('<ii', frames.shape[1], self.stride) + box.astype('<f8').tobytes())
                self._header_written = True
            f.write(np.ascontiguousarray(frames, dtype='<f4').tobytes())
        self.num_frames += len(frames)
        return frames


def read_frames(path):
    """-> (frames float32 [k,n,3], box float64 [3], stride)."""
    with open(path, 'rb') as f:
        if f.read(8) != MAGIC:
            raise ValueError('%s is not a mdpy_b200 frame file' % path)
        n, stride = struct.unpack('<ii', f.read(8))
        box = np.frombuffer(f.read(24), dtype='<f8').copy()
        data = np.frombuffer(f.read(), dtype='<f4')
    return data.reshape(-1, n, 3).copy(), box, stride
