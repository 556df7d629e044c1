"""ctypes binding of the wideband channeliser (include/nrsc5_b200.h, csrc/channelizer.cu): one cu8 capture at
23 814 000 S/s -> FM channels at 744 187.5 S/s cs16, the format nrsc5b_push_cs16 / input_push_cs16 take.
No CPU fallback: constructing a Channelizer without a CUDA device raises."""
from __future__ import annotations

import ctypes

import numpy as np

from .engine import EngineError, _check, load_library

WIDE_RATE = 23814000.0          # 32 x 744 187.5
TAPS, PERIOD, DECIM = 256, 11907, 32


def _lib():
    L = load_library()
    if not getattr(L, "_chan_ready", False):
        vp, sz, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        L.nrsc5b_chan_create.argtypes = [ctypes.POINTER(vp), ci, vp, ci]
        L.nrsc5b_chan_destroy.argtypes = [vp]
        L.nrsc5b_chan_destroy.restype = None
        L.nrsc5b_chan_tables.argtypes = [vp, vp, vp]
        L.nrsc5b_chan_make_tables.argtypes = [vp, ci, vp, vp]
        L.nrsc5b_chan_outputs.argtypes = [sz]
        L.nrsc5b_chan_outputs.restype = ctypes.c_longlong
        L.nrsc5b_chan_run_device.argtypes = [vp, vp, sz, vp, sz, vp]
        L.nrsc5b_chan_run.argtypes = [vp, vp, sz, vp]
        L._chan_ready = True
    return L


def make_tables(offsets_100khz):
    """The integer tables of the definition, computed on the host (no device): taps[nch][256][2], phasor[11907][2]."""
    off = np.ascontiguousarray(offsets_100khz, dtype=np.int32)
    taps = np.empty((off.size, TAPS, 2), dtype=np.int16)
    ph = np.empty((PERIOD, 2), dtype=np.int16)
    _check(_lib().nrsc5b_chan_make_tables(off.ctypes.data, off.size, taps.ctypes.data, ph.ctypes.data), "nrsc5b_chan_make_tables")
    return taps, ph


def outputs(nbytes: int) -> int:
    return int(_lib().nrsc5b_chan_outputs(nbytes & ~63))


class Channelizer:
    def __init__(self, offsets_100khz, device: int = 0):
        self._L = _lib()
        self.offsets = np.ascontiguousarray(offsets_100khz, dtype=np.int32)
        self.nch = int(self.offsets.size)
        self._h = ctypes.c_void_p()
        _check(self._L.nrsc5b_chan_create(ctypes.byref(self._h), device, self.offsets.ctypes.data, self.nch), "nrsc5b_chan_create")

    def close(self):
        if self._h:
            self._L.nrsc5b_chan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tables(self):
        taps = np.empty((self.nch, TAPS, 2), dtype=np.int16)
        ph = np.empty((PERIOD, 2), dtype=np.int16)
        _check(self._L.nrsc5b_chan_tables(self._h, taps.ctypes.data, ph.ctypes.data), "nrsc5b_chan_tables")
        return taps, ph

    def run(self, cu8: np.ndarray) -> np.ndarray:
        """Host capture (uint8, I/Q interleaved) -> int16 array [nch][2 * outputs] (I, Q interleaved)."""
        a = np.ascontiguousarray(cu8, dtype=np.uint8)
        n = outputs(a.size)
        out = np.empty((self.nch, 2 * max(n, 0)), dtype=np.int16)
        if n > 0:
            _check(self._L.nrsc5b_chan_run(self._h, a.ctypes.data, a.size, out.ctypes.data), "nrsc5b_chan_run")
        return out

    def run_device(self, d_cu8: int, nbytes: int, d_out: int, out_stride: int, stream: int = 0):
        _check(self._L.nrsc5b_chan_run_device(self._h, ctypes.c_void_p(d_cu8), nbytes, ctypes.c_void_p(d_out), out_stride,
                                             ctypes.c_void_p(stream)), "nrsc5b_chan_run_device")
