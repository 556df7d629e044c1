"""TEST INFRASTRUCTURE ONLY (tests/, bench.py's gates).  The channeliser's definition restated in numpy: exact integer arithmetic on the tables the
library publishes (nrsc5b_chan_make_tables).  The reference has no channeliser - there is nothing to pin this to but the
definition itself (include/nrsc5_b200.h) and the end-to-end check that the stations mixed into a wideband capture decode."""
import numpy as np

TAPS, PERIOD, DECIM = 256, 11907, 32


def channelize(cu8: np.ndarray, offsets, taps: np.ndarray, phasor: np.ndarray, n0: int = 0) -> np.ndarray:
    """cu8: uint8 I/Q interleaved (length a multiple of 64 is used); returns int16 [nch][2 * nout].  n0: index of the
    first output when cu8 is a slice of a longer capture starting at its sample 32 * n0 (the mixer runs on)."""
    a = np.asarray(cu8, dtype=np.uint8)
    a = a[: a.size & ~63]
    ns = a.size // 2
    nout = (ns - TAPS) // DECIM + 1 if ns >= TAPS else 0
    xr = a[0::2].astype(np.int64) - 127
    xi = a[1::2].astype(np.int64) - 127
    out = np.zeros((len(offsets), 2 * max(nout, 0)), dtype=np.int16)
    if nout <= 0:
        return out
    idx = (np.arange(nout)[:, None] * DECIM + np.arange(TAPS)[None, :])         # [nout][256] sample indices 32 n + u
    XR, XI = xr[idx], xi[idx]
    n = np.arange(nout, dtype=np.int64) + int(n0)
    for k, m in enumerate(offsets):
        wr = taps[k, :, 0].astype(np.int64)
        wi = taps[k, :, 1].astype(np.int64)
        ar = XR @ wr - XI @ wi
        ai = XI @ wr + XR @ wi
        vr = (ar + (1 << 12)) >> 13
        vi = (ai + (1 << 12)) >> 13
        step = (1600 * int(m)) % PERIOD
        q = (step * (n % PERIOD)) % PERIOD
        pr = phasor[q, 0].astype(np.int64)
        pi = phasor[q, 1].astype(np.int64)
        zr = (vr * pr + vi * pi + (1 << 14)) >> 15                                # v * conj(P)
        zi = (vi * pr - vr * pi + (1 << 14)) >> 15
        out[k, 0::2] = np.clip(zr, -32768, 32767).astype(np.int16)
        out[k, 1::2] = np.clip(zi, -32768, 32767).astype(np.int16)
    return out
