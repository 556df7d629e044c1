"""ctypes wrapper for oracle/_ref/liboracle.so — our plain-C restatement of the
reference FM receive chain (oracle/nrsc5_oracle.c).  TEST INFRASTRUCTURE ONLY.
"""
import ctypes
import os

import numpy as np

from reftap import _parse, RefLog  # same record format

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "_ref", "liboracle.so")
REC_BLOCK = 9
_lib = None


def available():
    return os.path.exists(PORT_SO)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(PORT_SO)
        L.orc_new.restype = ctypes.c_void_p
        L.orc_free.argtypes = [ctypes.c_void_p]
        L.orc_want_soft.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.orc_want_blocks.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.orc_push_cu8.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        L.orc_push_cs16.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        L.orc_log_size.restype = ctypes.c_size_t
        L.orc_log_size.argtypes = [ctypes.c_void_p]
        L.orc_log_data.restype = ctypes.c_void_p
        L.orc_log_data.argtypes = [ctypes.c_void_p]
        L.orc_am_new.restype = ctypes.c_void_p
        L.orc_am_free.argtypes = [ctypes.c_void_p]
        L.orc_am_push_cs16.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        L.orc_am_push_cu8.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        L.orc_am_log_size.restype = ctypes.c_size_t
        L.orc_am_log_size.argtypes = [ctypes.c_void_p]
        L.orc_am_log_data.restype = ctypes.c_void_p
        L.orc_am_log_data.argtypes = [ctypes.c_void_p]
        L.orc_halfband_fm.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        L.orc_viterbi.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_uint, ctypes.c_uint, ctypes.c_uint]
        L.orc_deinterleave_p1.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_deinterleave_pids.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p]
        L.orc_descramble.argtypes = [ctypes.c_void_p, ctypes.c_uint]
        L.orc_bit_errors_fm.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.orc_rs_decode.argtypes = [ctypes.c_void_p]
        L.orc_fix_header.argtypes = [ctypes.c_void_p]
        L.orc_p1_sync_lost.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_l2_new.restype = ctypes.c_void_p
        L.orc_l2_free.argtypes = [ctypes.c_void_p]
        L.orc_l2_frames.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
        L.orc_l2_log_size.restype = ctypes.c_size_t
        L.orc_l2_log_size.argtypes = [ctypes.c_void_p]
        L.orc_l2_log_data.restype = ctypes.c_void_p
        L.orc_l2_log_data.argtypes = [ctypes.c_void_p]
        L.orc_l2_lost.restype = ctypes.c_uint
        L.orc_l2_lost.argtypes = [ctypes.c_void_p]
        _lib = L
    return _lib


def decode(cu8: np.ndarray, chunk: int = 0, want_soft=False, want_blocks=False) -> RefLog:
    """cu8 (uint8) capture, or cs16 (int16, already decimated) capture."""
    L = lib()
    is_cs16 = np.asarray(cu8).dtype == np.int16
    a = np.ascontiguousarray(cu8, dtype=np.int16 if is_cs16 else np.uint8)
    n = a.size & (~1 if is_cs16 else ~3)
    o = L.orc_new()
    try:
        L.orc_want_soft(o, int(want_soft))
        L.orc_want_blocks(o, int(want_blocks))
        push = L.orc_push_cs16 if is_cs16 else L.orc_push_cu8
        item = 2 if is_cs16 else 1
        if chunk <= 0:
            push(o, a.ctypes.data, n)
        else:
            chunk &= ~3
            for off in range(0, n, chunk):
                push(o, a.ctypes.data + off * item, min(chunk, n - off))
        raw = ctypes.string_at(L.orc_log_data(o), L.orc_log_size(o))
    finally:
        L.orc_free(o)
    return _parse(raw)


def decode_am(samples: np.ndarray, chunk: int = 0) -> RefLog:
    """AM (MA1 / MA3) capture through oracle/nrsc5_oracle_am.c: int16 = cs16 at 46 511.72 S/s, uint8 = cu8 at
    1 488 375 S/s (decimated by 32 first, like input_push_cu8 in AM mode)."""
    L = lib()
    a = np.ascontiguousarray(samples)
    is_cs16 = a.dtype == np.int16
    assert is_cs16 or a.dtype == np.uint8
    n = a.size & (~1 if is_cs16 else ~3)
    push = L.orc_am_push_cs16 if is_cs16 else L.orc_am_push_cu8
    o = L.orc_am_new()
    try:
        if chunk <= 0:
            push(o, a.ctypes.data, n)
        else:
            chunk &= ~1 if is_cs16 else ~3
            for off in range(0, n, chunk):
                push(o, a.ctypes.data + a.itemsize * off, min(chunk, n - off))
        raw = ctypes.string_at(L.orc_am_log_data(o), L.orc_am_log_size(o))
    finally:
        L.orc_am_free(o)
    return _parse(raw)


def decimate_am(cu8: np.ndarray) -> np.ndarray:
    """The AM cu8 front end alone (five halfband stages, /32) from a zero state: int16 I/Q interleaved."""
    a = np.ascontiguousarray(cu8, dtype=np.uint8)
    n = a.size & ~3
    out = np.empty(2 * (n // 64), dtype=np.int16)
    L = lib()
    L.orc_am_decimate.restype = ctypes.c_size_t
    got = L.orc_am_decimate(ctypes.c_void_p(a.ctypes.data), ctypes.c_size_t(n), ctypes.c_void_p(out.ctypes.data))
    assert got == n // 64
    return out


def halfband_fm(cu8: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(cu8, dtype=np.uint8)
    npairs = a.size // 4
    out = np.empty(2 * npairs, dtype=np.int16)
    lib().orc_halfband_fm(a.ctypes.data, npairs, out.ctypes.data)
    return out


def viterbi(soft: np.ndarray, k=7, gens=(0o133, 0o171, 0o165)) -> np.ndarray:
    s = np.ascontiguousarray(soft, dtype=np.int8)
    n = s.size // 3
    out = np.empty(n, dtype=np.uint8)
    lib().orc_viterbi(s.ctypes.data, out.ctypes.data, k, n, *gens)
    return out


def deinterleave_p1(pm: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(pm, dtype=np.int8)
    out = np.empty(438528, dtype=np.int8)
    lib().orc_deinterleave_p1(a.ctypes.data, out.ctypes.data)
    return out


def deinterleave_pids(pm: np.ndarray, bc: int) -> np.ndarray:
    a = np.ascontiguousarray(pm, dtype=np.int8)
    out = np.empty(240, dtype=np.int8)
    lib().orc_deinterleave_pids(a.ctypes.data, bc, out.ctypes.data)
    return out


def descramble(bits: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(bits, dtype=np.uint8).copy()
    lib().orc_descramble(b.ctypes.data, b.size)
    return b


def rs_decode(block255: np.ndarray):
    b = np.ascontiguousarray(block255, dtype=np.uint8).copy()
    rc = lib().orc_rs_decode(b.ctypes.data)
    return rc, b


def fix_header(buf96: np.ndarray):
    b = np.ascontiguousarray(buf96, dtype=np.uint8).copy()
    ok = lib().orc_fix_header(b.ctypes.data)
    return ok, b


def p1_sync_lost(bits: np.ndarray):
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    pci = ctypes.c_uint32(0)
    lost = lib().orc_p1_sync_lost(b.ctypes.data, ctypes.byref(pci))
    return bool(lost), pci.value


def l2_frames(frames, raw=False):
    """L2 framing (oracle/nrsc5_oracle_l2.c) over a sequence of L1 PDUs: (lc, nbits, packed bits) or None
    (= frame_reset).  Returns the record stream (every frame followed by the L2 -> L3 calls it causes) and the
    number of times the sync-loss predicate fired."""
    from reftap import pack_frames
    L = lib()
    o = L.orc_l2_new()
    try:
        blob = pack_frames(frames)
        rc = L.orc_l2_frames(o, blob, len(blob))
        assert rc == 0
        rawlog = ctypes.string_at(L.orc_l2_log_data(o), L.orc_l2_log_size(o))
        lost = L.orc_l2_lost(o)
    finally:
        L.orc_l2_free(o)
    return (rawlog if raw else _parse(rawlog)), lost


def l1_to_l2_input(records):
    """The L1 record stream of a decode -> the frame list L2 sees: frame_reset where fine sync was entered
    (REC_SYNC, reference src/sync.c:405-409), then every REC_FRAME in order."""
    out = []
    for ty, r in records:
        if ty == 3:
            out.append(None)
        elif ty == 1:
            out.append((r["lc"], r["nbits"], r["bits"]))
    return out
