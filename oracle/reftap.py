"""ctypes wrapper for oracle/_ref/libnrsc5_ref.so (the unmodified reference
library + reftap.c harness).  TEST INFRASTRUCTURE ONLY: importable from tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs, never from the
product package.
"""
import ctypes
import os
import struct
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libnrsc5_ref.so")

REC_FRAME, REC_PIDS, REC_SYNC, REC_LOST_SYNC, REC_MER, REC_BER, REC_HDC, REC_SOFT_PM, REC_BLOCK = range(1, 10)
REC_L2_SERVICE, REC_L2_ALIGN, REC_L2_AAS, REC_L2_PACKET = 16, 17, 18, 19   # the L2 -> L3 calls (reftap_l2.c)
MODE_FM, MODE_AM = 0, 1

_lib = None


def available():
    return os.path.exists(REF_SO)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(REF_SO)
        L.reftap_log_size.restype = ctypes.c_size_t
        L.reftap_log_data.restype = ctypes.c_void_p
        L.reftap_decode.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_size_t]
        L.reftap_l2_frames.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]
        L.reftap_bench.restype = ctypes.c_double
        L.reftap_bench.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        _lib = L
    return _lib


@dataclass
class RefLog:
    """Decoded reftap log, in reference call order."""
    records: list = field(default_factory=list)  # (type, payload-dict)

    def of(self, t):
        return [p for (ty, p) in self.records if ty == t]

    @property
    def p1_frames(self):
        return [p["bits"] for p in self.of(REC_FRAME) if p["lc"] == 0]

    def frames(self, lc):
        return [p["bits"] for p in self.of(REC_FRAME) if p["lc"] == lc]

    @property
    def pids_frames(self):
        return [p["bits"] for p in self.of(REC_PIDS)]


def _parse(raw: bytes) -> RefLog:
    out = RefLog()
    off = 0
    n = len(raw)
    while off < n:
        ty, plen = struct.unpack_from("<II", raw, off)
        pay = raw[off + 8: off + 8 + plen]
        off += 8 + ((plen + 3) & ~3)
        if ty == REC_FRAME:
            lc, nbits = struct.unpack_from("<II", pay, 0)
            rec = {"lc": lc, "nbits": nbits, "bits": bytes(pay[8:])}
        elif ty == REC_PIDS:
            rec = {"bits": bytes(pay[:10]), "crc_ok": (pay[10] if len(pay) > 10 else None)}
        elif ty == REC_SYNC:
            f, psmi = struct.unpack_from("<fi", pay)
            flags = struct.unpack_from("<4i", pay, 8) if len(pay) >= 24 else (-1, -1, -1, -1)   # pli, hppi, aabi, rdbi
            rec = {"freq_offset": f, "psmi": psmi, "flags": list(flags)}
        elif ty == REC_LOST_SYNC:
            rec = {}
        elif ty == REC_MER:
            lo, up = struct.unpack("<ff", pay)
            rec = {"lower": lo, "upper": up}
        elif ty == REC_BER:
            rec = {"cber": struct.unpack("<f", pay)[0]}
        elif ty == REC_HDC:
            prog, cnt = struct.unpack_from("<II", pay, 0)
            rec = {"program": prog, "data": bytes(pay[8:8 + cnt])}
        elif ty == REC_SOFT_PM:
            bc = struct.unpack_from("<I", pay, 0)[0]
            rec = {"bc": bc, "soft": np.frombuffer(pay[4:], dtype=np.int8).copy()}
        elif ty == REC_BLOCK:
            st, se, ang, pr, pi, cfo, start = struct.unpack("<iifffiq", pay[:32])
            rec = {"state": st, "samperr": se, "angle": ang, "phase": complex(pr, pi), "cfo": cfo, "start": start}
        elif ty in (REC_L2_SERVICE, REC_L2_ALIGN, REC_L2_AAS, REC_L2_PACKET):
            rec = parse_l2(ty, pay)
        else:
            raise ValueError(f"bad record type {ty}")
        out.records.append((ty, rec))
    return out


def parse_l2(ty, pay):
    """Payload of one L2 -> L3 call record (same layout in the reference tap, the oracle and the engine)."""
    if ty == REC_L2_SERVICE:
        k = ("program", "access", "type", "codec_mode", "blend_control", "gain", "common_delay", "latency")
        return dict(zip(k, struct.unpack("<8i", pay[:32])))
    if ty == REC_L2_ALIGN:
        return dict(zip(("program", "stream_id", "offset"), struct.unpack("<3I", pay[:12])))
    if ty == REC_L2_AAS:
        return {"data": bytes(pay)}
    k = ("program", "stream_id", "seq", "shape", "flags", "size")
    rec = dict(zip(k, struct.unpack("<6I", pay[:24])))
    rec["data"] = bytes(pay[24:24 + rec["size"]])
    return rec


def pack_frames(frames) -> bytes:
    """frames: iterable of (lc, nbits, packed bits) or None (= frame_reset) -> the byte layout reftap_l2_frames,
    the oracle's L2 entry and nrsc5b_l2_frames take."""
    out = bytearray()
    for f in frames:
        if f is None:
            out += struct.pack("<II", 0, 0)
            continue
        lc, nbits, bits = f
        nb = (nbits + 7) // 8
        assert len(bits) >= nb
        out += struct.pack("<II", lc, nbits) + bytes(bits[:nb]) + bytes((-nb) % 4)
    return bytes(out)


def l2_frames(frames, mode=MODE_FM) -> RefLog:
    """Feed L1 PDUs straight into the reference's frame_push() on a fresh handle; the log holds every frame
    followed by the L2 -> L3 calls it caused."""
    L = lib()
    raw_in = pack_frames(frames)
    L.reftap_reset()
    rc = L.reftap_l2_frames(raw_in, len(raw_in), mode)
    assert rc == 0
    raw = ctypes.string_at(L.reftap_log_data(), L.reftap_log_size())
    L.reftap_reset()
    return _parse(raw)


def decode(samples: np.ndarray, mode=MODE_FM, chunk=0, want_soft=False, want_l2=False) -> RefLog:
    """Run the reference on a capture (uint8 cu8 array, or int16 cs16 array).  want_l2: also record the L2 -> L3
    calls (types 16..19, reftap_l2.c) between the other records."""
    L = lib()
    a = np.ascontiguousarray(samples)
    is_cs16 = a.dtype == np.int16
    assert is_cs16 or a.dtype == np.uint8
    L.reftap_reset()
    L.reftap_want_soft(1 if want_soft else 0)
    L.reftap_want_l2(1 if want_l2 else 0)
    rc = L.reftap_decode(a.ctypes.data, a.size, mode, int(is_cs16), chunk)
    assert rc == 0
    raw = ctypes.string_at(L.reftap_log_data(), L.reftap_log_size())
    L.reftap_reset()
    L.reftap_want_l2(0)
    return _parse(raw)


def bench(buffers, mode=MODE_FM, reps=1):
    """Time `len(buffers)` independent reference sessions, one host thread each.
    Returns wall seconds."""
    L = lib()
    arrs = [np.ascontiguousarray(b) for b in buffers]
    is_cs16 = arrs[0].dtype == np.int16
    n = len(arrs)
    ptrs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in arrs])
    lens = (ctypes.c_size_t * n)(*[a.size for a in arrs])
    return L.reftap_bench(ptrs, lens, n, mode, int(is_cs16), reps)


def fnv1a32(b: bytes) -> int:
    h = 0x811C9DC5
    for x in b:
        h = ((h ^ x) * 0x01000193) & 0xFFFFFFFF
    return h
