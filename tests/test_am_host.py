"""The AM engine's code (nrsc5_b200/csrc/am.cuh, AM_HD functions) compiled for the CPU and run with one lane
(tests/am_host.cu) against the oracle and the golden vectors, and with the 32 lanes of the GPU warp emulated by fibres:
the same source the GPU runs."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import common
import port
import reftap
from nrsc5_b200 import synth_am

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libam_host.so")

pytestmark = pytest.mark.skipif(not port.available(), reason="oracle/_ref/liboracle.so not built")


def _lib():
    src = [os.path.join(HERE, "am_host.cu"), os.path.join(common.ROOT, "nrsc5_b200", "csrc", "am.cuh"),
           os.path.join(common.ROOT, "nrsc5_b200", "csrc", "am_tables.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(p) > os.path.getmtime(SO) for p in src):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.run(["nvcc", "-std=c++17", "--expt-extended-lambda", "--expt-relaxed-constexpr", "-O2", "-shared", "-Xcompiler", "-fPIC", "-o", SO,
                        src[0], "-L" + os.path.join(common.ROOT, "oracle", "_ref"), "-loracle",
                        "-Xlinker", "-rpath", "-Xlinker", os.path.join(common.ROOT, "oracle", "_ref")], check=True)
    L = ctypes.CDLL(SO)
    L.am_host_decode.restype = ctypes.c_long
    L.am_host_decode.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
    L.am_host_decode_lanes.restype = ctypes.c_long
    L.am_host_decode_lanes.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                       ctypes.c_int]
    return L


def host_decode(cs16, lanes=1, order=1):
    a = np.ascontiguousarray(cs16, dtype=np.int16)
    buf = ctypes.create_string_buffer(8 << 20)
    if lanes == 1:
        n = _lib().am_host_decode(a.ctypes.data, a.size & ~1, buf, len(buf))
    else:
        n = _lib().am_host_decode_lanes(a.ctypes.data, a.size & ~1, buf, len(buf), lanes, order)
    assert n != -2, "lanes ended with different private AmState copies"
    assert n != -3, "lanes met at different AM_SYNC() sites"
    assert n >= 0
    return reftap._parse(buf.raw[:n])


@pytest.mark.parametrize("name", list(common.AM_CASES))
def test_am_engine_code_on_host_matches_oracle(name):
    cap = synth_am.make_am_ma1(**common.AM_CASES[name])
    got = host_decode(cap.cs16)
    ref = port.decode_am(cap.cs16)
    assert common.summarize(got) == common.summarize(ref)


@pytest.mark.parametrize("name,order", [("ma1_noisy", 1), ("ma1_noisy", -1), ("ma3_noisy", 1), ("ma3_noisy", -1)])
def test_am_engine_code_with_32_emulated_lanes(name, order):
    """k_am's warp emulated by 32 fibres (tests/am_host.cu): the lane-strided work split and the placement of the
    AM_SYNC() barriers, with the lanes scheduled in ascending and in descending order."""
    cap = synth_am.make_am_ma1(**common.AM_CASES[name])
    got = host_decode(cap.cs16, lanes=32, order=order)
    ref = port.decode_am(cap.cs16)
    assert common.summarize(got) == common.summarize(ref)
