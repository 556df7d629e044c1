"""bench.py's supplementary legs (BASELINE configs 3 and 4) run on the CPU emulation of the kernels at toy sizes:
their capture generation, push / rewind / drain plumbing, gates and JSON line - not their numbers."""
import json
import os
import sys
import types

import pytest

import common
import port

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))

pytestmark = pytest.mark.skipif(not port.available(), reason="oracle/_ref/liboracle.so not built")


@pytest.fixture(scope="module", autouse=True)
def emulated_engine():
    import build_emu
    from nrsc5_b200 import engine as eng
    so = build_emu.build()
    saved = (eng.lib_path, eng._lib)
    eng.lib_path = lambda: so
    eng._lib = None
    yield
    eng.lib_path, eng._lib = saved


def test_am_config4_leg(capsys):
    import bench
    bench.am_leg(types.SimpleNamespace(am_streams=3, am_frames=10, steps=2))
    out = json.loads([ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")][-1])
    assert out["value"] > 0 and out["e2e"]["value"] > 0 and out["p1_frames_per_channel"] >= 16
    assert out["e2e"]["h2d_bytes_per_step"] > 3 * 2_000_000
