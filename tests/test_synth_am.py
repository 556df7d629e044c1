"""The AM MA1 / MA3 modulator (nrsc5_b200/synth_am.py) against the UNMODIFIED reference: every P1, P3 and PIDS frame
the reference decodes from a generated capture is a frame the generator put in, and the reference's events on
it match the committed golden file (tests/golden/synth_am.json) - the known answers for the AM rows."""
import numpy as np
import pytest

import common
import reftap
from nrsc5_b200 import synth_am

pytestmark = pytest.mark.skipif(not reftap.available(), reason="reference oracle not built")


def _pack(b):
    return np.packbits(np.asarray(b, dtype=np.uint8)).tobytes()


@pytest.mark.parametrize("name", list(common.AM_CASES))
def test_reference_decodes_generated_am_frames(name):
    cap = synth_am.make_am_ma1(**common.AM_CASES[name])
    log = reftap.decode(cap.cs16, mode=reftap.MODE_AM)
    frames = [p for t, p in log.records if t == reftap.REC_FRAME]
    p1 = [p["bits"] for p in frames if p["lc"] == 0 and p["nbits"] == 3750]
    psmi = common.AM_CASES[name].get("psmi", 1)
    p3 = [p["bits"] for p in frames if p["lc"] == 1 and p["nbits"] == (24000 if psmi == 1 else 30000)]
    pids = [p["bits"] for t, p in log.records if t == reftap.REC_PIDS]
    gen_p1 = {_pack(b) for fr in cap.p1_frames.values() for b in fr}
    gen_p3 = {_pack(b) for b in cap.p3_frames.values()}
    gen_pids = {_pack(b) for b in cap.pids_frames}
    assert len(p1) >= 32 and all(b in gen_p1 for b in p1)
    assert len(p3) >= 4 and all(b in gen_p3 for b in p3)
    assert len(pids) >= 60 and all(b in gen_pids for b in pids)
    syncs = [p for t, p in log.records if t == reftap.REC_SYNC]
    assert len(syncs) == 1 and syncs[0]["psmi"] == psmi
    g = common.golden("synth_am.json")[name]
    if common.fnv1a32(cap.cs16[:1 << 18].tobytes()) == g["input_fnv"]:
        assert common.summarize(log) == g["events"]
