"""GPU parity on awkward inputs through the C ABI: dropouts with re-acquisition, streams that start in the middle of
a frame, inputs shorter than one acquisition window, pure noise, misuse of the push calls.

Their CPU twins run on the emulation of the kernels (tests/test_emu_engine.py)."""
import numpy as np
import pytest

import common
import port
import reftap
import nrsc5_b200
from nrsc5_b200 import engine as eng
from nrsc5_b200 import synth
from test_gpu_chain import kinds, oracle_kinds, pdus, run_engine

pytestmark = pytest.mark.gpu


def _same_as_oracle(cu8, chunk=None):
    ref = port.decode(cu8)
    recs = run_engine([cu8], chunk=chunk)[0]
    frames = [(r["lc"], r["nbits"], r["bits"]) for t, r in recs if t == eng.REC_FRAME]
    want = [(p["lc"], p["nbits"], p["bits"]) for t, p in ref.records if t == reftap.REC_FRAME]
    assert frames == want
    assert pdus(recs)[1] == ref.pids_frames
    assert kinds(recs) == oracle_kinds(ref)
    return recs, ref


def test_dropout_and_reacquisition():
    """Two transmissions with different carrier offsets and timing, separated by 0.4 s of noise: the garbage frame
    decoded across the gap, the sync loss it causes (or not - whatever the reference does) and the second
    acquisition must come out as in the oracle, record for record."""
    a = synth.make_fm_mp1(nframes=2, seed=501, lead_in=700, cfo_hz=80.0, noise_lsb=3.0, tail_blocks=3)
    b = synth.make_fm_mp1(nframes=2, seed=502, lead_in=1333, cfo_hz=-640.0, noise_lsb=3.0, start_bc=9, tail_blocks=2)
    rng = np.random.default_rng(503)
    gap = np.clip(np.rint(rng.standard_normal(2 * 600000) * 6 + 127), 0, 255).astype(np.uint8)
    cu8 = np.concatenate([a.cu8, gap, b.cu8])
    recs, ref = _same_as_oracle(cu8)
    assert kinds(recs).count("S") >= 2                      # it did lock twice
    _same_as_oracle(cu8, chunk=3 << 20)


def test_stream_starting_mid_frame_mp3():
    """The capture starts at block 11 of a frame, on an odd block: P1 must wait for the next block 0, the PX1
    interleaver for the next even block (decode.c:383-399)."""
    cap = synth.make_fm(psmi=3, nframes=4, seed=77, lead_in=50, tail_blocks=3, start_bc=11, noise_lsb=2.0)
    recs, ref = _same_as_oracle(cap.cu8)
    assert sum(1 for t, r in recs if t == eng.REC_FRAME and r["lc"] == 1) >= 4


def test_noise_only_and_short_inputs():
    rng = np.random.default_rng(9)
    noise = np.clip(np.rint(rng.standard_normal(2 * 1500000) * 20 + 127), 0, 255).astype(np.uint8)
    recs, ref = _same_as_oracle(noise)                       # whatever the reference makes of noise (normally nothing)
    short = noise[: 2 * 100000]                              # less than one acquisition window: nothing may come out
    assert run_engine([short])[0] == []
    assert run_engine([noise[:0]])[0] == []


def test_push_misuse_is_refused():
    with nrsc5_b200.Engine(nstreams=2, input_capacity=1 << 20, log_capacity=1 << 16) as e:
        ok = np.zeros(4096, dtype=np.uint8)
        with pytest.raises(eng.EngineError):
            e.push_cu8(0, ok[:4095])                         # not a multiple of 4 (input.c:103 asserts)
        with pytest.raises(eng.EngineError):
            e.push_cu8(2, ok)                                # no such stream
        with pytest.raises(eng.EngineError):
            e.push_cs16(0, np.zeros(64, dtype=np.int16))     # a cu8 engine
        with pytest.raises(eng.EngineError):
            e.push_cu8(0, np.zeros((1 << 20) + 4096, dtype=np.uint8))   # more than the stream's buffer can ever hold
        e.push_cu8(0, ok)
        e.process()
        assert e.drain(0) == [] and e.drain(1) == []
