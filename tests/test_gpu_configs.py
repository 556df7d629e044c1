"""BASELINE configs 3 and 5 at their own stream counts, every stream against the CPU oracle.

The streams are bench.py's: D distinct synthetic captures replicated with distinct start offsets (zero and non-zero),
so every stream takes its own acquisition path; the oracle decodes every distinct (capture, offset) view and every
L1 PDU (P1 / P3 / PIDS) and sync event of every stream must be the oracle's, in order.  bench.py applies the same
gate to its timed workloads (`parity_gate`)."""
import numpy as np
import pytest

import bench
import nrsc5_b200
from nrsc5_b200 import engine as eng
from nrsc5_b200 import synth

pytestmark = pytest.mark.gpu


def _run(views, nbytes, log_cap):
    S = len(views)
    with nrsc5_b200.Engine(nstreams=S, input_capacity=nbytes + 4096, log_capacity=log_cap) as e:
        for s, v in enumerate(views):
            e.push_cu8(s, np.ascontiguousarray(v))
        e.process()
        recs = e.drain_all()
        st = e.stats()
    return recs, st


def test_config3_64_mp3_streams_bit_exact():
    """64 concurrent FM MP3 streams (reference src/decode.c:344-437, src/sync.c:552-573): P1, PIDS and P3 through
    interleaver IV, the decode groups enabled on demand while 64 streams wait for them."""
    S, F = 64, 4
    caps = [synth.make_fm_mp3(nframes=F, seed=11 + i, lead_in=0, tail_blocks=2, cfo_hz=(0.0, 80.0)[i % 2]).cu8 for i in range(2)]
    views, nbytes = bench.stream_views(caps, S, 0)
    recs, _ = _run(views, nbytes, (F + 1) * (18272 + 64) + 8 * F * (576 + 32) + 128 * 1024)
    gate = bench.parity_gate(views, recs, what="config 3")
    assert gate["distinct_views"] == S
    p3 = [sum(1 for t, r in rr if t == eng.REC_FRAME and r["lc"] == 1) for rr in recs]
    assert min(p3) >= 8 * (F - 3) - 2 and max(p3) >= 8 * (F - 2) - 2         # P3 starts after two frames of interleaver fill


def test_config5_128_mp1_streams_bit_exact():
    """bench.py's headline workload (BASELINE config 5's per-GPU shard): 128 FM MP1 streams x 4 L1 frames, 4 distinct
    captures (clean, CFO, CFO + noise) at 32 distinct offsets each."""
    S, F = 128, 4
    caps = bench.make_captures(4, F)
    views, nbytes = bench.stream_views(caps, S, 0)
    recs, st = _run(views, nbytes, (F + 1) * (18272 + 64) + 96 * 1024)
    gate = bench.parity_gate(views, recs, what="config 5")
    assert gate["distinct_views"] == S and gate["pdus_compared"] >= S * (3 + 48)
    assert int(st.p1_frames) >= 3 * S
