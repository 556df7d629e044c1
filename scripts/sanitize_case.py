"""One bounded decode for compute-sanitizer (memcheck / racecheck / synccheck): BASELINE config 5's streams (MP1) or
config 3's (MP3, P3 through interleaver IV) with L2 framing on the device, checked against the oracle afterwards so
that the run under the tool is also a parity run.  usage: sanitize_case.py mp1|mp3 [streams] [frames]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                               # noqa: E402
import nrsc5_b200                                          # noqa: E402
from nrsc5_b200 import synth                               # noqa: E402

kind = sys.argv[1]
S = int(sys.argv[2]) if len(sys.argv) > 2 else (128 if kind == "mp1" else 64)
F = int(sys.argv[3]) if len(sys.argv) > 3 else (1 if kind == "mp1" else 3)
if kind == "mp1":
    caps = bench.make_captures(4, F)
else:
    caps = [synth.make_fm_mp3(nframes=F, seed=11 + i, lead_in=0, tail_blocks=2, cfo_hz=(0.0, 80.0)[i % 2]).cu8 for i in range(2)]
views, nbytes = bench.stream_views(caps, S, 0)
with nrsc5_b200.Engine(nstreams=S, input_capacity=nbytes + 4096, log_capacity=1 << 20) as e:
    e.enable_l2(True)
    for s, v in enumerate(views):
        e.push_cu8(s, np.ascontiguousarray(v))
    e.process()
    recs = e.drain_all()
print(kind, S, F, bench.parity_gate(views, recs, what="sanitizer case"))
