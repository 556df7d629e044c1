"""One small AM decode (MA3, 10 frames, one stream) for compute-sanitizer; checked against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import nrsc5_b200, common, port, bench
from nrsc5_b200 import synth_am
name = sys.argv[1] if len(sys.argv) > 1 else "ma3_clean"
cap = synth_am.make_am_ma1(**common.AM_CASES[name])
c = cap.cs16[: cap.cs16.size & ~1]
with nrsc5_b200.Engine(nstreams=1, input_capacity=2 * c.size + 4096, log_capacity=4 << 20, mode="am") as e:
    e.push_cs16(0, c); e.process(); recs = e.drain(0)
print(name, len(recs), "records; gate:", bench.parity_gate([c], [recs], am=True, what="sanitizer AM case"))
