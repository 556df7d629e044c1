"""Regenerates tests/golden/l2.json from the UNMODIFIED reference (oracle/_ref/libnrsc5_ref.so, frame.c observed
through the link-time taps of oracle/reftap_l2.c): the L2 -> L3 call stream of support/sample.xz (decoded through
the public API) and of the generated PDU sequences of tests/l2_cases.py (fed to the reference's frame_push).
Run from the repo root in the build container: python tests/golden/make_golden_l2.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import reftap  # noqa: E402
from common import load_sample  # noqa: E402
from l2_cases import AM_BITS, L2_CASES, l2_digest  # noqa: E402
from nrsc5_b200 import synth_l2  # noqa: E402

out = {}
log = reftap.decode(load_sample(), want_l2=True)
out["sample_xz"] = l2_digest([(t, r) for t, r in log.records if t in (1, 16, 17, 18, 19)])
for name, kw in L2_CASES.items():
    fr = synth_l2.make_l2_sequence(**kw)
    ref = reftap.l2_frames(fr, mode=1 if kw.get("nbits") in AM_BITS else 0)
    out[name] = l2_digest(ref.records)
    print(name, len(out[name]))
with open(os.path.join(ROOT, "tests", "golden", "l2.json"), "w") as f:
    json.dump(out, f, separators=(",", ":"))
