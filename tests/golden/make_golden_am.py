"""Regenerates tests/golden/synth_am.json: the synthetic AM MA1 / MA3 captures (common.AM_CASES, cs16 at 46 511.72 S/s)
decoded by the UNMODIFIED reference (oracle/_ref/libnrsc5_ref.so, AM mode).  Build container only; the JSON is
committed.  It is the known answer the AM rows of the scope table will be held to.

    python tests/golden/make_golden_am.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import common  # noqa: E402
import reftap  # noqa: E402
from nrsc5_b200 import synth_am  # noqa: E402


def main():
    out = {}
    for name, kw in common.AM_CASES.items():
        cap = synth_am.make_am_ma1(**kw)
        log = reftap.decode(cap.cs16, mode=reftap.MODE_AM)
        out[name] = {"kwargs": kw, "input_fnv": common.fnv1a32(cap.cs16[:1 << 18].tobytes()),
                     "events": common.summarize(log)}
        kinds = [e[0] + (str(e[1]) if e[0] == "F" else "") for e in out[name]["events"]]
        print(name, {k: kinds.count(k) for k in sorted(set(kinds))})
    json.dump(out, open(os.path.join(HERE, "synth_am.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
