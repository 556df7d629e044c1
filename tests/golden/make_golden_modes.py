"""Regenerates tests/golden/synth_fm_modes.json: synthetic FM captures in the service modes MP2, MP5, MP6 and MP11
(common.FM_MODE_CASES) decoded by the UNMODIFIED reference (oracle/_ref/libnrsc5_ref.so).  Build container only;
the JSON is committed.

    python tests/golden/make_golden_modes.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import common  # noqa: E402
import reftap  # noqa: E402
from nrsc5_b200 import synth  # noqa: E402


def main():
    out = {}
    for name, kw in common.FM_MODE_CASES.items():
        cap = synth.make_fm(**kw)
        log = reftap.decode(cap.cu8, want_soft=True)
        soft = b"".join(p["soft"].tobytes() for p in log.of(reftap.REC_SOFT_PM))
        out[name] = {
            "kwargs": kw,
            "input_fnv": common.fnv1a32(cap.cu8[:1 << 20].tobytes()),
            "events": common.summarize(log),
            "soft_blocks": len(log.of(reftap.REC_SOFT_PM)),
            "soft_fnv": common.fnv1a32(soft),
            "generated_p1_fnv": [common.fnv1a32(synth.pack_bits(b)) for b in cap.p1_frames],
            "generated_p3_fnv": [common.fnv1a32(synth.pack_bits(b)) for b in cap.p3_frames],
            "generated_p4_fnv": [common.fnv1a32(synth.pack_bits(b)) for b in cap.p4_frames],
        }
        kinds = [e[0] + (str(e[1]) + "/" + str(e[2]) if e[0] == "F" else "") for e in out[name]["events"]]
        print(name, {k: kinds.count(k) for k in sorted(set(kinds))})
    json.dump(out, open(os.path.join(HERE, "synth_fm_modes.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
