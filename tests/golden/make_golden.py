"""Regenerates tests/golden/*.json by running the UNMODIFIED reference
(oracle/_ref/libnrsc5_ref.so, built by oracle/Makefile from /root/reference)
on support/sample.xz and on the synthetic parity matrix.  Run in the build
container only (needs /root/reference); the JSON files are committed.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402
import common  # noqa: E402
import reftap  # noqa: E402
from nrsc5_b200 import synth  # noqa: E402


def main():
    raw = common.load_sample()
    log = reftap.decode(raw, want_soft=True)
    soft = b"".join(p["soft"].tobytes() for p in log.of(reftap.REC_SOFT_PM))
    hdc = b"".join(p["data"] for p in log.of(reftap.REC_HDC))
    out = {
        "source": "support/sample.xz decoded by the unmodified reference (fftshim FFT, SSE Viterbi)",
        "input_fnv": common.fnv1a32(raw[:1 << 20].tobytes()),
        "events": common.summarize(log),
        "soft_blocks": len(log.of(reftap.REC_SOFT_PM)),
        "soft_fnv": common.fnv1a32(soft),
        "hdc_packets": len(log.of(reftap.REC_HDC)),
        "hdc_bytes": len(hdc),
        "hdc_fnv": common.fnv1a32(hdc),
    }
    json.dump(out, open(os.path.join(HERE, "sample_xz.json"), "w"), indent=0)
    cases = {}
    for name, kw in common.SYNTH_CASES.items():
        cap = synth.make_fm_mp1(**kw)
        log = reftap.decode(cap.cu8, want_soft=True)
        soft = b"".join(p["soft"].tobytes() for p in log.of(reftap.REC_SOFT_PM))
        cases[name] = {
            "kwargs": kw,
            "input_fnv": common.fnv1a32(cap.cu8[:1 << 20].tobytes()),
            "events": common.summarize(log),
            "soft_blocks": len(log.of(reftap.REC_SOFT_PM)),
            "soft_fnv": common.fnv1a32(soft),
            "generated_p1_fnv": [common.fnv1a32(synth.pack_bits(b)) for b in cap.p1_frames],
        }
        print(name, len(cases[name]["events"]))
    json.dump(cases, open(os.path.join(HERE, "synth_fm.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
