"""Regenerates tests/golden/api_sample_xz.json: support/sample.xz pushed through the PUBLIC API
(nrsc5_open_pipe / nrsc5_pipe_samples_cu8 in 32768-byte calls, as the reference CLI does) of the UNMODIFIED
reference library oracle/_ref/libnrsc5_ref.so.  Build container only; the JSON is committed.

    python tests/golden/make_golden_api.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import common  # noqa: E402
import nrsc5_api  # noqa: E402


def main():
    raw = common.load_sample()
    ev = nrsc5_api.run(os.path.join(ROOT, "oracle", "_ref", "libnrsc5_ref.so"), raw.tobytes())
    kinds = {}
    for e in ev:
        kinds[e[0]] = kinds.get(e[0], 0) + 1
    out = {"source": "support/sample.xz through the public API of the unmodified reference, 32768-byte pushes",
           "input_fnv": common.fnv1a32(raw[:1 << 20].tobytes()), "counts": kinds, "events": ev}
    json.dump(out, open(os.path.join(HERE, "api_sample_xz.json"), "w"), indent=0)
    print(kinds)


if __name__ == "__main__":
    main()
