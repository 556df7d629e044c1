"""Diagnostic: the dropout capture of tests/test_gpu_edge.py decoded in one push and in 3 MiB pushes, several times;
prints where the per-block records (REC_BLOCK: state, timing, NCO, window position) of two runs first differ."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import nrsc5_b200
from nrsc5_b200 import engine as eng, synth

a = synth.make_fm_mp1(nframes=2, seed=501, lead_in=700, cfo_hz=80.0, noise_lsb=3.0, tail_blocks=3)
b = synth.make_fm_mp1(nframes=2, seed=502, lead_in=1333, cfo_hz=-640.0, noise_lsb=3.0, start_bc=9, tail_blocks=2)
rng = np.random.default_rng(503)
gap = np.clip(np.rint(rng.standard_normal(2 * 600000) * 6 + 127), 0, 255).astype(np.uint8)
cu8 = np.concatenate([a.cu8, gap, b.cu8])
cu8 = cu8[: cu8.size & ~3]

def run(chunk):
    with nrsc5_b200.Engine(nstreams=1, input_capacity=cu8.size + 4096, log_capacity=16 << 20) as e:
        out = []
        if chunk is None:
            e.push_cu8(0, cu8); e.process(); out = e.drain(0)
        else:
            for off in range(0, cu8.size, chunk):
                e.push_cu8(0, cu8[off: off + chunk]); e.process(); out += e.drain(0)
    return out

def blocks(recs):
    return [(r["state"], r["samperr"], round(r["angle"], 5), r["cfo"], r["start"]) for t, r in recs if t == eng.REC_BLOCK]

def kinds(recs):
    m = {1: "F", 2: "P", 3: "S", 4: "L", 5: "M", 6: "B"}
    return "".join(m.get(t, "") for t, _ in recs)

base = run(None)
print("single push:", len(blocks(base)), "blocks", kinds(base).count("P"), "PIDS", kinds(base).count("S"), "SYNC")
for trial, chunk in enumerate([None, 3 << 20, 3 << 20, 1 << 20, 276480 * 2, 3 << 20]):
    r = run(chunk)
    bb, b0 = blocks(r), blocks(base)
    k = next((i for i, (x, y) in enumerate(zip(bb, b0)) if x != y), None)
    print("chunk", chunk, ":", len(bb), "blocks", kinds(r).count("P"), "PIDS; first differing block:", k,
          (bb[k - 1: k + 2], b0[k - 1: k + 2]) if k is not None else "", "kinds equal:", kinds(r) == kinds(base))
