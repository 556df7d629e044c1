"""Diagnostics for the fast P1 Viterbi: fallback counts on clean / noisy / random frames."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from nrsc5_b200 import engine as eng, synth

L = 146176
rng = np.random.default_rng(1)
bits = rng.integers(0, 2, L, dtype=np.uint8)
enc = synth.conv_encode_tb(bits).reshape(-1)
cases = {}
if enc is not None:
    e = (2.0 * enc.astype(np.float32) - 1.0)
    punct = np.ones(3 * L, dtype=bool); punct[5::6] = False
    for name, sigma, amp in [("clean127", 0.0, 127), ("mild", 0.5, 60), ("noisy", 1.0, 40)]:
        x = e * amp + rng.normal(0, sigma * amp, e.size)
        x = np.clip(np.round(x), -127, 127).astype(np.int8)
        x[~punct] = 0
        cases[name] = x
cases["random"] = rng.integers(-127, 128, 3 * L).astype(np.int8)
for name, x in cases.items():
    out, fb = eng.viterbi_k7(np.stack([x, x]), L, want_fallbacks=True)
    print(name, "fallbacks", fb, "of 2", flush=True)
