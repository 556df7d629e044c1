"""Cycle counts of the AM K=9 decoder's two phases on the GPU (stage entry, 256 frames at once = config 4's stream count)."""
import os, sys, ctypes
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from nrsc5_b200 import engine as eng
import test_gpu_stages as st
eng.load_library().nrsc5b_debug_set(8)
rng = np.random.default_rng(3)
for n, gens in ((3750, (0o561, 0o657, 0o711)), (24000, (0o561, 0o753, 0o711))):
    _, sym = st._k9_frames(rng, 4, n, gens, 0.01)
    sym = np.tile(sym, (64, 1))
    for wu in (128, 16):
        print("len", n, "jobs", sym.shape[0], "warmup", wu, flush=True)
        got, rounds, redone = eng.viterbi_k9(sym, gens, warmup=wu)
