"""Developer aid: run one capture through the engine and the oracle and print
where they diverge (block params, soft bits, PDUs)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import nrsc5_b200
from nrsc5_b200 import engine as eng, synth
import port, reftap, common

name = sys.argv[1] if len(sys.argv) > 1 else "mp1_clean"
if name == "sample":
    cu8 = common.load_sample()
else:
    cu8 = synth.make_fm_mp1(**common.SYNTH_CASES[name]).cu8
ref = port.decode(cu8, want_soft=True, want_blocks=True)
with nrsc5_b200.Engine(nstreams=1, input_capacity=cu8.size + 4096, log_capacity=64 << 20, emit_soft=True) as e:
    e.push_cu8(0, cu8[: cu8.size & ~3])
    e.process()
    recs = e.drain(0)
    st = e.stats()
    print("stats blocks", st.blocks, "frames", st.p1_frames, "launches", st.kernel_launches)
rb = [p for t, p in ref.records if t == reftap.REC_BLOCK]
gb = [p for t, p in recs if t == eng.REC_BLOCK]
print("blocks ref/gpu", len(rb), len(gb))
for i, (a, b) in enumerate(zip(rb, gb)):
    flag = "" if (a["state"], a["samperr"], a["cfo"], a["start"]) == (b["state"], b["samperr"], b["cfo"], b["start"]) else " <<<"
    if i < 12 or flag:
        print(i, "ref", a["state"], a["samperr"], round(a["angle"], 6), a["cfo"], a["start"], "| gpu", b["state"], b["samperr"], round(b["angle"], 6), b["cfo"], b["start"], flag)
sa = [p["soft"] for t, p in recs if t == eng.REC_SOFT_PM]
sb = [p["soft"] for p in ref.of(reftap.REC_SOFT_PM)]
print("soft blocks", len(sa), len(sb))
for i, (x, y) in enumerate(zip(sa, sb)):
    d = x.astype(int) - y.astype(int)
    if i < 6 or np.abs(d).max() > 1:
        print(" soft blk", i, "ndiff", int((d != 0).sum()), "max", int(np.abs(d).max()))
m = {1: "F", 2: "P", 3: "S", 4: "L", 5: "M", 6: "B"}
gk = "".join(m[t] for t, _ in recs if t in m)
rk = "".join(e[0] for e in common.summarize(ref))
print("kinds equal", gk == rk)
if gk != rk:
    print(gk); print(rk)
gp1 = [r["bits"] for t, r in recs if t == eng.REC_FRAME]
print("P1 equal", gp1 == ref.p1_frames, len(gp1), len(ref.p1_frames))
for i, (a, b) in enumerate(zip(gp1, ref.p1_frames)):
    if a != b:
        print("  frame", i, "bit diffs", bin(int.from_bytes(a, "big") ^ int.from_bytes(b, "big")).count("1"))
gpi = [r["bits"] for t, r in recs if t == eng.REC_PIDS]
print("PIDS equal", gpi == ref.pids_frames, len(gpi), len(ref.pids_frames), sum(a != b for a, b in zip(gpi, ref.pids_frames)))
for k, t in [("S", eng.REC_SYNC), ("M", eng.REC_MER), ("B", eng.REC_BER)]:
    print(k, [r for tt, r in recs if tt == t][:4], [p for tt, p in ref.records if tt == t][:4])
dbg = [(t, r["dbg"]) for t, r in recs if t in (10, 11)]
if dbg:
    print("debug records:")
    for i in range(0, min(len(dbg), 80), 2):
        print("  ", dbg[i:i + 2])
