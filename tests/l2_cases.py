"""The L2 framing parity matrix: name -> synth_l2.make_l2_sequence kwargs (+ the reference mode)."""
L2_CASES = {
    "p1_fm": dict(),
    "p1_fm_fixed": dict(fixed=True, nframes=40),
    "p1_fm_fixed_b": dict(fixed=True, nframes=20, seed=11),
    "p3_mp3": dict(nbits=4608, lc=1),
    "p4_mp11_fixed": dict(nbits=4608, lc=2, fixed=True, nframes=60, seed=3),
    "p3_mp2": dict(nbits=2304, lc=1),
    "p1_am": dict(nbits=3750),
    "p3_ma1_fixed": dict(nbits=24000, lc=1, fixed=True, nframes=30),
    "p3_ma3": dict(nbits=30000, lc=1),
}
AM_BITS = (3750, 24000, 30000)


def l2_digest(records):
    """Order-preserving digest of an L2 record stream (frames and the L2 -> L3 calls they cause)."""
    from common import fnv1a32
    out = []
    for ty, r in records:
        if ty == 1:
            out.append(["F", r["lc"], r["nbits"], fnv1a32(r["bits"])])
        elif ty == 16:
            out.append(["V"] + [r[k] for k in ("program", "access", "type", "codec_mode", "blend_control", "gain", "common_delay", "latency")])
        elif ty == 17:
            out.append(["A", r["program"], r["stream_id"], r["offset"]])
        elif ty == 18:
            out.append(["S", len(r["data"]), fnv1a32(r["data"])])
        elif ty == 19:
            out.append(["K", r["program"], r["stream_id"], r["seq"], r["shape"], r["flags"], r["size"], fnv1a32(r["data"])])
    return out
