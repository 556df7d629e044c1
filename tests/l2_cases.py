"""The L2 framing parity matrix: name -> synth_l2.make_l2_sequence kwargs (+ the reference mode)."""
L2_CASES = {
    "p1_fm": dict(),
    "p1_fm_fixed": dict(fixed=True, nframes=40),
    "p1_fm_fixed_b": dict(fixed=True, nframes=36, seed=11),
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


def mutated_sequence(trial: int):
    """A generated PDU sequence with up to five random byte errors per frame (most of them in the first 400 bytes,
    where headers, locations, HEF and PSD live): exercises the early-return / resynchronisation branches of
    frame_process.  Returns (frames, am)."""
    import numpy as np
    from nrsc5_b200 import synth_l2
    rng = np.random.default_rng(1000 + trial)
    nbits = [4608, 146176, 3750, 24000, 4608, 2304][trial % 6]
    fixed = (trial % 3 == 1) and nbits not in (2304, 3750)
    fr = synth_l2.make_l2_sequence(seed=500 + trial, nframes=10 if nbits == 146176 else 24, nbits=nbits, fixed=fixed, lc=trial % 3)
    out = []
    for f in fr:
        if f is None:
            out.append(None)
            continue
        b = bytearray(f[2])
        for _ in range(int(rng.integers(0, 6))):
            pos = int(rng.integers(0, min(len(b), 400) if rng.random() < 0.7 else len(b)))
            b[pos] ^= int(rng.integers(1, 256))
        out.append((f[0], f[1], bytes(b)))
    return out, nbits in AM_BITS


def malformed_sequences():
    """Two PDU sequences on which the reference's frame.c leaves its buffers (so it is not run on them): a CCC that
    passes its FCS-16 but announces more subchannel bytes than a PDU holds (frame.c:493-496 would read in front of
    the buffer), and header expansion fields that run past la_location (frame.c:608: the byte count wraps).  The
    engine and the oracle define the same safe outcome for both."""
    import numpy as np
    from nrsc5_b200 import synth_l2 as g
    rng = np.random.default_rng(99)
    n = g.pdu_len(4608)
    # (a) CCC with a 60 000-byte subchannel
    src = g.FixedDataSource(rng, 4, [])
    src.ccc_stream = bytearray(b"\x7e" * 12 + g.hdlc_frame(bytes([0, 0, 0, 0x60, 0xEA])) + b"\x7e")
    a = [None]
    for f in range(12):
        part = g.audio_pdu(rng, [40, 41, 42], seq=f, psd=g.hdlc_frame(bytes([0x21, f]) + bytes(20)))
        a.append((1, 4608, g.frame_from_pdu(g.fill_pdu(rng, [part], n, src.tail()), 4608, g.PCI_AUDIO_FIXED)))
    # (b) la_location inside the header expansion fields
    b = [None]
    for f in range(6):
        hef = g.hef_fields(1, ptype=22, pdu_len_field=300, with_loc=True)
        pdu = bytearray(g.audio_pdu(rng, [50, 60, 70], codec=0, stream=0, seq=f, hef=hef, psd=g.hdlc_frame(bytes([0x21]) + bytes(30))))
        if f % 2:
            pdu[13] = 14 + 6 + 2                       # 3 locations of 16 bit = 6 bytes; la points into the HEF
            head = bytearray(pdu[:96])
            g.protect_header(head)
            pdu[:96] = head
        b.append((1, 4608, g.frame_from_pdu(g.fill_pdu(rng, [bytes(pdu)], n), 4608, g.PCI_AUDIO)))
    return {"ccc_overlong": a, "hef_past_la": b}


def stress_sequences():
    """PDU sequences that reach the rarely taken branches of the L2 kernel (the reference handles all of them, so the
    oracle stays pinned to it): more packets in one frame than the kernel's shared-memory packet table holds (1024;
    thread 0 then checks the CRC-8 itself), a PSD byte stream that overruns the 8212-byte HDLC buffer before the next
    flag (frame.c:381-386), and - `ev_overflow` - more events than the kernel's per-frame staging area holds."""
    import numpy as np
    from nrsc5_b200 import synth_l2 as g
    n = g.pdu_len(146176)

    def frame(rng, npdus, psd_of, npk=63):
        parts = [g.audio_pdu(rng, [1] * npk, codec=0, stream=0, seq=k & 63, psd=psd_of(k)) for k in range(npdus)]
        return (0, 146176, g.frame_from_pdu(g.fill_pdu(rng, parts, n), 146176, g.PCI_AUDIO))

    out = {}
    rng = np.random.default_rng(5)
    out["many_packets"] = [None] + [frame(rng, 36, lambda k: b"") for _ in range(2)]
    rng = np.random.default_rng(6)
    filler = bytes(b for b in rng.integers(0, 256, 260, dtype=np.uint8).tobytes() if b != 0x7E)[:210]
    msg = g.hdlc_frame(bytes([0x21]) + bytes(range(40)))
    # frame 0: a flag, then 60 PDUs x 210 flag-free PSD bytes = 12 600 > 8212: the buffer overruns and the scanner
    # waits for the next flag; frame 1: messages again
    out["hdlc_overrun"] = [None,
                           frame(rng, 60, lambda k: (b"\x7e" if k == 0 else b"") + filler, npk=10),
                           frame(rng, 40, lambda k: msg, npk=10)]
    rng = np.random.default_rng(7)
    out["ev_overflow"] = [None, frame(rng, 60, lambda k: b"")]
    return out
