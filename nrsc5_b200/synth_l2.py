"""Generator of L1 PDUs with real L2 content, for the parity tests of the L2 framing row.

The reference ships no transmitter; the layouts below are read off its receiver,
reference src/frame.c:
  * audio PDU: 8 RS parity bytes + header fields (`parse_header`, :181-196), packet locations
    (`parse_location`, :315-326, 12 or 16 bit wide by codec mode / stream, :267-287), optional header
    expansion fields (`parse_hef`, :198-265), PSD bytes (HDLC, `parse_hdlc` :369-391, `aas_push` :343-367)
    up to `la_location`, then the packets, each closed by its CRC-8 (:130-136, :617-620);
  * the first 96 bytes are one shortened RS(255,247) codeword (`fix_header`, :158-179);
  * fixed data at the tail of a PDU: sync byte, CCC bytes, subchannel bytes (`process_fixed_data`, :458-514),
    subchannel blocks of 4 marker bytes + 255 payload bytes carrying HDLC frames (:441-446);
  * PCI bits spread over the frame and the per-byte bit reversal (`frame_push`, :645-714).
"""
from __future__ import annotations

import numpy as np

from .synth import rs_parity

PCI_AUDIO, PCI_AUDIO_OPP, PCI_AUDIO_FIXED, PCI_AUDIO_FIXED_OPP, PCI_FIXED = 0x38D8D3, 0xCE3634, 0xE3634C, 0x8D8D33, 0x3634CE

# frame length -> (first PCI bit, PCI spacing, PCI bits), reference src/frame.c:651-690
FRAME_GEOMETRY = {146176: (146176 - 30000, 1248, 24), 4608: (120, 184, 24), 2304: (120, 88, 24),
                  3750: (120, 160, 22), 24000: (120, 992, 24), 30000: (120, 1240, 24)}


def pdu_len(nbits: int) -> int:
    return (nbits - FRAME_GEOMETRY[nbits][2]) // 8


def crc8(data: bytes) -> int:
    c = 0xFF
    for b in data:
        c ^= b
        for _ in range(8):
            c = ((c << 1) ^ 0x31) & 0xFF if c & 0x80 else (c << 1) & 0xFF
    return c


def fcs16(data: bytes) -> int:
    c = 0xFFFF
    for b in data:
        c ^= b
        for _ in range(8):
            c = (c >> 1) ^ 0x8408 if c & 1 else c >> 1
    return c


def hdlc_frame(payload: bytes, good_fcs: bool = True) -> bytes:
    """Opening flag + escaped (payload + FCS-16); the closing flag is the next frame's opening one."""
    f = fcs16(payload) ^ 0xFFFF
    if not good_fcs:
        f ^= 0x0100
    body = payload + bytes([f & 0xFF, f >> 8])
    out = bytearray([0x7E])
    for b in body:
        if b in (0x7E, 0x7D):
            out += bytes([0x7D, b ^ 0x20])
        else:
            out.append(b)
    return bytes(out)


def protect_header(pdu96: bytearray) -> None:
    """Fill bytes 0..7 with the RS(255,247) parity of bytes 8..95 (block[254 - i] = buf[i])."""
    msg = bytes(159) + bytes(pdu96[95 - k] for k in range(88))
    par = rs_parity(msg)
    for k in range(8):
        pdu96[7 - k] = par[k]


def pack_locations(locs, width: int) -> bytes:
    n = len(locs)
    out = bytearray((width * n + 4) // 8)
    if width == 16:
        for j, v in enumerate(locs):
            out[2 * j] = v & 0xFF
            out[2 * j + 1] = v >> 8
    else:
        for j, v in enumerate(locs):
            q = (j // 2) * 3
            if j % 2 == 0:
                out[q] = v & 0xFF
                out[q + 1] |= (v >> 8) & 0xF
            else:
                out[q + 1] |= (v & 0xF) << 4
                out[q + 2] = v >> 4
    return bytes(out)


def loc_width(codec: int, stream: int) -> int:
    if codec in (1, 2, 3):
        return 12 if stream == 0 else 16
    return 12 if codec in (10, 13) else 16


def audio_pdu(rng, sizes, codec=0, stream=0, pdu_seq=0, blend=2, delay=0, common=24, latency=4, pfirst=0, plast=0,
              seq=0, hef: bytes = b"", psd: bytes = b"", bad_crc=(), header_errors=0, min_len=110) -> bytes:
    """One audio PDU holding len(sizes) packets of the given payload sizes."""
    nop = len(sizes)
    width = loc_width(codec, stream)
    locb = (width * nop + 4) // 8
    la = 14 + locb + len(hef) + len(psd) - 1
    pad = 0
    total = la + 1 + sum(s + 1 for s in sizes)
    if total < min_len:                       # the walk needs more than 96 bytes; pad the PSD region with idle flags
        pad = min_len - total
        la += pad
    assert la < 256
    locs, pos, body = [], la + 1, bytearray()
    for j, s in enumerate(sizes):
        pay = rng.integers(0, 256, s, dtype=np.uint8).tobytes()
        c = crc8(pay)
        if j in bad_crc:
            c ^= 0x5A
        body += pay + bytes([c])
        pos += s + 1
        locs.append(pos - 1)
    fields = bytearray(6)
    fields[0] = (codec & 15) | ((stream & 3) << 4) | ((pdu_seq & 3) << 6)
    fields[1] = ((pdu_seq >> 2) & 1) | ((blend & 3) << 1) | ((delay & 31) << 3)
    fields[2] = (common & 0x3F) | ((latency & 3) << 6)
    fields[3] = ((latency >> 2) & 1) | ((pfirst & 1) << 1) | ((plast & 1) << 2) | ((seq & 31) << 3)
    fields[4] = ((seq >> 5) & 1) | ((nop & 0x3F) << 1) | ((1 if hef else 0) << 7)
    fields[5] = la
    pdu = bytearray(8) + fields + pack_locations(locs, width) + hef + psd + bytes([0x7E]) * pad + body
    head = bytearray(pdu[:96])
    protect_header(head)
    pdu[:96] = head
    for k in range(header_errors):
        pdu[3 + 9 * k] ^= 0x41 + k
    return bytes(pdu)


def hef_fields(prog: int, access: int = 0, ptype: int = 0, pdu_len_field: int | None = None, with_loc: bool = False,
               marker: int | None = None) -> bytes:
    """Header expansion: class indicator, program number (+ optional PDU length), access / program type,
    optionally a type-3 field and a type-4 field with a PDU marker (frame.c:198-265)."""
    out = [0x80 | 0x00 | 0x1]
    if pdu_len_field is None:
        out.append(0x80 | 0x10 | (prog << 1))
    else:
        out += [0x10 | (prog << 1) | 1, (pdu_len_field >> 7) & 0x7F, 0x80 | (pdu_len_field & 0x7F)]
    if with_loc:
        out += [0x30 | 0x8, 0x11, 0x22, 0x33, 0x80 | 0x44]
    if marker is not None:
        out += [0x40 | 0x8 | 0x5, (marker >> 14) & 0x7F, (marker >> 7) & 0x7F, 0x80 | (marker & 0x7F)]
    out += [0x20 | (access << 3) | (ptype >> 7), ptype & 0x7F]
    return bytes(out)


def frame_from_pdu(pdu: bytes, nbits: int, pci: int) -> bytes:
    """PDU bytes + PCI -> the frame's bits as frame_push() receives them, packed MSB first."""
    first, step, npci = FRAME_GEOMETRY[nbits]
    assert len(pdu) == (nbits - npci) // 8
    logical = np.zeros(nbits, dtype=np.uint8)
    pos = first + step * np.arange(npci)
    is_pci = np.zeros(nbits, dtype=bool)
    is_pci[pos] = True
    logical[pos] = [(pci >> (23 - h)) & 1 for h in range(npci)]
    data = np.unpackbits(np.frombuffer(pdu, dtype=np.uint8))
    rest = np.flatnonzero(~is_pci)
    logical[rest[:data.size]] = data
    i = np.arange(nbits)
    base = i & ~7
    span = np.minimum(nbits - base, 8)
    bits = np.zeros(nbits, dtype=np.uint8)
    bits[base + span - 1 - (i & 7)] = logical
    return np.packbits(bits).tobytes()


def fill_pdu(rng, parts, nbytes: int, tail: bytes = b"") -> bytes:
    """Audio PDUs back to back, random filler up to the fixed-data tail."""
    body = b"".join(parts)
    gap = nbytes - len(body) - len(tail)
    assert gap >= 0, (len(body), len(tail), nbytes)
    fill = rng.integers(0, 256, gap, dtype=np.uint8)
    if gap and not tail:
        fill[-1] = 0x12                        # unequal nibbles: no fixed-data sync (frame.c:448-456)
    return body + fill.tobytes() + tail


class FixedDataSource:
    """Fixed-data tail of successive PDUs: sync byte of width `width`, a CCC announcing the subchannel lengths,
    and per subchannel a stream of marker + 255-byte blocks that carry HDLC frames."""

    def __init__(self, rng, width: int, lengths, messages_per_sub=4, misalign: int = 0):
        self.width = width
        self.lengths = list(lengths)
        ccc = bytes([0x00]) + b"".join(bytes([0, 0, l & 0xFF, l >> 8]) for l in self.lengths)
        # the receiver only scans the CCC bytes once it has seen the same sync width three times in a row
        # (frame.c:466-477) and forgets everything at a frame_reset: idle flags first, and the CCC keeps repeating
        self.ccc_stream = bytearray((b"\x7e" * (3 * width) + hdlc_frame(ccc) + b"\x7e") * 64)
        self.ccc_pos = 0
        self.sent = []
        self.sub_streams = []
        for si, _ in enumerate(self.lengths):
            hd = bytearray()
            msgs = []
            for m in range(messages_per_sub):
                body = bytes([0x21]) + rng.integers(0, 256, 40 + 37 * m + 11 * si, dtype=np.uint8).tobytes()
                good = not (m == 1 and si == 0)
                hd += hdlc_frame(body, good_fcs=good)
                if good:
                    msgs.append(body[1:])
            hd += b"\x7e"
            while len(hd) % 255:
                hd += b"\x7e"
            blocks = bytearray(rng.integers(0, 256, misalign, dtype=np.uint8).tobytes())
            for _ in range(6):                 # the receiver joins mid-stream (after the CCC): the messages keep coming
                for k in range(0, len(hd), 255):
                    blocks += bytes([0x7D, 0x3A, 0xE2, 0x42]) + hd[k:k + 255]
            self.sub_streams.append(blocks)
            self.sent.append(msgs)
        self.sub_pos = [0] * len(self.lengths)

    def tail(self) -> bytes:
        out = bytearray()
        for si, l in enumerate(self.lengths):
            s = self.sub_streams[si]
            chunk = bytes(s[self.sub_pos[si]:self.sub_pos[si] + l])
            chunk += bytes([0x7E]) * (l - len(chunk))
            self.sub_pos[si] += l
            out += chunk
        c = bytes(self.ccc_stream[self.ccc_pos:self.ccc_pos + self.width])
        c += bytes([0x7E]) * (self.width - len(c))
        self.ccc_pos += self.width
        out += c
        out.append(((self.width // 2) << 4) | (self.width // 2))
        return bytes(out)


def make_l2_sequence(seed: int = 7, nframes: int = 12, nbits: int = 146176, fixed: bool = False, lc: int = 0):
    """A frame_reset followed by `nframes` L1 PDUs of `nbits` bits carrying two programs (one with an enhanced
    stream), PSD messages with escapes and a bad FCS, header expansion fields, both location widths, corrected and
    uncorrectable headers, packets with CRC errors and, with `fixed`, a fixed-data tail.  Returns the frame list
    [(lc, nbits, packed bits) | None] for reftap.l2_frames / port.l2_frames / Engine.l2_frames."""
    rng = np.random.default_rng(seed)
    n = pdu_len(nbits)
    src = FixedDataSource(rng, 4, [37, 61], misalign=3) if fixed else None
    psd_msgs = [hdlc_frame(bytes([0x21]) + bytes([0x7E, 0x7D, 0x00, 0x51]) + rng.integers(0, 256, 60, dtype=np.uint8).tobytes()),
                hdlc_frame(bytes([0x21]) + rng.integers(0, 256, 33, dtype=np.uint8).tobytes(), good_fcs=False),
                hdlc_frame(bytes([0x22]) + rng.integers(0, 256, 20, dtype=np.uint8).tobytes()),
                hdlc_frame(b""), hdlc_frame(bytes([0x21]) + rng.integers(0, 256, 150, dtype=np.uint8).tobytes())]
    psd_stream = b"".join(psd_msgs) * 3 + b"\x7e"
    psd_pos = 0
    frames = [None]
    small = n < 4000
    for f in range(nframes):
        tail = src.tail() if src else b""
        parts = []
        room = n - len(tail) - 8
        take = psd_stream[psd_pos:psd_pos + 45]
        psd_pos = (psd_pos + 45) % max(1, len(psd_stream) - 45)
        hdr_err = 5 if f == 5 else (f % 5 if f != 7 else 4)
        if small:
            sizes = [int(x) for x in rng.integers(20, 60, 3)]
            parts.append(audio_pdu(rng, sizes, codec=0, stream=0, pdu_seq=f & 7, latency=4, seq=(3 * f) & 63, psd=take,
                                   pfirst=f & 1, plast=(f >> 1) & 1, hef=hef_fields(0, ptype=5), header_errors=hdr_err,
                                   bad_crc=(1,) if f == 3 else ()))
            if f % 3 == 1 and n > 400:
                parts.append(audio_pdu(rng, [30, 31], codec=13, stream=0, pdu_seq=f & 7, seq=f & 63,
                                       hef=hef_fields(2, access=1, ptype=130, marker=0x12345)))
        else:
            sizes = [int(x) for x in rng.integers(100, 400, 28 + (f % 5))]
            parts.append(audio_pdu(rng, sizes, codec=0, stream=0, pdu_seq=f & 7, latency=4, seq=(29 * f) & 63, psd=take,
                                   pfirst=f & 1, plast=(f >> 1) & 1, hef=hef_fields(0, ptype=5) if f % 4 else b"",
                                   header_errors=hdr_err, bad_crc=(2, 9) if f == 3 else ()))
            sizes = [int(x) for x in rng.integers(40, 120, 8)]
            parts.append(audio_pdu(rng, sizes, codec=2 if f < 8 else 10, stream=0, pdu_seq=(f + 1) & 7, latency=2, seq=(8 * f) & 63,
                                   blend=1, delay=17 + (f > 6), common=11,
                                   hef=hef_fields(1, access=f > 9, ptype=22, pdu_len_field=700 + f, with_loc=True),
                                   psd=hdlc_frame(bytes([0x21, f]) + bytes(range(30)))))
            sizes = [int(x) for x in rng.integers(30, 90, 5)]
            parts.append(audio_pdu(rng, sizes, codec=2, stream=1, pdu_seq=f & 7, seq=(5 * f) & 63,
                                   hef=hef_fields(1, ptype=22)))
            if f == 2:
                parts.append(audio_pdu(rng, [50, 60], codec=0, stream=3, hef=b""))          # invalid stream id: skipped
                parts.append(audio_pdu(rng, [70], codec=0, stream=0, hef=hef_fields(3)))
            if f == 6:                                                                           # locations out of order
                bad = bytearray(audio_pdu(rng, [50, 60, 70], codec=0, stream=0))
                bad[14], bad[16] = bad[16], bad[14]
                bad[15], bad[17] = bad[17], bad[15]
                head = bytearray(bad[:96]); protect_header(head); bad[:96] = head
                parts.append(bytes(bad))
        assert sum(map(len, parts)) <= room
        pdu = fill_pdu(rng, parts, n, tail)
        pci = (PCI_AUDIO_FIXED if f % 2 else PCI_AUDIO_FIXED_OPP) if fixed else (PCI_AUDIO if f % 2 else PCI_AUDIO_OPP)
        if f == 9 and not fixed:
            pci = PCI_FIXED                                                                      # no audio in this one
        frames.append((lc, nbits, frame_from_pdu(pdu, nbits, pci)))
        if f == 8:
            frames.append(None)                                                                  # re-entering fine sync
    return frames
