"""GPU parity tests, stage by stage, through the C ABI (libnrsc5_b200.so)
against the CPU oracle (oracle/nrsc5_oracle.c)."""
import numpy as np
import pytest

import port
from nrsc5_b200 import engine as eng
from nrsc5_b200 import synth

pytestmark = pytest.mark.gpu


def test_halfband_bit_exact():
    rng = np.random.default_rng(0)
    cu8 = rng.integers(0, 256, 4 * 200000, dtype=np.uint8)
    cu8[:4000] = 255
    cu8[4000:8000] = 0
    cu8[8000:12000:2] = 255
    assert np.array_equal(eng.halfband_fm(cu8), port.halfband_fm(cu8))


def test_fft2048_matches_fp64():
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((8, 2048)) + 1j * rng.standard_normal((8, 2048))).astype(np.complex64)
    x[0] = 0
    x[0, 1] = 1
    got = eng.fft2048(x)
    ref = np.fft.fft(x.astype(np.complex128), axis=1)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < 2e-6, err        # fp32 FFT: ~log2(N)*eps relative


@pytest.mark.parametrize("length", [80, 2304, 4608])
def test_viterbi_random_soft_bit_exact(length):
    rng = np.random.default_rng(length)
    nframes = 6
    soft = rng.integers(-127, 128, (nframes, 3 * length), dtype=np.int8)
    soft[:, 5::6] = 0
    got = eng.viterbi_k7(soft, length)
    for f in range(nframes):
        assert np.array_equal(got[f], port.viterbi(soft[f]))


def test_viterbi_saturating_and_noisy():
    rng = np.random.default_rng(3)
    length = 4608
    u = rng.integers(0, 2, length, dtype=np.uint8)
    c = synth.conv_encode_tb(u).reshape(-1).astype(np.int16)
    full = ((2 * c - 1) * 127).astype(np.int8)             # drives int16 metrics into saturation
    noisy = np.clip((2 * c - 1) * 40 + rng.normal(0, 45, c.size), -127, 127).astype(np.int8)
    soft = np.stack([full, noisy])
    got = eng.viterbi_k7(soft, length)
    assert np.array_equal(got[0], port.viterbi(full)) and np.array_equal(got[0], u)
    assert np.array_equal(got[1], port.viterbi(noisy))


def test_viterbi_fast_path_and_fallback_agree():
    """The register-resident fast path proves its result or hands the frame to the exact fallback: encoded
    (even noisy) frames stay on the fast path, the saturating frame and pure noise must fall back - and every
    result equals the sequential decoder's."""
    rng = np.random.default_rng(8)
    length = 4608
    u = rng.integers(0, 2, length, dtype=np.uint8)
    c = synth.conv_encode_tb(u).reshape(-1).astype(np.int16)
    clean = ((2 * c - 1) * 60).astype(np.int8)
    clean[5::6] = 0
    noisy = np.clip((2 * c - 1) * 40 + rng.normal(0, 40, c.size), -127, 127).astype(np.int8)
    noisy[5::6] = 0
    got, fb = eng.viterbi_k7(np.stack([clean, noisy]), length, want_fallbacks=True)
    assert fb == 0
    assert np.array_equal(got[0], u) and np.array_equal(got[1], port.viterbi(noisy))
    full = ((2 * c - 1) * 127).astype(np.int8)               # could saturate the reference's int16 metrics
    rnd = rng.integers(-127, 128, 3 * length).astype(np.int8)
    got, fb = eng.viterbi_k7(np.stack([full, rnd]), length, want_fallbacks=True)
    assert fb >= 1
    assert np.array_equal(got[0], port.viterbi(full)) and np.array_equal(got[1], port.viterbi(rnd))


def test_viterbi_p1_length_bit_exact():
    rng = np.random.default_rng(4)
    length = 146176
    u = rng.integers(0, 2, length, dtype=np.uint8)
    c = synth.conv_encode_tb(u).reshape(-1).astype(np.int16)
    soft = np.clip((2 * c - 1) * 30 + rng.normal(0, 40, c.size), -127, 127).astype(np.int8)
    soft[5::6] = 0
    got = eng.viterbi_k7(soft[None, :], length)[0]
    assert np.array_equal(got, port.viterbi(soft))


def test_rs_decode_bit_exact():
    rng = np.random.default_rng(5)
    blocks = []
    for t in range(600):
        hdr = np.frombuffer(synth.audio_pdu_header(rng=rng), dtype=np.uint8)
        blk = np.zeros(255, dtype=np.uint8)
        blk[254 - np.arange(96)] = hdr
        for p in rng.choice(255 if t % 3 == 0 else 96, t % 8, replace=False):
            blk[p if t % 3 == 0 else 254 - p] ^= rng.integers(1, 256)
        if t % 25 == 24:
            blk = rng.integers(0, 256, 255, dtype=np.uint8)
        blocks.append(blk)
    blocks = np.stack(blocks)
    rc, fixed = eng.rs_decode(blocks)
    for i in range(blocks.shape[0]):
        rc_ref, ref = port.rs_decode(blocks[i])
        assert rc[i] == rc_ref and np.array_equal(fixed[i], ref), i

def _rs_words(rng, n):
    """Valid (255,247) codewords - a generated 96-byte header behind 159 zeros, or a random message encoded by the
    synth encoder where available - hit by 5..8 symbol errors anywhere in the block: beyond the code's radius, where
    the result (rejected, or "corrected" into another word) depends on what the decoder makes of an ambiguous locator."""
    out = np.zeros((n, 255), dtype=np.uint8)
    for i in range(n):
        hdr = np.frombuffer(synth.audio_pdu_header(rng=rng), dtype=np.uint8)
        out[i, 254 - np.arange(96)] = hdr
        ne = 5 + (i & 3)
        pos = rng.choice(255 if i % 2 else np.arange(159, 255), ne, replace=False)
        out[i, pos] ^= rng.integers(1, 256, ne).astype(np.uint8)
    return out


def rs_beyond_radius(n):
    rng = np.random.default_rng(2024)
    blocks = _rs_words(rng, n)
    rc, fixed = eng.rs_decode(blocks)
    nfail = ncorr = 0
    for i in range(n):
        rc_ref, ref = port.rs_decode(blocks[i])
        assert rc[i] == rc_ref and np.array_equal(fixed[i], ref), (i, rc[i], rc_ref)
        nfail += rc_ref < 0
        ncorr += rc_ref > 0
    return nfail, ncorr


def test_rs_beyond_the_correction_radius_equals_reference():
    """120 000 words with 5..8 symbol errors: the warp decoder (inversion-free Berlekamp-Massey across lanes) must do
    exactly what decode_rs_char does with them - reject most, mis-correct the ones whose locator happens to split -
    byte for byte and count for count (the oracle's decoder is pinned to the unmodified reference's decode_rs_char
    on the same kind of words in tests/test_oracle.py)."""
    nfail, ncorr = rs_beyond_radius(120000)
    assert nfail > 100000 and ncorr > 50          # both outcomes occur



def _k9_frames(rng, njobs, n, gens, noise):
    """Tail-biting rate-1/3 K=9 code words of random bits as hard symbols; `noise` = fraction of flipped symbols; every
    fifth symbol punctured (0), as the AM chain's E1/E2/E3 depuncturing leaves them."""
    bits = rng.integers(0, 2, (njobs, n), dtype=np.uint8)
    sym = np.zeros((njobs, 3 * n), dtype=np.int8)
    for j in range(njobs):
        b = bits[j].astype(np.int64)
        reg = np.zeros(n, dtype=np.int64)
        for q in range(9):
            reg |= np.roll(b, q) << (8 - q)                                  # register at bit i: bits i-8 .. i, newest on top
        for k, g in enumerate(gens):
            v = reg & g
            par = np.zeros(n, dtype=np.int64)
            for q in range(9):
                par ^= (v >> q) & 1
            sym[j, k::3] = np.where(par == 1, 1, -1)
    flip = rng.random(sym.shape) < noise
    sym = np.where(flip, -sym, sym).astype(np.int8)
    sym[:, 4::5] = 0
    return bits, sym


@pytest.mark.gpu
@pytest.mark.parametrize("n,gens", [(80, (0o561, 0o753, 0o711)), (3750, (0o561, 0o657, 0o711)), (3751, (0o561, 0o657, 0o711)),
                                    (3752, (0o561, 0o657, 0o711)), (24000, (0o561, 0o753, 0o711))])
def test_viterbi_k9_equals_the_reference_decoder(n, gens):
    """The AM decoder alone (radix-8 single-warp recursion + segmented traceback, csrc/am.cuh) against the oracle's
    conv_dec restatement with K = 9: clean, noisy and pure-noise frames of every length the chain uses (PIDS 80, P1 3750,
    P3 24000; 3751 / 3752 for the two shapes of a short last group) - bit for bit, ties and all.  Then the same frames
    with warm-ups of 3 steps, which leave most traceback walkers on a wrong path and most chunks of the recursion with
    wrong metrics: the repair rounds / the sequential re-run must bring back the reference decoder's output."""
    from nrsc5_b200 import engine as eng
    rng = np.random.default_rng(n)
    njobs = 6 if n > 4000 else 12
    _, clean = _k9_frames(rng, njobs // 3, n, gens, 0.0)
    _, noisy = _k9_frames(rng, njobs // 3, n, gens, 0.08)
    junk = rng.integers(-1, 2, (njobs - 2 * (njobs // 3), 3 * n)).astype(np.int8)
    sym = np.concatenate([clean, noisy, junk])
    want = np.stack([port.viterbi(sym[j], k=9, gens=gens) for j in range(sym.shape[0])])
    got, rounds, redone = eng.viterbi_k9(sym, gens)
    assert np.array_equal(got, want), np.argwhere(got != want)[:5]
    got3, rounds3, redone3 = eng.viterbi_k9(sym, gens, warmup=3, chunk_warmup=3)
    assert np.array_equal(got3, want)
    if n >= 3750:
        assert rounds3.max() >= 1 and rounds3.max() >= rounds.max()          # the traceback's repair loop really ran
        assert redone3.max() >= 1 and redone[: 2 * (njobs // 3)].max() == 0   # and so did the recursion's; decodable frames never need it
