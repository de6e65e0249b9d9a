"""Shared helpers for the parity tests: build seeded captures with the oracle TX + channel."""
import numpy as np

CMP_EXACT = ("trigger", "link", "burst_len", "found", "frame_start", "n_syms", "sig_ok", "encoding", "length",
             "frame_symbols", "n_rows", "accepted", "decoded", "crc_ok")


def make_psdu(O, rng, length, seq=0):
    if length >= 28:
        return O.mac_frame(rng.integers(0, 256, length - 28, dtype=np.uint8).tobytes(), seq=seq)
    body = rng.integers(0, 256, max(length - 4, 0), dtype=np.uint8).tobytes()
    return body + int(O.crc32(body)).to_bytes(4, "little")


def make_capture(O, rng, specs, snr_db=30.0, cfo=0.0, taps=((0, 1.0),), lead=100, gap=1000, seed=0, gain=0.6):
    """specs: list of (enc, psdu_len).  Returns (iq complex64, list of psdus)."""
    parts, psdus = [np.zeros(lead, np.complex64)], []
    for i, (enc, ln) in enumerate(specs):
        p = make_psdu(O, rng, ln, seq=i)
        psdus.append(p)
        parts.append(O.tx_frame(p, enc, seed=1 + i % 127))
        parts.append(np.zeros(gap, np.complex64))
    x = np.concatenate(parts).astype(np.complex64)
    sigma = gain * 10 ** (-snr_db / 20) if snr_db is not None else 0.0
    y = O.channel(x, gain=gain, cfo=cfo, noise_sigma=sigma, taps=taps, seed=seed)
    return y, psdus


def assert_frames_equal(gpu, ref, float_exact=True):
    assert len(gpu.frames) == len(ref.frames), (len(gpu.frames), len(ref.frames))
    for k in CMP_EXACT:
        assert np.array_equal(gpu.frames[k], ref.frames[k]), (k, gpu.frames[k], ref.frames[k])
    if float_exact:
        for k in ("freq_short", "freq_long"):
            assert np.array_equal(gpu.frames[k], ref.frames[k], equal_nan=True), (k, gpu.frames[k], ref.frames[k])
    ok = ref.frames["sig_ok"] == 1
    assert np.allclose(gpu.frames["snr"][ok], ref.frames["snr"][ok], rtol=1e-9, atol=1e-9, equal_nan=True)   # double log10: libm vs device
    for i in range(len(ref.frames)):
        if ref.frames[i]["decoded"]:
            assert gpu.psdu(i) == ref.psdu(i), ("psdu", i)
