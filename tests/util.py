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
    bad = ~np.isclose(gpu.frames["snr"][ok], ref.frames["snr"][ok], rtol=1e-9, atol=1e-9, equal_nan=True)   # double log10: libm vs device
    assert not bad.any(), ("snr", np.nonzero(ok)[0][bad][:5], gpu.frames["snr"][ok][bad][:5], ref.frames["snr"][ok][bad][:5])
    for i in range(len(ref.frames)):
        if ref.frames[i]["decoded"]:
            assert gpu.psdu(i) == ref.psdu(i), ("psdu", i)


def oversize_burst(rng, n_data_syms=30, length=4000, enc=0):
    """Preamble + a SIGNAL field announcing `length` bytes (more than decode_mac accepts: its tag is refused and the rows
    behind it feed whatever collection is open) + BPSK data symbols; built with the independent numpy model."""
    import ref_model as M
    S = np.zeros(64, complex)
    for k, s in M.STS_POS.items():
        S[k % 64] = s * np.sqrt(13 / 6) * (1 + 1j)
    L = np.zeros(64, complex)
    for i, k in enumerate(range(-26, 27)):
        L[k % 64] = M.LTS[i]
    sts_t, lts_t = M.time_symbol(S), M.time_symbol(L)
    parts = [np.tile(sts_t, 3)[:160], np.concatenate([lts_t[32:], lts_t, lts_t])]
    sig = np.zeros(48, np.uint8)
    sig[M.interleave_perm(0)] = M.conv_encode(M.signal_bits(enc, length))
    t = M.time_symbol(M.ofdm_symbol(M.map_bits(sig, 0), M.PILOT_POLARITY[0]))
    parts.append(np.concatenate([t[48:], t]))
    for n in range(n_data_syms):
        t = M.time_symbol(M.ofdm_symbol(M.map_bits(rng.integers(0, 2, 48), 0), M.PILOT_POLARITY[(n + 1) % 127]))
        parts.append(np.concatenate([t[48:], t]))
    return np.concatenate(parts).astype(np.complex64)


def capture_with_collection_over_many_bursts(O, rng, n_interferers=6, spacing=1100):
    """One BPSK 1/2 frame cut short again and again by stronger bursts whose SIGNAL fields announce oversize PSDUs:
    decode_mac refuses their tags and keeps collecting the first frame's symbols through all of them."""
    p = make_psdu(O, rng, 230, seq=1)                      # 79 symbols
    a = O.tx_frame(p, 0, seed=5)
    x = np.zeros(100 + a.size + spacing * (n_interferers + 2) + 3000, np.complex64)
    x[100:100 + a.size] += 0.25 * a
    pos = 100 + 1300
    for _ in range(n_interferers):
        b = oversize_burst(rng, n_data_syms=(spacing - 500) // 80)
        x[pos:pos + b.size] += b
        pos += spacing
    return O.channel(x, gain=0.6, cfo=0.002, noise_sigma=0.003, seed=12)


def adversarial_stream(O, trunc=500, n_zone=230):
    """Three ordinary frames, then `n_zone` back-to-back preambles cut to `trunc` samples with not one idle sample
    between them, then four ordinary frames.  trunc = 500: the plateaus follow each other more closely than MIN_GAP
    plus a plateau, so two sync_short trigger chains that start at different points of the zone never merge.
    trunc = 641 (whole two-symbol frames): every SIGNAL decodes but no burst is long enough for a data symbol, so
    decode_mac carries a pending tag across the whole zone."""
    rng = np.random.default_rng(3)
    head, _ = make_capture(O, rng, [(2, 300)] * 3, snr_db=30, seed=1, gap=900)
    tail, _ = make_capture(O, rng, [(2, 300)] * 4, snr_db=30, seed=2, gap=900, lead=600)
    f = O.tx_frame(make_psdu(O, rng, 2), 0, seed=3)
    zone = O.channel(np.concatenate([f[:trunc]] * n_zone), gain=0.6, noise_sigma=0.6 * 10 ** (-30 / 20), seed=7)
    return np.concatenate([head, zone, tail]).astype(np.complex64)


def oracle_segment_decoder(O, y, **kw):
    """decode(lo, end, state, final) callback of sharding.reconcile on top of the oracle."""
    def decode(lo, end, st, final):
        if st is None:
            return O.rx(y[lo:end], want_carrier=False, final=final, **kw).frames
        h = st["hist"]
        return O.rx(y[lo - h:end], want_carrier=False, final=final, hist=h, min_pos=st["min_pos"], fo_carry=st["fo_carry"], **kw).frames
    return decode


def gpu_segment_decoder(h, y):
    """decode(lo, end, state, final) callback of sharding.reconcile on top of the library: the capture lives in device
    memory once, a segment is a pair of offsets into it (wifi_b200_rx_batch_dev / _rx_batch_dev_state)."""
    import threading
    import torch
    from importlib import import_module
    w = import_module("gnuradio-wifi-imagetransfer_b200.wifi_b200")
    dev = torch.from_numpy(np.ascontiguousarray(y).view(np.float32)).cuda()
    lock = threading.Lock()        # one handle: a call and the fetch of its results belong together

    def decode(lo, end, st, final):
        off = np.array([lo, end], np.uint64)
        with lock:
            if st is None:
                return h.rx_batch_dev(dev.data_ptr(), off, final=final, fetch=True).frames
            state = np.zeros(1, w.LINK_STATE_DTYPE)
            state["min_pos"], state["fo_carry"], state["hist"] = st["min_pos"], st["fo_carry"], st["hist"]
            return h.rx_batch_dev_state(dev.data_ptr(), off, state, final=final, fetch=True).frames
    decode.keepalive = dev
    return decode
