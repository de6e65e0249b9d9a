#!/usr/bin/env python3
"""Pins the oracle to the REAL gr-ieee802-11: one command on any machine that has GNU Radio 3.10 and gr-ieee802-11
(maint-3.10) installed.  This image has neither, which is why parity is "unpinned" today.

    python tools/make_upstream_fixtures.py            # writes tests/golden/upstream_fixture.npz + .json
    python -m pytest tests/test_upstream_fixture.py   # oracle (CPU) and library (-m gpu) against what upstream produced

What it does
  1. builds the seeded captures of tests/golden/make_oracle_regression.py (all 8 MCS through the 3-tap channel, CFO,
     AWGN) plus one capture per equalizer -- with the ORACLE's TX and Philox channel, so the input is the committed,
     reproducible one; the TX side is pinned separately (step 3);
  2. runs every capture through the receive chain of the reference, wired exactly as gnu_radio/IRS_tranceiver.py:268-341
     wires it (delay 16, conjugate, multiply, moving_average_cc(48) / _ff(64), complex_to_mag(_squared), divide,
     ieee802_11.sync_short(0.56, 2), delay(320), sync_long(320), stream_to_vector(64), fft_vcc(64, forward, rectangular,
     shift), frame_equalizer(algo, 5.89e9, 10e6), decode_mac) and records what comes out: the PDUs of decode_mac 'out'
     (bytes + the snr / nomfreq / freqofs of their dict), the equalised points of frame_equalizer 'symbols', the 48-byte
     rows on frame_equalizer's stream output with their `wifi_start` tags, and sync_short's tag offsets;
  3. feeds the PSDUs through the reference's transmit chain (ieee802_11.mac is bypassed: PDUs go to the hier block's
     `mac_in`, gnu_radio/wifi_phy_hier.grc) and records the bursts on `samp_out`.
The fixture is small (a few hundred kB).  tests/test_upstream_fixture.py skips itself while the fixture is absent.

Where upstream is not a function of its input (moving_average re-seeding, GNU Radio's chunking) the comparison in the
test is the one SURVEY 8c defines: trigger set, frame_start, SIGNAL fields and PSDU bytes exactly; equalised points
within 2e-3; CFO estimates within 1e-5 rad/sample.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden", "upstream_fixture")


def need_gnuradio():
    try:
        from gnuradio import blocks, fft, gr  # noqa: F401
        from gnuradio.fft import window  # noqa: F401
        import ieee802_11  # noqa: F401
        import pmt  # noqa: F401
    except ImportError as e:
        sys.exit("this script needs GNU Radio 3.10 with gr-ieee802-11 (maint-3.10) importable: %s" % e)


def captures():
    """name -> (iq, psdus, encodings, equalizer): the regression capture under each equalizer + a marginal-SNR one."""
    from oracle import oracle as O
    from util import make_capture
    out = {}
    specs = [(e, 60 + 90 * e) for e in range(8)] + [(7, 1528), (3, 296)]
    taps = ((0, 1.0), (1, 0.4 * np.exp(1j * 1.0)), (3, 0.2 * np.exp(-2j)))
    y, psdus = make_capture(O, np.random.default_rng(7), specs, snr_db=24, cfo=0.009, taps=taps, seed=5, gap=900)
    for algo in range(4):
        out["regression_algo%d" % algo] = (y, psdus, [s[0] for s in specs], algo)
    y2, p2 = make_capture(O, np.random.default_rng(8), [(5, 400)] * 12 + [(7, 700)] * 12, snr_db=17, cfo=-0.004, seed=6, gap=1100)
    out["marginal_snr"] = (y2, p2, [5] * 12 + [7] * 12, 0)
    y3, p3 = make_capture(O, np.random.default_rng(9), [(1, 333), (3, 333), (5, 333), (7, 333)] * 4, snr_db=40, seed=7, gap=700)
    out["rate34_tails"] = (y3, p3, [1, 3, 5, 7] * 4, 0)       # clean channel: the truncated-traceback tail (DESIGN.md choice 2)
    return out


class Collector:
    """Message sink + stream probes built from embedded-python-style blocks."""

    def __init__(self, gr, pmt):
        self.pdus, self.symbols = [], []

        class Sink(gr.basic_block):
            def __init__(s, name, store, conv):
                gr.basic_block.__init__(s, name=name, in_sig=None, out_sig=None)
                s.message_port_register_in(pmt.intern("in"))
                s.set_msg_handler(pmt.intern("in"), lambda msg: store.append(conv(msg)))

        def pdu(msg):
            meta = pmt.to_python(pmt.car(msg)) or {}
            return ({str(k): (float(v) if isinstance(v, (int, float)) else str(v)) for k, v in meta.items()}, bytes(bytearray(pmt.u8vector_elements(pmt.cdr(msg)))))

        def sym(msg):
            return np.array(pmt.c32vector_elements(pmt.cdr(msg)), np.complex64)

        self.mac_sink = Sink("mac_out_sink", self.pdus, pdu)
        self.sym_sink = Sink("carrier_sink", self.symbols, sym)


def run_rx(y, algo):
    from gnuradio import blocks, fft, gr
    from gnuradio.fft import window
    import ieee802_11
    import pmt
    tb = gr.top_block()
    src = blocks.vector_source_c(y.tolist(), False, 1, [])
    delay16 = blocks.delay(gr.sizeof_gr_complex, 16)
    conj = blocks.conjugate_cc()
    mult = blocks.multiply_vcc(1)
    mavg_c = blocks.moving_average_cc(48, 1, 4000, 1)
    mavg_f = blocks.moving_average_ff(48 + 16, 1, 4000, 1)
    c2m, c2m2, div = blocks.complex_to_mag(1), blocks.complex_to_mag_squared(1), blocks.divide_ff(1)
    sshort = ieee802_11.sync_short(0.56, 2, False, False)
    delay320 = blocks.delay(gr.sizeof_gr_complex, 320)
    slong = ieee802_11.sync_long(320, False, False)
    s2v = blocks.stream_to_vector(gr.sizeof_gr_complex, 64)
    fftb = fft.fft_vcc(64, True, window.rectangular(64), True, 1)
    eq = ieee802_11.frame_equalizer(ieee802_11.Equalizer(algo), 5.89e9, 10e6, False, False)
    dec = ieee802_11.decode_mac(False, False)
    rows = blocks.vector_sink_b(48)
    short_out = blocks.tag_debug(gr.sizeof_gr_complex, "sync_short")
    short_out.set_display(False)
    short_out.set_save_all(True)
    col = Collector(gr, pmt)
    tb.connect(src, delay16, conj, (mult, 1))
    tb.connect(src, (mult, 0))
    tb.connect(mult, mavg_c, c2m, (div, 0))
    tb.connect(src, c2m2, mavg_f, (div, 1))
    tb.connect(delay16, (sshort, 0))
    tb.connect(mavg_c, (sshort, 1))
    tb.connect(div, (sshort, 2))
    tb.connect(sshort, (slong, 0))
    tb.connect(sshort, delay320, (slong, 1))
    tb.connect(sshort, short_out)
    tb.connect(slong, s2v, fftb, eq, dec)
    tb.connect(eq, rows)
    tb.msg_connect((dec, "out"), (col.mac_sink, "in"))
    tb.msg_connect((eq, "symbols"), (col.sym_sink, "in"))
    tb.run()
    time.sleep(0.2)           # message handlers drain
    tags = [(int(t.offset), float(pmt.to_double(t.value))) for t in short_out.current_tags() if pmt.symbol_to_string(t.key) == "wifi_start"]
    row_tags = []
    for t in rows.tags():
        if pmt.symbol_to_string(t.key) == "wifi_start" and pmt.is_dict(t.value):
            d = pmt.to_python(t.value)
            row_tags.append((int(t.offset), int(d.get("frame_bytes", -1)), int(d.get("encoding", -1)), float(d.get("snr", 0.0)), float(d.get("freq_offset", 0.0))))
    return {"pdus": col.pdus, "symbols": np.array(col.symbols, np.complex64).reshape(-1, 48) if col.symbols else np.zeros((0, 48), np.complex64),
            "rows": np.array(rows.data(), np.uint8).reshape(-1, 48), "short_tags": tags, "row_tags": row_tags}


def run_tx(psdus, encs):
    """mac_in -> samp_out of the hier block generated from gnu_radio/wifi_phy_hier.grc (grcc it first; GRC_HIER_PATH)."""
    from gnuradio import blocks, gr
    import pmt
    try:
        from wifi_phy_hier import wifi_phy_hier
    except ImportError:
        return None           # the hier block was not generated: RX-only fixture
    bursts = []
    for psdu, enc in zip(psdus, encs):
        tb = gr.top_block()
        phy = wifi_phy_hier(bandwidth=10e6, chan_est=0, encoding=enc, frequency=5.89e9, sensitivity=0.56)
        zero = blocks.vector_source_c([0.0] * 64, False, 1, [])
        sink = blocks.vector_sink_c(1)
        tb.connect(zero, (phy, 0))
        tb.connect((phy, 0), sink)
        tb.start()
        msg = pmt.cons(pmt.make_dict(), pmt.init_u8vector(len(psdu), list(bytearray(psdu))))
        phy.to_basic_block()._post(pmt.intern("mac_in"), msg)
        time.sleep(0.5)
        tb.stop()
        tb.wait()
        bursts.append(np.array(sink.data(), np.complex64))
    return bursts


def main():
    need_gnuradio()
    arrays, meta = {}, {"made_with": {}, "captures": {}}
    import gnuradio
    meta["made_with"] = {"gnuradio": getattr(gnuradio.gr, "version", lambda: "?")() if hasattr(gnuradio, "gr") else "?", "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    for name, (y, psdus, encs, algo) in captures().items():
        r = run_rx(y, algo)
        arrays[name + "/iq"] = y
        arrays[name + "/symbols"] = r["symbols"]
        arrays[name + "/rows"] = r["rows"]
        arrays[name + "/pdu_blob"] = np.frombuffer(b"".join(p for _, p in r["pdus"]), np.uint8)
        meta["captures"][name] = {"equalizer": algo, "sent_psdus_hex": [p.hex() for p in psdus], "sent_encodings": encs,
                                  "pdu_lengths": [len(p) for _, p in r["pdus"]], "pdu_meta": [m for m, _ in r["pdus"]],
                                  "short_tags": r["short_tags"], "row_tags": r["row_tags"]}
        print(name, "PDUs", len(r["pdus"]), "rows", len(r["rows"]), "tags", len(r["short_tags"]))
    first = next(iter(captures().values()))
    tx = run_tx(first[1], first[2])
    if tx is not None:
        for i, b in enumerate(tx):
            arrays["tx/%d" % i] = b
        meta["tx"] = {"psdus_hex": [p.hex() for p in first[1]], "encodings": first[2], "note": "scrambler seed runs 1, 2, ... per frame of one mapper instance; here every frame has its own instance: seed 1"}
    np.savez_compressed(OUT + ".npz", **arrays)
    json.dump(meta, open(OUT + ".json", "w"), indent=1)
    print("wrote", OUT + ".npz", OUT + ".json")


if __name__ == "__main__":
    main()
