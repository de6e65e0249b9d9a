"""GNU Radio adapter.  GNU Radio is not installed in the build image, so the adapter is
driven through the small stand-ins under tests/fake_gr (the test plays the scheduler); where GNU Radio is
installed the same module binds to the real `gnuradio.gr` / `pmt`."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAKE = os.path.join(ROOT, "tests", "fake_gr")


def _run(code, with_stub):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([ROOT] + ([FAKE] if with_stub else []))
    return subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)


def test_adapter_needs_gnuradio_and_says_so():
    try:
        import gnuradio  # noqa: F401
        pytest.skip("GNU Radio is installed here")
    except ImportError:
        pass
    r = _run("import wifi_b200.gr_adapter", with_stub=False)
    assert r.returncode != 0 and "needs GNU Radio" in r.stderr
    # the rest of the package never imports it
    r = _run("import wifi_b200, sys; assert 'gnuradio' not in sys.modules and 'pmt' not in sys.modules", with_stub=False)
    assert r.returncode == 0, r.stderr


def test_adapter_imports_against_the_gr_api_surface_it_uses():
    r = _run("import wifi_b200.gr_adapter as g; print(sorted(n for n in dir(g) if n.startswith('wifi_')))", with_stub=True)
    assert r.returncode == 0, r.stderr
    assert "wifi_phy_hier_b200" in r.stdout and "wifi_rx_b200" in r.stdout and "wifi_tx_b200" in r.stdout


def test_block_yml_matches_the_reference_hier_block():
    import json
    import yaml
    y = yaml.safe_load(open(os.path.join(ROOT, "gnuradio-wifi-imagetransfer_b200", "grc", "ieee802_11_wifi_phy_hier_b200.block.yml")))
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "hier_constants.json")))
    ids = [p["id"] for p in y["parameters"]]
    assert ids[:5] == ["bandwidth", "chan_est", "encoding", "frequency", "sensitivity"]       # wifi_phy_hier.grc parameters
    assert [p.get("id", p.get("label")) for p in y["inputs"]] == ["samp_in", "mac_in"]
    assert [p.get("id", p.get("label")) for p in y["outputs"]] == ["samp_out", "mac_out", "carrier"]
    assert y["category"] == "[IEEE802.11]"
    assert sorted(ref["hier_defaults"]) == sorted(ids[:5])                                     # parameters of the reference's hier block
    dflt = {p["id"]: str(p["default"]) for p in y["parameters"]}
    assert float(dflt["bandwidth"]) == float(ref["hier_defaults"]["bandwidth"]) and float(dflt["sensitivity"]) == float(ref["hier_defaults"]["sensitivity"])
    assert float(dflt["frequency"]) == float(ref["hier_defaults"]["frequency"])


@pytest.mark.gpu
def test_hier_block_ports_against_the_oracle(O):
    """mac_in -> samp_out (burst + packet_len tag) and samp_in -> mac_out / carrier, scheduler played by hand."""
    sys.path.insert(0, FAKE)
    try:
        import importlib
        import wifi_b200
        g = importlib.import_module("gnuradio-wifi-imagetransfer_b200.gr_adapter")
        import pmt
        blk = g.wifi_phy_hier_b200(bandwidth=10e6, chan_est=1, encoding=5, frequency=5.89e9, sensitivity=0.56, want_carrier=True,
                                   max_samples=1 << 18)
        assert blk.hier_in == ["mac_in"] and blk.hier_out == ["mac_out", "carrier"]
        assert len(blk.connections) == 2 and len(blk.msg_connections) == 3
        m = wifi_b200.mac()
        rng = np.random.default_rng(5)
        psdus, parts = [], [np.zeros(300, np.complex64)]
        for i in range(3):
            pdu = m.app_in(rng.integers(0, 256, 200 + 50 * i, dtype=np.uint8).tobytes())
            psdus.append(pdu[1])
            blk.tx.handlers["mac_in"](g.pdu_from_python({}, pdu[1]))
        # play the scheduler on the TX block: small output buffers, bursts straddle calls
        got = []
        while True:
            out = np.zeros(1000, np.complex64)
            n = blk.tx.general_work([], [out])
            if n == 0:
                break
            got.append(out[:n].copy())
            blk.tx.written[0] = blk.tx.written.get(0, 0) + n
        stream = np.concatenate(got)
        refs = [O.tx_frame(p, 5, seed=i + 1) for i, p in enumerate(psdus)]
        assert np.array_equal(stream, np.concatenate(refs))
        starts = np.cumsum([0] + [len(r) for r in refs[:-1]])
        assert [(t[1], t[2], t[3]) for t in blk.tx.tags] == [(int(s), "packet_len", len(r)) for s, r in zip(starts, refs)]
        # RX: the bursts through the oracle's channel, fed in uneven chunks
        for r in refs:
            parts += [0.6 * r, np.zeros(1200, np.complex64)]
        x = np.concatenate(parts).astype(np.complex64)
        y = O.channel(x, gain=1.0, cfo=0.003, noise_sigma=0.6 * 10 ** (-28 / 20), seed=3)
        pos = 0
        for n in (777, 5000, 123, 10 ** 9):
            chunk = y[pos:pos + n]
            pos += len(chunk)
            assert blk.rx.general_work([chunk], []) == 0
        assert blk.rx.consumed[0] == len(y)
        blk.rx.stop()
        ref = O.rx(y, algo=1, want_carrier=True)
        pdus = [g.pdu_to_python(p) for p in blk.rx.published["mac_out"]]
        assert [p[1] for p in pdus] == ref.pdus() == [p[:-4] for p in psdus]
        assert all(set(p[0]) == {"snr", "nomfreq", "freqofs", "dlt", "encoding"} and p[0]["dlt"] == 105 for p in pdus)
        car = np.array([pmt.cdr(c) for c in blk.rx.published["carrier"]])
        ok = np.nonzero(ref.frames["crc_ok"])[0]
        want = np.concatenate([ref.carrier[int(ref.frames["row_off"][i]):int(ref.frames["row_off"][i]) + int(ref.frames["n_rows"][i])] for i in ok])
        assert car.shape == want.shape and np.array_equal(car, want)
        blk.set_encoding(3)
        assert blk.get_encoding() == 3
    finally:
        sys.path.remove(FAKE)
        for k in [k for k in sys.modules if k == "pmt" or k.startswith("gnuradio") and "imagetransfer" not in k]:
            del sys.modules[k]
