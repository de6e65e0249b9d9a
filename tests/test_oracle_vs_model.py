"""The oracle against an independent float64 numpy model built from the standard's formulas."""
import numpy as np
import pytest

import ref_model as R
from util import make_capture


@pytest.mark.parametrize("enc", range(8))
def test_tx_waveform_and_mapper_indices(O, enc):
    rng = np.random.default_rng(enc)
    psdu = rng.integers(0, 256, 77 + enc, dtype=np.uint8).tobytes()
    want, idx = R.tx_frame(psdu, enc, 1 + 9 * enc)
    got = O.tx_frame(psdu, enc, 1 + 9 * enc)
    assert got.size == want.size
    assert np.abs(got - want).max() < 1e-5                  # TX IQ |delta| <= 1e-5
    assert np.array_equal(O.tx_symbols(psdu, enc, 1 + 9 * enc), idx)
    assert abs(np.mean(np.abs(got[320:]) ** 2) - 1.0) < 0.15   # unit mean power


@pytest.mark.parametrize("enc", [0, 3, 5, 7])
def test_cross_decode(O, enc):
    """model RX decodes the oracle's TX; oracle RX decodes the model's TX."""
    rng = np.random.default_rng(10 + enc)
    psdu = O.mac_frame(rng.integers(0, 256, 50, dtype=np.uint8).tobytes(), seq=enc)
    iq = O.tx_frame(psdu, enc, 17)
    e, ln, data = R.rx_frame(iq, 0)
    assert (e, ln, data) == (enc, len(psdu), psdu)
    m, _ = R.tx_frame(psdu, enc, 17)
    x = np.concatenate([np.zeros(200), 0.5 * m, np.zeros(600)]).astype(np.complex64)
    x = O.channel(x, noise_sigma=0.5 * 10 ** (-35 / 20), seed=enc, cfo=0.003)
    r = O.rx(x, algo=0)
    assert [r.psdu(i) for i in range(len(r.frames)) if r.frames[i]["crc_ok"]] == [psdu]


def test_rx_estimates_against_truth(O):
    rng = np.random.default_rng(42)
    y, psdus = make_capture(O, rng, [(4, 300)], snr_db=30, cfo=0.0126, lead=100)
    r = O.rx(y, algo=0)
    f = r.frames[0]
    assert 130 <= f["trigger"] <= 145 and 165 <= f["frame_start"] + f["trigger"] - 138 + 0 <= 175 + 10
    total = float(f["freq_short"]) - float(f["freq_long"])
    assert abs(total - 0.0126) < 3e-4                       # rad/sample
    assert abs(f["snr"] - 30) < 2.5
    assert f["burst_len"] == y.size - f["trigger"] and f["crc_ok"] == 1
    # equalised points sit on the 16-QAM grid
    pts = r.carrier[f["row_off"]:f["row_off"] + f["n_rows"]].reshape(-1)
    grid = O.constellation(4)
    assert np.mean(np.min(np.abs(pts[:, None] - grid[None, :]), axis=1)) < 0.06


def test_loopback_all_equalizers_all_mcs(O):
    rng = np.random.default_rng(7)
    for algo in range(4):
        specs = [(e, 64 + 16 * e) for e in range(8)]
        y, psdus = make_capture(O, rng, specs, snr_db=36, cfo=-0.008, seed=algo)
        r = O.rx(y, algo=algo)
        assert r.pdus() == [p[:-4] for p in psdus], algo


def test_oversize_and_limits(O):
    with pytest.raises(ValueError):
        O.tx_frame(bytes(1529), 7, 1)
    with pytest.raises(ValueError):
        O.mac_frame(bytes(1501), 0)
    assert O.tx_frame(bytes(1528), 0, 1).size == 80 * (5 + 511) + 1   # MAX_SYM
