"""Regenerates tests/golden/oracle_regression.json: digests of what the CPU oracle produces for fixed seeds (TX IQ of
every MCS, a channelised multi-frame capture, its RX frame table, decisions, equalised points and PSDUs, hard and
soft).  The oracle defines parity for the CUDA library, so a change of these digests is a change of the numerical
contract and must be deliberate:  python tests/golden/make_oracle_regression.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build():
    from oracle import oracle as O
    from util import make_capture, make_psdu
    out = {"tx": {}, "rx": {}}
    rng = np.random.default_rng(2024)
    for enc in range(8):
        psdu = make_psdu(O, rng, 100 + 150 * enc, seq=enc)
        out["tx"][str(enc)] = {"psdu_len": len(psdu), "iq_sha256": sha(O.tx_frame(psdu, enc, seed=17 + enc))}
    specs = [(e, 60 + 90 * e) for e in range(8)] + [(7, 1528), (3, 296)]
    taps = ((0, 1.0), (1, 0.4 * np.exp(1j * 1.0)), (3, 0.2 * np.exp(-2j)))
    y, _ = make_capture(O, np.random.default_rng(7), specs, snr_db=24, cfo=0.009, taps=taps, seed=5, gap=900)
    out["capture_sha256"] = sha(y)
    for algo in range(4):
        for soft in (False, True):
            r = O.rx(y, algo=algo, soft=soft)
            f = r.frames
            out["rx"]["algo%d_%s" % (algo, "soft" if soft else "hard")] = {
                "triggers": [int(v) for v in f["trigger"]], "frame_start": [int(v) for v in f["frame_start"]],
                "encoding": [int(v) for v in f["encoding"]], "length": [int(v) for v in f["length"]], "crc_ok": [int(v) for v in f["crc_ok"]],
                "freq_sha256": sha(np.stack([f["freq_short"], f["freq_long"]])), "rows_sha256": sha(r.rows), "carrier_sha256": sha(r.carrier),
                "pdus_sha256": hashlib.sha256(b"".join(r.pdus())).hexdigest()}
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_regression.json")
    json.dump(build(), open(path, "w"), indent=1)
    print("wrote", path)
