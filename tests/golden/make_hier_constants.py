#!/usr/bin/env python
"""Extracts the PHY constants written in the reference's hier block into a small fixture.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_hier_constants.py
Output: tests/golden/hier_constants.json (committed).  Source: gnu_radio/wifi_phy_hier.grc
 - ofdm_carrier_allocator: occupied_carriers, pilot_carriers, pilot_symbols, sync_words (:336-405)
 - fft_vxx_0_0 window (:459-479), cyclic prefixer cp_len / rolloff (:406-424)
 - sync_short threshold / min_plateau (:716-734), sync_long sync_length (:698-715), window sizes
"""
import json
import os
import re

import yaml

SRC = "/root/reference/gnu_radio/wifi_phy_hier.grc"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hier_constants.json")


def main():
    doc = yaml.safe_load(open(SRC))
    blocks = {b["name"]: b for b in doc["blocks"]}
    alloc = blocks["digital_ofdm_carrier_allocator_cvc_0_0_0"]["parameters"]
    env = {"range": range, "list": list}
    occupied = [list(c) for c in eval(alloc["occupied_carriers"], env)]
    pilots_c = [list(c) for c in eval(alloc["pilot_carriers"], env)]
    pilots_s = [list(c) for c in eval(alloc["pilot_symbols"], env)]
    sync = [[[complex(v).real, complex(v).imag] for v in w] for w in eval(alloc["sync_words"], env)]
    var = {b["name"]: b["parameters"].get("value") for b in doc["blocks"] if b["id"] in ("variable", "parameter")}
    par = {b["name"]: b["parameters"] for b in doc["blocks"]}
    out = {
        "source": "gnu_radio/wifi_phy_hier.grc",
        "occupied_carriers": occupied, "pilot_carriers": pilots_c, "pilot_symbols": pilots_s, "sync_words": sync,
        "fft_len": int(alloc["fft_len"]),
        "ifft_window": par["fft_vxx_0_0"]["window"],
        "cp_len": par["digital_ofdm_cyclic_prefixer_0_0"]["cp_len"],
        "rolloff": par["digital_ofdm_cyclic_prefixer_0_0"]["rolloff"],
        "sync_short": {k: par["sync_short"][k] for k in ("threshold", "min_plateau")},
        "sync_long": {"sync_length": par["sync_long"]["sync_length"]},
        "variables": {k: v for k, v in var.items() if v is not None},
        "moving_average_cc_length": par["blocks_moving_average_xx_0"]["length"],
        "moving_average_ff_length": par["blocks_moving_average_xx_1"]["length"],
        "delay_short": par["blocks_delay_0_0"]["delay"], "delay_long": par["blocks_delay_0"]["delay"],
        "hier_defaults": {k: par[k]["value"] for k in ("bandwidth", "chan_est", "encoding", "frequency", "sensitivity")},
    }
    json.dump(out, open(OUT, "w"), indent=0)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
