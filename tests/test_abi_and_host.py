"""CPU-side checks of the product: the C-ABI library loads and exports every symbol that
include/wifi_b200.h declares, fails loudly without a GPU, and never routes through the oracle."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gnuradio-wifi-imagetransfer_b200")


def declared_symbols():
    h = open(os.path.join(ROOT, "include", "wifi_b200.h")).read()
    return sorted(set(re.findall(r"\b(wifi_b200_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol(W):
    lib = W.wifi_b200.lib()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(W.wifi_b200.EXPORTS) == syms
    assert lib.wifi_b200_abi_version() == 1
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(PKG, "libwifi_b200.so")]).decode()
    exported = set(re.findall(r" T (wifi_b200_\w+)", out))
    assert exported == set(syms)


def test_struct_layouts_match_header(W):
    w = W.wifi_b200
    assert w.FRAME_DTYPE.itemsize == 96 and ctypes.sizeof(w.Cfg) == 64
    assert ctypes.sizeof(w.ChanSeg) == w.CHANSEG_DTYPE.itemsize == 176
    from oracle import oracle as O
    assert O.FRAME_DTYPE == w.FRAME_DTYPE


def test_no_cpu_fallback(W):
    w = W.wifi_b200
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert w.device_count() == 0
    with pytest.raises(w.WifiB200Error) as e:
        W.Handle()
    assert e.value.code == w.E_NODEVICE
    with pytest.raises(w.WifiB200Error):
        W.wifi_phy_hier()


def test_mac_block_mirror(W):
    """ieee802_11.mac semantics (IRS_tranceiver.py:271) -- host-side integer work, no GPU needed."""
    import zlib
    m = W.mac([0x23] * 6, [0x42] * 6, [0xff] * 6)
    meta, p0 = m.app_in(b"abc")
    _, p1 = m.app_in(b"abc")
    assert meta == {"crc_included": True}
    assert p0[22:24] == b"\x00\x00" and p1[22:24] == (1 << 4).to_bytes(2, "little")
    assert p0[-4:] == zlib.crc32(p0[:-4]).to_bytes(4, "little")
    assert W.mac.phy_in(({}, p0[:-4])) == ({}, b"abc")
    # what "Extract Pics" does with a mac_out PDU (IRS_tranceiver_epy_block_2.py:34-35)
    import struct
    payload = struct.pack("=L", 3) + b"xyz"
    _, p = m.app_in(payload)
    assert p[:-4][24:][4:] == b"xyz"
    with pytest.raises(ValueError):
        m.app_in(bytes(1501))
    with pytest.raises(ValueError):
        W.mac([1, 2, 3], [0] * 6, [0] * 6)
    assert W.wifi_b200.n_sym(7, 1528) == 57 and W.wifi_b200.frame_samples(0, 1500) == 40481


def test_product_never_touches_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                # comments may cite the oracle; code may not include, import, link or dlopen it
                if re.search(r"^\s*(from|import)\s+\.*oracle|liboracle|#\s*include[^\n]*oracle|dlopen|CDLL\([^)]*oracle", txt, re.M):
                    bad.append(f)
    assert not bad, bad
    txt = open(os.path.join(ROOT, "include", "wifi_b200.h")).read()
    assert "oracle" not in txt.replace("the oracle's orc_frame", "")


def test_sources_do_not_name_banned_memcpy_batch_calls():
    pat = re.compile("cu" + "daMemcpy(3D)?Batch" + "Async|cuMemcpy(3D)?Batch" + "Async")
    for base in (PKG, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        for dirpath, _, files in os.walk(base):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def test_pcap_tap_and_parse_mac(W, tmp_path):
    """wireshark_connector / parse_mac stand-ins (reference taps, IRS_tranceiver.grc:446-495,1022-1038)."""
    import struct
    m = W.mac()
    pdus = [({"dlt": 105}, m.app_in(bytes([i]) * (10 + i))[1][:-4]) for i in range(3)]      # mac_out carries no FCS
    path = str(tmp_path / "wifi.pcap")
    with W.pcap.PcapWriter(path) as w:
        for i, p in enumerate(pdus):
            w.write(p, ts=1000.5 + i)
    raw = open(path, "rb").read()
    assert struct.unpack_from("<IHH", raw, 0) == (0xA1B2C3D4, 2, 4) and struct.unpack_from("<I", raw, 20)[0] == 105
    lt, recs = W.pcap.read_pcap(path)
    assert lt == 105 and [r[1] for r in recs] == [p[1] for p in pdus] and abs(recs[1][0] - 1001.5) < 1e-6
    h = W.pcap.parse_mac(pdus[2][1])
    assert h["frame_control"] == 0x0008 and h["type"] == 2 and h["subtype"] == 0 and h["seq_nr"] == 2 and h["frag_nr"] == 0
    assert h["addr1"] == "42:42:42:42:42:42" and h["addr2"] == "23:23:23:23:23:23" and h["addr3"] == "ff:ff:ff:ff:ff:ff"
    assert h["payload_len"] == 12 and W.pcap.parse_mac(b"short") is None
