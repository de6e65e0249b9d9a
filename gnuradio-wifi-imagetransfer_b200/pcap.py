"""Observability taps of the reference flowgraphs (disabled there, gnu_radio/IRS_tranceiver.grc:446-495,1022-1038):
`foo.wireshark_connector(LINKTYPE 127 / 105) -> file_sink /tmp/wifi.pcap` and `ieee802_11.parse_mac`.

`PcapWriter` stores `mac_out` PDUs (meta['dlt'] = 105 = LINKTYPE_IEEE802_11, frames without FCS) in a classic
little-endian pcap file that Wireshark opens; `parse_mac` returns what upstream's parse_mac block logs for a data
frame (frame control, duration, the three addresses, sequence / fragment number).  Host-side byte shuffling only.
"""
import struct
import time

LINKTYPE_IEEE802_11 = 105
_MAGIC = 0xA1B2C3D4


class PcapWriter:
    def __init__(self, path, linktype=LINKTYPE_IEEE802_11, snaplen=65535):
        self._f = open(path, "wb")
        self._f.write(struct.pack("<IHHiIII", _MAGIC, 2, 4, 0, 0, snaplen, linktype))
        self.count = 0

    def write(self, pdu, ts=None):
        """pdu = (meta, bytes) as published on mac_out"""
        _meta, data = pdu
        data = bytes(data)
        ts = time.time() if ts is None else float(ts)
        sec = int(ts)
        self._f.write(struct.pack("<IIII", sec, int(round((ts - sec) * 1e6)) % 1000000, len(data), len(data)))
        self._f.write(data)
        self.count += 1

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def read_pcap(path):
    """-> (linktype, [(timestamp, bytes)])"""
    with open(path, "rb") as f:
        magic, _vmaj, _vmin, _tz, _sig, _snap, linktype = struct.unpack("<IHHiIII", f.read(24))
        if magic != _MAGIC:
            raise ValueError("not a little-endian microsecond pcap file")
        out = []
        while True:
            h = f.read(16)
            if len(h) < 16:
                break
            sec, usec, incl, _orig = struct.unpack("<IIII", h)
            out.append((sec + usec * 1e-6, f.read(incl)))
    return linktype, out


def parse_mac(mpdu):
    """Fields of an 802.11 data/management header (24 bytes) as upstream parse_mac prints them."""
    if len(mpdu) < 24:
        return None
    fc, dur = struct.unpack_from("<HH", mpdu, 0)
    seq = struct.unpack_from("<H", mpdu, 22)[0]
    fmt = lambda b: ":".join("%02x" % x for x in b)     # noqa: E731
    return {"frame_control": fc, "type": (fc >> 2) & 3, "subtype": (fc >> 4) & 15, "duration": dur,
            "addr1": fmt(mpdu[4:10]), "addr2": fmt(mpdu[10:16]), "addr3": fmt(mpdu[16:22]),
            "seq_nr": seq >> 4, "frag_nr": seq & 15, "payload_len": len(mpdu) - 24}
