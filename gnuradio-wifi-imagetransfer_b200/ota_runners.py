"""GNU-Radio-free stand-ins for the reference's two over-the-air flowgraphs, with an IQ file (or any file-like
object) where the SDR stands, so that a transmission can be recorded and replayed without hardware:

  IrsUser  (gnu_radio/IRS_user.py / IRS_user.grc)
      UDP :52001 (socket_pdu UDP_SERVER, mtu 10000, IRS_user.grc:299-312) -> ieee802_11.mac (IRS_user.py:192)
      -> wifi_phy_hier TX (:154-160) -> x multi_const 0.5 (:79,196) -> packet_pad2(100, 1000) (:193)
      -> [soapy HackRF sink :167-173]  ==> interleaved fc32 (or int16 "sc16") IQ written to a file
  IrsAp    (gnu_radio/IRS_AP.py)
      [uhd.usrp_source, cpu_format fc32 :163-177] ==> IQ read from a file (fc32 or sc16)
      -> flattened RX chain (:268-285,294-316) -> "Extract Pics" data[24:][4:] (IRS_AP_epy_block_2.py) -> UDP localhost:10010

All PHY work happens in libwifi_b200.so; these classes only move bytes.  sc16 files go through
wifi_b200_rx_batch_sc16 (conversion on the GPU).
"""
import argparse
import socket

import numpy as np

from .wifi_phy_hier import mac, wifi_phy_hier


class IrsUser:
    """Transmit side.  Defaults are the flowgraph's variables (IRS_user.py:76-82: samp_rate 1e6, encoding 2, multi_const 0.5)."""

    def __init__(self, out, in_port=52001, encoding=2, multi_const=0.5, samp_rate=1e6, freq=5.89e9, device=0, mtu=10000, fmt="fc32",
                 sc16_scale=1.0 / 16384.0):
        if fmt not in ("fc32", "sc16"):
            raise ValueError("fmt must be fc32 or sc16")
        self.out = open(out, "wb") if isinstance(out, str) else out
        self.multi_const, self.mtu, self.fmt, self.sc16_scale = float(multi_const), mtu, fmt, float(sc16_scale)
        self.phy = wifi_phy_hier(bandwidth=samp_rate, encoding=encoding, frequency=freq, device=device, max_samples=1 << 18)
        self.mac = mac([0x23] * 6, [0x42] * 6, [0xff] * 6)
        self.sock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        self.sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.sock.bind(("", in_port))
        self.in_port = self.sock.getsockname()[1]
        self.sock.settimeout(0.2)
        self.stats = {"datagrams_in": 0, "bursts_out": 0, "samples_out": 0, "dropped_oversize": 0}

    def set_encoding(self, encoding):
        self.phy.set_encoding(encoding)

    def set_multi_const(self, v):
        self.multi_const = float(v)

    def handle_datagram(self, data):
        data = data[:self.mtu]
        self.stats["datagrams_in"] += 1
        try:
            pdu = self.mac.app_in(data)
        except ValueError:                       # > 1500 bytes: upstream mac throws, catch_exceptions swallows it
            self.stats["dropped_oversize"] += 1
            return
        burst = self.phy.mac_in(pdu)
        self.phy.samp_out.clear()
        x = np.concatenate([np.zeros(100, np.complex64), np.complex64(self.multi_const) * burst, np.zeros(1000, np.complex64)]).astype(np.complex64)
        if self.fmt == "sc16":
            iq = np.clip(np.rint(x.view(np.float32) / np.float32(self.sc16_scale)), -32768, 32767).astype(np.int16)
            self.out.write(iq.tobytes())
        else:
            self.out.write(x.tobytes())
        self.stats["bursts_out"] += 1
        self.stats["samples_out"] += x.size

    def serve(self, max_datagrams=None, idle_timeouts=None):
        n = idle = 0
        while (max_datagrams is None or n < max_datagrams) and (idle_timeouts is None or idle < idle_timeouts):
            try:
                data, _ = self.sock.recvfrom(65536)
            except socket.timeout:
                idle += 1
                continue
            idle = 0
            self.handle_datagram(data)
            n += 1
        self.out.flush()

    def close(self):
        self.sock.close()
        self.out.flush()
        self.phy.handle.close()


class IrsAp:
    """Receive side.  Defaults are the flowgraph's variables (IRS_AP.py:77-88: samp_rate 1e6, chan_est 0)."""

    def __init__(self, out_addr=("localhost", 10010), chan_est=0, samp_rate=1e6, freq=5.89e9, device=0, fmt="fc32", sc16_scale=1.0 / 16384.0,
                 chunk=1 << 20):
        if fmt not in ("fc32", "sc16"):
            raise ValueError("fmt must be fc32 or sc16")
        self.fmt, self.sc16_scale, self.chunk = fmt, float(sc16_scale), int(chunk)
        self.phy = wifi_phy_hier(bandwidth=samp_rate, chan_est=chan_est, frequency=freq, device=device, max_samples=4 * self.chunk)
        self.sock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        self.out_addr = out_addr
        self.stats = {"samples_in": 0, "pdus_out": 0}

    def set_chan_est(self, chan_est):
        self.phy.set_chan_est(chan_est)

    def _emit(self, pdus):
        for _meta, mpdu in pdus:
            self.sock.sendto(bytes(mpdu[24:][4:]), self.out_addr)     # "Extract Pics"
            self.stats["pdus_out"] += 1

    def run(self, src):
        """Decode a whole recording: `src` is a path or a binary file object of interleaved fc32 / int16 IQ."""
        f = open(src, "rb") if isinstance(src, str) else src
        item = 8 if self.fmt == "fc32" else 4
        while True:
            raw = f.read(self.chunk * item)
            raw = raw[:len(raw) - len(raw) % item]
            if not raw:
                break
            if self.fmt == "fc32":
                x = np.frombuffer(raw, np.complex64)
            else:                                                       # the streaming entry point takes fc32: same rule as the GPU converter
                x = (np.frombuffer(raw, np.int16).astype(np.float32) * np.float32(self.sc16_scale)).view(np.complex64)
            self.stats["samples_in"] += x.size
            self._emit(self.phy.samp_in(x))
        self._emit(self.phy.samp_in(np.zeros(0, np.complex64), flush=True))
        return self.stats

    def close(self):
        self.sock.close()
        self.phy.handle.close()


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    u = sub.add_parser("user", help="UDP datagrams -> IQ file (IRS_user)")
    u.add_argument("out")
    u.add_argument("--in-port", type=int, default=52001)
    u.add_argument("--encoding", type=int, default=2)
    u.add_argument("--multi-const", type=float, default=0.5)
    u.add_argument("--format", default="fc32", choices=["fc32", "sc16"])
    a = sub.add_parser("ap", help="IQ file -> decoded patches on UDP (IRS_AP)")
    a.add_argument("src")
    a.add_argument("--out-host", default="localhost")
    a.add_argument("--out-port", type=int, default=10010)
    a.add_argument("--chan-est", type=int, default=0)
    a.add_argument("--format", default="fc32", choices=["fc32", "sc16"])
    args = ap.parse_args()
    if args.cmd == "user":
        t = IrsUser(args.out, args.in_port, args.encoding, args.multi_const, fmt=args.format)
        print("listening on UDP :%d, writing %s IQ to %s" % (t.in_port, args.format, args.out))
        try:
            t.serve()
        finally:
            t.close()
    else:
        r = IrsAp((args.out_host, args.out_port), args.chan_est, fmt=args.format)
        try:
            print(r.run(args.src))
        finally:
            r.close()


if __name__ == "__main__":
    main()
