"""Host-side mirror of the reference's PHY interface.

`wifi_phy_hier` mirrors the GRC-generated hierarchical block of
gnu_radio/wifi_phy_hier.grc (constructor keywords :83-99,299-315,442-458,501-517,681-697,
call site gnu_radio/IRS_tranceiver.py:178-184; setters used at :386,:427,:442) with the
same ports: message in `mac_in`, stream out `samp_out`, stream in `samp_in`, message out
`mac_out` and `carrier` (:587-680).  `mac` mirrors ieee802_11.mac (IRS_tranceiver.py:271).
PDUs are (meta dict, bytes) pairs -- the Python image of pmt.cons(dict, u8vector).
Everything that computes runs in libwifi_b200.so on the GPU; there is no CPU path.
"""
import math

import numpy as np

from . import wifi_b200 as _w

BPSK_1_2, BPSK_3_4, QPSK_1_2, QPSK_3_4, QAM16_1_2, QAM16_3_4, QAM64_2_3, QAM64_3_4 = range(8)
LS, LMS, COMB, STA = range(4)
LINKTYPE_IEEE802_11 = 105


class mac:
    """[UPSTREAM] ieee802_11.mac(src, dst, bss): 'app in' -> 'phy out', 'phy in' -> 'app out'."""

    def __init__(self, src_mac=(0x23,) * 6, dst_mac=(0x42,) * 6, bss_mac=(0xff,) * 6):
        for m in (src_mac, dst_mac, bss_mac):
            if len(m) != 6:
                raise ValueError("wrong mac address size")
        self.src, self.dst, self.bss = bytes(src_mac), bytes(dst_mac), bytes(bss_mac)
        self.seq = 0

    def app_in(self, payload):
        payload = bytes(payload)
        if len(payload) > 1500:
            raise ValueError("Frame too large (> 1500)")   # upstream throws std::invalid_argument
        psdu = _w.mac_frame(payload, self.seq, self.src, self.dst, self.bss)
        self.seq += 1
        return ({"crc_included": True}, psdu)

    @staticmethod
    def phy_in(pdu):
        meta, data = pdu
        if len(data) < 24:
            return None
        return (meta, data[24:])


class wifi_phy_hier:
    def __init__(self, bandwidth=10e6, chan_est=LS, encoding=BPSK_1_2, frequency=5.89e9, sensitivity=0.56,
                 device=0, max_samples=1 << 22, max_frames=0, want_carrier=False, soft_decision=False):
        self.bandwidth, self.chan_est, self.encoding = float(bandwidth), int(chan_est), int(encoding)
        self.frequency, self.sensitivity = float(frequency), float(sensitivity)
        self._h = _w.Handle(bandwidth=bandwidth, frequency=frequency, sensitivity=sensitivity, chan_est=int(chan_est),
                            encoding=int(encoding), device=device, max_samples=max_samples, max_frames=max_frames,
                            want_carrier=want_carrier, soft_decision=soft_decision)
        self._want_carrier = bool(want_carrier)
        self._mac_out_cb, self._carrier_cb = [], []
        self.samp_out = []          # bursts produced by mac_in, in order

    # --- GRC-generated accessors ---
    def get_bandwidth(self):
        return self.bandwidth

    def set_bandwidth(self, bandwidth):
        self.bandwidth = float(bandwidth)
        self._h.set_param(_w.P_BANDWIDTH, bandwidth)

    def get_chan_est(self):
        return self.chan_est

    def set_chan_est(self, chan_est):
        self.chan_est = int(chan_est)
        self._h.set_param(_w.P_CHAN_EST, int(chan_est))

    def get_encoding(self):
        return self.encoding

    def set_encoding(self, encoding):
        self.encoding = int(encoding)
        self._h.set_param(_w.P_ENCODING, int(encoding))

    def get_frequency(self):
        return self.frequency

    def set_frequency(self, frequency):
        self.frequency = float(frequency)
        self._h.set_param(_w.P_FREQUENCY, frequency)

    def get_sensitivity(self):
        return self.sensitivity

    def set_sensitivity(self, sensitivity):
        self.sensitivity = float(sensitivity)
        self._h.set_param(_w.P_SENSITIVITY, sensitivity)

    def set_soft_decision(self, on):
        """Extension (no reference counterpart): max-log LLR demapper + soft-decision Viterbi."""
        self._h.set_param(_w.P_SOFT_DECISION, 1 if on else 0)

    @property
    def handle(self):
        return self._h

    # --- ports ---
    def msg_connect_mac_out(self, callback):
        self._mac_out_cb.append(callback)

    def msg_connect_carrier(self, callback):
        self._carrier_cb.append(callback)

    def mac_in(self, pdu):
        """One PDU (meta, psdu bytes incl. FCS) -> one burst of 80*(5+N_SYM)+1 samples on samp_out."""
        _meta, psdu = pdu
        iq, _ = self._h.tx([bytes(psdu)])
        self.samp_out.append(iq)
        return iq

    def mac_in_many(self, pdus, encodings=None, seeds=None):
        iq, off = self._h.tx([bytes(p[1]) for p in pdus], enc=encodings, seed=seeds)
        bursts = [iq[int(off[i]):int(off[i + 1])] for i in range(len(pdus))]
        self.samp_out.extend(bursts)
        return bursts

    def _meta(self, f):
        total = float(f["freq_short"]) - float(f["freq_long"])
        return {"snr": float(f["snr"]), "nomfreq": self.frequency, "freqofs": total * self.bandwidth / (2 * math.pi),
                "dlt": LINKTYPE_IEEE802_11, "encoding": int(f["encoding"]), "sample_index": int(f["trigger"])}

    def samp_in(self, samples, flush=False):
        """Feed a chunk of the continuous stream; returns the PDUs decode_mac publishes."""
        self._h.rx_push(samples, flush=flush)
        out = []
        car = None
        while True:
            got = self._h.rx_pop()
            if not got:
                break
            for f, data in got:
                out.append((self._meta(f), data))
                if self._want_carrier and self._carrier_cb:
                    # frame_equalizer's 'symbols' PDUs (48 equalised points per data symbol) of the frames this
                    # push completed; the rows belong to the batch the push just ran
                    if car is None:
                        car = self._h.rows(carrier=True)[1]
                    r0, r1 = int(f["row_off"]), int(f["row_off"]) + int(f["n_rows"])
                    if 0 <= r0 and r1 <= len(car):
                        for r in range(r0, r1):
                            for cb in self._carrier_cb:
                                cb(({}, car[r]))
        for pdu in out:
            for cb in self._mac_out_cb:
                cb(pdu)
        return out

    def rx(self, samples, link_off=None, final=True):
        """Whole-capture form (one or many independent links): returns (mac_out PDUs, RxResult)."""
        res = self._h.rx_batch(samples, link_off, final=final)
        pdus = []
        for i in np.nonzero(res.frames["crc_ok"])[0]:
            pdus.append((self._meta(res.frames[i]), res.psdu(i)[:-4]))
        if self._want_carrier and self._carrier_cb:
            rows, car = self._h.rows(carrier=True)
            for i in np.nonzero(res.frames["n_rows"])[0]:
                f = res.frames[i]
                for r in range(f["row_off"], f["row_off"] + f["n_rows"]):
                    for cb in self._carrier_cb:
                        cb(({}, car[r]))
        for pdu in pdus:
            for cb in self._mac_out_cb:
                cb(pdu)
        return pdus, res
