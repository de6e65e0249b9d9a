"""GNU Radio adapter: the `wifi_phy_hier` ports as real GNU Radio blocks.

Importing this module needs GNU Radio 3.10 (`gnuradio.gr`, `pmt`); the rest of the package does not.
The reference builds its PHY as a GRC hierarchical block (gnu_radio/wifi_phy_hier.grc, id `wifi_phy_hier`
:14, category `[IEEE802.11]` :5) with

    stream in  `samp_in`  (:641-660)   message in  `mac_in`  (:661-680)
    stream out `samp_out` (:587-604)   message out `mac_out` (:623-640), `carrier` (:605-622)

and instantiates it as `wifi_phy_hier(bandwidth=, chan_est=, encoding=, frequency=, sensitivity=)`
(gnu_radio/IRS_tranceiver.py:178-184).  `wifi_phy_hier_b200` below is a `gr.hier_block2` with the same
constructor keywords, port names and setters, so a flowgraph swaps the import and nothing else:

    from wifi_b200.gr_adapter import wifi_phy_hier_b200 as wifi_phy_hier

Inside, two blocks stand where the reference wires ~25:
  * `wifi_tx_b200`  message `mac_in` -> stream out (one burst of 80*(5+N_SYM)+1 samples per PDU, `packet_len`
    tag on its first sample -- what mapper ... ofdm_cyclic_prefixer produce, wifi_phy_hier.grc:749-765);
  * `wifi_rx_b200`  stream in -> messages `mac_out` (decode_mac 'out': cons(dict{snr, nomfreq, freqofs, dlt},
    u8vector)) and `carrier` (frame_equalizer 'symbols': cons(dict(), c32vector[48])) (:736-748,755-759).
Both call the same facade (`wifi_phy_hier.py`) that the GNU-Radio-free users call; all arithmetic runs in
libwifi_b200.so.  `grc/ieee802_11_wifi_phy_hier_b200.block.yml` makes the block appear in GRC.
"""
import collections
import queue
import threading

import numpy as np

try:
    import pmt
    from gnuradio import gr
except ImportError as e:  # pragma: no cover - exercised by tests/test_gr_adapter.py with and without stubs
    raise ImportError("wifi_b200.gr_adapter needs GNU Radio 3.10 (gnuradio.gr, pmt); "
                      "use wifi_b200.wifi_phy_hier without GNU Radio") from e

from . import wifi_b200 as _w
from .wifi_phy_hier import wifi_phy_hier as _facade

_MAC_IN, _MAC_OUT, _CARRIER = "mac_in", "mac_out", "carrier"


def pdu_to_python(msg):
    """pmt.cons(dict, u8vector) -> (dict, bytes)"""
    meta = pmt.to_python(pmt.car(msg)) if pmt.is_dict(pmt.car(msg)) else {}
    return (meta or {}, bytes(bytearray(pmt.u8vector_elements(pmt.cdr(msg)))))


def pdu_from_python(meta, data):
    """(dict, bytes) -> pmt.cons(dict, u8vector), the shape decode_mac publishes"""
    d = pmt.make_dict()
    for k, v in meta.items():
        d = pmt.dict_add(d, pmt.intern(str(k)), pmt.from_long(int(v)) if isinstance(v, (int, np.integer)) else pmt.from_double(float(v)))
    return pmt.cons(d, pmt.init_u8vector(len(data), list(bytearray(data))))


class wifi_tx_b200(gr.basic_block):
    """message `mac_in` -> complex stream with a `packet_len` tag per burst"""

    def __init__(self, phy):
        gr.basic_block.__init__(self, name="wifi_tx_b200", in_sig=None, out_sig=[np.complex64])
        self._phy = phy
        self._q = collections.deque()      # [burst ndarray, samples already written]
        self._lock = threading.Lock()
        self._ready = threading.Event()    # set while bursts wait: an idle work() call sleeps on it instead of spinning
        self.message_port_register_in(pmt.intern(_MAC_IN))
        self.set_msg_handler(pmt.intern(_MAC_IN), self.handle_mac_in)

    def handle_mac_in(self, msg):
        _meta, psdu = pdu_to_python(msg)
        burst = self._phy.mac_in(({}, psdu))        # raises on oversize PSDUs like upstream's mapper
        self._phy.samp_out.clear()
        with self._lock:
            self._q.append([burst, 0])
            self._ready.set()

    def general_work(self, input_items, output_items):
        out = output_items[0]
        produced = 0
        # A source block that returns 0 is called again at once by GNU Radio's thread-per-block scheduler: with no PDU
        # queued, wait (bounded, so stop() is honoured) the way pdu_to_tagged_stream blocks on its message queue.
        if not self._q:
            self._ready.wait(0.01)
        with self._lock:
            while self._q and produced < len(out):
                burst, done = self._q[0]
                if done == 0:
                    self.add_item_tag(0, self.nitems_written(0) + produced, pmt.intern("packet_len"), pmt.from_long(len(burst)))
                n = min(len(burst) - done, len(out) - produced)
                out[produced:produced + n] = burst[done:done + n]
                produced += n
                if done + n == len(burst):
                    self._q.popleft()
                else:
                    self._q[0][1] = done + n
            if not self._q:
                self._ready.clear()
        return produced


class wifi_rx_b200(gr.basic_block):
    """complex stream -> messages `mac_out` and (if connected) `carrier`.

    `general_work` only hands the samples to a worker thread, so the scheduler thread never stalls (a live SDR
    source overflows if its consumer blocks for milliseconds).  The worker pushes them into the library, which
    buffers until `stream_batch` new samples wait (WIFI_P_STREAM_BATCH): one pipeline run costs 1-3 ms whatever
    its size, 131072 samples are 6.5 ms of a 20 Msps stream."""

    def __init__(self, phy, stream_batch=131072):
        gr.basic_block.__init__(self, name="wifi_rx_b200", in_sig=[np.complex64], out_sig=None)
        self._phy = phy
        self.message_port_register_out(pmt.intern(_MAC_OUT))
        self.message_port_register_out(pmt.intern(_CARRIER))
        phy.msg_connect_carrier(self._publish_carrier)
        phy.handle.set_param(_w.P_STREAM_BATCH, int(stream_batch))
        self._q = queue.Queue()
        self._worker = None
        self._error = None

    def _publish_carrier(self, pdu):
        _meta, pts = pdu
        self.message_port_pub(pmt.intern(_CARRIER), pmt.cons(pmt.make_dict(), pmt.init_c32vector(len(pts), [complex(v) for v in pts])))

    def _publish(self, pdus):
        for meta, mpdu in pdus:
            meta = {k: meta[k] for k in ("snr", "nomfreq", "freqofs", "dlt", "encoding")}     # upstream decode_mac's dict (+ the MCS)
            self.message_port_pub(pmt.intern(_MAC_OUT), pdu_from_python(meta, mpdu))

    def _run(self):
        try:
            idle = True
            while True:
                try:
                    x = self._q.get(timeout=0.05)
                except queue.Empty:
                    if not idle:                       # the stream paused: decode what is buffered (empty push = run now)
                        self._publish(self._phy.samp_in(np.zeros(0, np.complex64)))
                        idle = True
                    continue
                if x is None:
                    break
                idle = False
                self._publish(self._phy.samp_in(x))
            self._publish(self._phy.samp_in(np.zeros(0, np.complex64), flush=True))   # the newest burst is held until flushed
        except Exception as e:   # surfaced by the next general_work / stop
            self._error = e

    def start(self):
        if self._worker is None:
            self._worker = threading.Thread(target=self._run, name="wifi_rx_b200", daemon=True)
            self._worker.start()
        return True

    def general_work(self, input_items, output_items):
        if self._error is not None:
            raise self._error
        if self._worker is None:
            self.start()
        x = input_items[0]
        if len(x):
            self._q.put(np.array(x, dtype=np.complex64))      # copy: the scheduler reuses its buffer
            self.consume(0, len(x))
        return 0

    def stop(self):
        if self._worker is not None:
            self._q.put(None)
            self._worker.join()
            self._worker = None
        if self._error is not None:
            raise self._error
        return True


class wifi_phy_hier_b200(gr.hier_block2):
    """Drop-in for the GRC-generated `wifi_phy_hier` (same keywords, ports and setters)."""

    def __init__(self, bandwidth=10e6, chan_est=0, encoding=0, frequency=5.89e9, sensitivity=0.56, device=0, max_samples=1 << 22,
                 want_carrier=False, stream_batch=131072):
        gr.hier_block2.__init__(self, "WiFi PHY Hier (B200)",
                                gr.io_signature(1, 1, gr.sizeof_gr_complex * 1),
                                gr.io_signature(1, 1, gr.sizeof_gr_complex * 1))
        self.message_port_register_hier_in(_MAC_IN)
        self.message_port_register_hier_out(_MAC_OUT)
        self.message_port_register_hier_out(_CARRIER)
        self.bandwidth, self.chan_est, self.encoding = bandwidth, int(chan_est), int(encoding)
        self.frequency, self.sensitivity = frequency, sensitivity
        self._phy = _facade(bandwidth=bandwidth, chan_est=int(chan_est), encoding=int(encoding), frequency=frequency,
                            sensitivity=sensitivity, device=device, max_samples=max_samples, want_carrier=want_carrier)
        self.tx = wifi_tx_b200(self._phy)
        self.rx = wifi_rx_b200(self._phy, stream_batch)
        self.connect((self, 0), (self.rx, 0))                                  # samp_in   (wifi_phy_hier.grc:762-764)
        self.connect((self.tx, 0), (self, 0))                                  # samp_out  (:752)
        self.msg_connect((self, _MAC_IN), (self.tx, _MAC_IN))                  # mac_in    (:765)
        self.msg_connect((self.rx, _MAC_OUT), (self, _MAC_OUT))                # mac_out   (:757)
        self.msg_connect((self.rx, _CARRIER), (self, _CARRIER))                # carrier   (:759)

    # the accessors GRC generates for a hier block's parameters (used at IRS_tranceiver.py:386,427,442)
    def get_bandwidth(self):
        return self.bandwidth

    def set_bandwidth(self, bandwidth):
        self.bandwidth = bandwidth
        self._phy.set_bandwidth(bandwidth)

    def get_chan_est(self):
        return self.chan_est

    def set_chan_est(self, chan_est):
        self.chan_est = int(chan_est)
        self._phy.set_chan_est(int(chan_est))

    def get_encoding(self):
        return self.encoding

    def set_encoding(self, encoding):
        self.encoding = int(encoding)
        self._phy.set_encoding(int(encoding))

    def get_frequency(self):
        return self.frequency

    def set_frequency(self, frequency):
        self.frequency = frequency
        self._phy.set_frequency(frequency)

    def get_sensitivity(self):
        return self.sensitivity

    def set_sensitivity(self, sensitivity):
        self.sensitivity = sensitivity
        self._phy.set_sensitivity(sensitivity)
