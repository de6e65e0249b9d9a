"""BASELINE.json configs[0], [1] and [4] at their stated sizes on the GPU, each compared with the CPU oracle over the
whole workload (frame table and every PSDU byte) -- the tests in test_gpu_parity.py cover the same paths on small shapes.

configs[0]  IRS_tranceiver loopback: kodim01.png cut into 1500-byte PSDUs (352 frames), BPSK 1/2, 20 MHz, AWGN 20 dB
configs[1]  one 10 s 20 Msps stream (200 Msamples), 16-QAM 1/2, per-frame CFO, 3-tap multipath, 25 dB
configs[4]  feature-map payloads (detach_image pieces of a (30, 30, 128) float32 map per Kodak image), all 8 MCS x an
            SNR ladder, packet-error-rate table identical to the oracle's
"""
import json
import os
import pickle
import struct

import numpy as np
import pytest

from util import assert_frames_equal

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_config0_kodim01_in_1500_byte_psdus_bpsk_20db(O, W):
    """/root/reference/images/kodim01.png (fixture tests/golden/kodim01_payload.bin) -> ieee802_11.mac framing (addresses of
    IRS_tranceiver.py:271, sequence numbers from 0) -> 1500-byte PSDUs -> TX on the GPU (mapper's running scrambler seed)
    -> x0.6, packet_pad2(100, 1000) (IRS_tranceiver.py:277,295) -> AWGN at 20 dB -> RX on the GPU.  Every PDU comes back,
    the file is rebuilt from the `data[24:]` slices, and the frame table is the oracle's."""
    import hashlib
    meta = json.load(open(os.path.join(HERE, "golden", "kodim01_payload.json")))
    data = open(os.path.join(HERE, "golden", "kodim01_payload.bin"), "rb").read()
    assert hashlib.sha256(data).hexdigest() == meta["sha256"] and len(data) == 517565
    phy = W.wifi_phy_hier(bandwidth=20e6, chan_est=0, encoding=0, frequency=5.89e9, sensitivity=0.56, max_samples=1 << 24, max_frames=1024)
    m = W.mac([0x23] * 6, [0x42] * 6, [0xff] * 6)
    pdus = [m.app_in(data[i:i + 1472]) for i in range(0, len(data), 1472)]
    assert len(pdus) == 352 and all(len(p[1]) == 1500 for p in pdus[:-1])
    bursts = phy.mac_in_many(pdus)
    assert bursts[0].size == 40481 and np.array_equal(bursts[5], O.tx_frame(pdus[5][1], 0, seed=6))
    parts = []
    for b in bursts:
        parts += [np.zeros(100, np.complex64), b, np.zeros(1000, np.complex64)]
    x = np.concatenate(parts)
    y = O.channel(x, gain=0.6, noise_sigma=0.6 * 10 ** (-20 / 20), seed=0)
    got, res = phy.rx(y)
    ref = O.rx(y, algo=0, want_carrier=False, bw=20e6, freq=5.89e9)
    assert_frames_equal(res, ref)
    assert len(got) == 352 and all(g[0]["encoding"] == 0 for g in got)
    assert b"".join(g[1][24:] for g in got) == data
    phy.handle.close()


def test_config1_full_10s_capture_16qam_multipath_cfo(O, W):
    """The bench's `--workload c2` capture (TX + Philox channel on the GPU: 17270 frames, 200 Msamples) decoded in one call
    and by the oracle over the whole stream on one host core (about 20 s): identical tables, every frame's FCS good."""
    import torch
    import bench
    saved = (bench.ENC, bench.PSDU_LEN, bench.GAP, bench.SNR_DB, bench.N_DBPS, bench.TAPS)
    try:
        bench.set_workload("c2")
        fpl = 17270
        n = bench.LEAD + fpl * (bench.frame_samples() + bench.GAP)
        assert n >= 199_900_000
        h = W.Handle(chan_est=0, encoding=bench.ENC, max_samples=n + 1024, max_frames=fpl + 1024)
        cap, link_off, psdus = bench.build_capture(h, W, torch, 1, fpl, seed=1)
        res = h.rx_batch_dev(cap.data_ptr(), link_off, final=True, fetch=True)
        y = cap.cpu().numpy().view(np.complex64)
        del cap
        ref = O.rx(y, algo=0, want_carrier=False)
        assert_frames_equal(res, ref)
        assert len(res.frames) == fpl and int(res.frames["crc_ok"].sum()) == fpl
        assert res.pdus() == [p[:-4] for p in psdus]
        h.close()
    finally:
        bench.ENC, bench.PSDU_LEN, bench.GAP, bench.SNR_DB, bench.N_DBPS, bench.TAPS = saved


def latent_of(img):
    """img2msg's latent -- a (30, 30, 128) float32 map per image; the trained weights are absent from the reference, so a
    seeded surrogate of the same shape and dtype stands in."""
    return np.random.default_rng(1000 + img).standard_normal((30, 30, 128)).astype(np.float32)


def feature_map_payloads(n_images=6):
    """What upload_featuremap_udp.py puts on the wire (:32-48) for each image: the latent cut by detach_image
    (image_detach_rebuild.py:6-32) into (10, 10, 1) pieces with their (y, x, c) positions, shuffled (seeded here), each sent
    as `=L` length + pickle -- through the package's featuremap module."""
    import wifi_b200
    fm = wifi_b200.featuremap
    out = []
    for img in range(n_images):
        out += [fm.to_datagram(p) for p in fm.detach(latent_of(img), random_state=img)]
    return out


def build_ladder(W, h, pay, mcs_list, snrs, fpp, seed=44):
    """One link per (MCS, SNR) point: lead-in noise, then fpp frames (mac framing of the datagrams, TX on the GPU), each
    followed by a gap; per-frame CFO; AWGN of the point's SNR from the Philox channel.  Returns (device capture, link_off,
    link_len, psdus)."""
    import torch
    m = W.mac()
    gap, lead = 1100, 128
    psdus, encs = [], []
    k = 0
    for enc in mcs_list:
        for _ in snrs:
            for _ in range(fpp):
                psdus.append(m.app_in(pay[k % len(pay)])[1])
                encs.append(enc)
                k += 1
    encs = np.array(encs, np.uint8)
    lens = np.array([W.wifi_b200.frame_samples(int(e), len(p)) for e, p in zip(encs, psdus)], np.int64)
    n, n_links = len(psdus), len(mcs_list) * len(snrs)
    tx = torch.empty(2 * int(lens.sum()), dtype=torch.float32, device="cuda")
    tot, boff = h.tx_dev(psdus, tx.data_ptr(), int(lens.sum()), enc=encs, seed=(np.arange(n) % 127 + 1).astype(np.uint8))
    assert tot == lens.sum()
    link_id = np.arange(n) // fpp
    stride = lens + gap
    link_len = np.array([lead + stride[link_id == l].sum() for l in range(n_links)], np.int64)
    link_off = np.concatenate([[0], np.cumsum(link_len)]).astype(np.uint64)
    out_off = np.zeros(n, np.int64)
    for l in range(n_links):
        idx = np.nonzero(link_id == l)[0]
        out_off[idx] = int(link_off[l]) + lead + np.concatenate([[0], np.cumsum(stride[idx])[:-1]])
    snr_of_link = np.array(snrs, np.float64)[np.arange(n_links) % len(snrs)]
    seg = np.zeros(n + n_links, W.wifi_b200.CHANSEG_DTYPE)
    seg["in_off"][:n], seg["in_len"][:n], seg["out_off"][:n], seg["n"][:n] = boff[:-1], lens, out_off, stride
    seg["noise_sigma"][:n] = 0.6 * 10 ** (-snr_of_link[link_id] / 20)
    seg["out_off"][n:], seg["n"][n:] = link_off[:-1], lead
    seg["noise_sigma"][n:] = 0.6 * 10 ** (-snr_of_link / 20)
    seg["n0"] = seg["out_off"]
    seg["gain"], seg["n_taps"], seg["seed"] = 0.6, 1, seed
    seg["tap_re"][:, 0] = 1.0
    seg["cfo"][:n] = np.random.default_rng(5).uniform(-0.01, 0.01, n)
    cap = torch.zeros(2 * int(link_off[-1]), dtype=torch.float32, device="cuda")
    h.channel_dev(tx.data_ptr(), cap.data_ptr(), seg)
    del tx
    return cap, link_off, link_len, psdus


def per_table(frames, n_mcs, n_snr, fpp):
    ok = np.zeros(n_mcs * n_snr, np.int64)
    np.add.at(ok, frames["link"][frames["crc_ok"] == 1], 1)
    return 1.0 - ok.reshape(n_mcs, n_snr) / fpp


def test_config4_per_ladder_all_mcs_feature_map_payloads(O, W):
    """8 MCS x 16 SNRs (0, 2, .. 30 dB) x 200 frames: 25600 feature-map datagrams through mac framing, TX, the Philox
    channel and RX on the GPU; one link per (MCS, SNR) point.  The oracle decodes the same 128 links on the host cores.
    The two PER tables are the same table, frame for frame; the ladder has its waterfall (PER 1 at the bottom of every
    column, 0 at the top) and every delivered piece unpickles to the piece that was sent."""
    pay = feature_map_payloads()
    assert len(pay) == 6 * 1152 and 500 < len(pay[0]) < 700
    snrs = list(range(0, 32, 2))
    fpp = 200
    h = W.Handle(chan_est=0, max_samples=1 << 28, max_frames=8 * 16 * fpp + 4096)
    cap, link_off, link_len, psdus = build_ladder(W, h, pay, list(range(8)), snrs, fpp)
    res = h.rx_batch_dev(cap.data_ptr(), link_off, final=True, fetch=True)
    y = cap.cpu().numpy().view(np.complex64)
    del cap
    ref = O.rx_links(y, link_off[:-1].astype(np.int64), link_len, n_threads=os.cpu_count() or 1, algo=0, want_carrier=False)
    assert_frames_equal(res, ref)
    per = per_table(res.frames, 8, len(snrs), fpp)
    assert np.array_equal(per, per_table(ref.frames, 8, len(snrs), fpp))
    assert (per[:, -1] <= 0.02).all() and per[7, 0] == 1.0 and per[4:, 1].min() == 1.0, per
    assert all(np.all(np.diff(per[e]) <= 0.05) for e in range(8)), per          # monotone waterfall (binomial slack)
    # the waterfall moves to higher SNR with the rate: the first SNR with PER < 10 % does not decrease along 1/2-rate MCS
    first_ok = [int(np.argmax(per[e] < 0.1)) for e in (0, 2, 4)]
    assert first_ok == sorted(first_ok) and first_ok[0] < first_ok[2], first_ok
    # delivered payloads are the datagrams that were sent (the consumer's slice, IRS_tranceiver_epy_block_2.py:34-35)
    sent = {p[24:-4] for p in psdus}
    pdus = res.pdus()
    assert len(pdus) == int((res.frames["crc_ok"] == 1).sum()) and all(p[24:] in sent for p in pdus[::97])
    (pos, piece) = pickle.loads(pdus[-1][24:][4:])
    assert piece.shape == (10, 10, 1) and piece.dtype == np.float32 and len(pos) == 3
    # the viewer's side (download_featuremap_udp.py:53-69): at the top of the ladder every piece of the links' images arrives
    # and rebuild_image gives the latent back; lower down the map has holes where frames were lost
    fm = W.featuremap
    top = res.frames["link"] == 7 * len(snrs) + len(snrs) - 1                 # 64-QAM 3/4 at 30 dB
    got = [fm.from_datagram(res.psdu(i)[:-4][24:][4:]) for i in np.nonzero(top & (res.frames["crc_ok"] == 1))[0]]
    first = (7 * len(snrs) + len(snrs) - 1) * fpp
    sent_here = [fm.from_datagram(pay[(first + j) % len(pay)][4:]) for j in range(fpp)]
    assert len(got) >= fpp - 4
    img = (first % len(pay)) // 1152
    assert (first % len(pay)) % 1152 + fpp <= 1152                            # all of this link's pieces belong to one image
    rebuilt, want = fm.rebuild(got, (30, 30, 128)), fm.rebuild(sent_here, (30, 30, 128))
    mask = rebuilt != 0
    assert mask.sum() >= 0.97 * (want != 0).sum() and np.array_equal(rebuilt[mask], want[mask])
    assert np.array_equal(want[want != 0], latent_of(img)[want != 0])
    h.close()


def test_config4_soft_decisions_over_the_ladder(O, W):
    """The same ladder, reduced (3 MCS x 8 SNRs x 100 frames), in soft-decision mode (extension: the oracle defines it):
    tables identical to the oracle's, and the soft decoder never loses to the hard one by more than binomial noise while
    winning clearly wherever the SNR grid samples a column's waterfall."""
    pay = feature_map_payloads(1)
    mcs, snrs, fpp = [1, 4, 7], [2, 6, 10, 14, 18, 22, 26, 30], 100
    per = {}
    for soft in (False, True):
        h = W.Handle(chan_est=0, max_samples=1 << 25, max_frames=len(mcs) * len(snrs) * fpp + 2048, soft_decision=soft)
        cap, link_off, link_len, _ = build_ladder(W, h, pay, mcs, snrs, fpp, seed=45)
        res = h.rx_batch_dev(cap.data_ptr(), link_off, final=True, fetch=True)
        y = cap.cpu().numpy().view(np.complex64)
        del cap
        ref = O.rx_links(y, link_off[:-1].astype(np.int64), link_len, n_threads=os.cpu_count() or 1, algo=0, want_carrier=False, soft=soft)
        assert_frames_equal(res, ref)
        per[soft] = per_table(res.frames, len(mcs), len(snrs), fpp)
        h.close()
    assert (per[True] <= per[False] + 0.08).all(), (per[False], per[True])
    # (a 4 dB grid does not sample every column's waterfall: BPSK 3/4 falls between 2 and 6 dB)
    assert sum((per[False][e] - per[True][e]).max() >= 0.3 for e in range(len(mcs))) >= 2, (per[False], per[True])


@pytest.mark.parametrize("fpp", [4760, 5000])
def test_frames_beyond_whole_viterbi_waves_are_decoded_beside_them(W, fpp):
    """A call with more frames than one wave of the per-thread Viterbi kernel (148 SMs x 4 blocks x 64 = 37888) decodes
    the whole waves with that kernel and the remainder (192 frames: one trellis per warp; 2112: per four lanes) in a
    second launch beside it.  Same frame table and the same PSDU bytes as ONE launch of the per-thread kernel over all."""
    import hashlib
    rng = np.random.default_rng(fpp)
    pay = [rng.integers(0, 256, int(rng.integers(20, 70)), dtype=np.uint8).tobytes() for _ in range(997)]
    h = W.Handle(chan_est=0, max_samples=1 << 28, max_frames=8 * fpp + 4096)
    try:
        cap, link_off, _, _ = build_ladder(W, h, pay, list(range(8)), [14.0], fpp)        # marginal SNR for the upper MCS: failing frames too
        tables, stores = [], []
        for form in (0, 3):
            h.set_param(W.wifi_b200.P_VITERBI_FORM, form)
            res = h.rx_batch_dev(cap.data_ptr(), link_off, final=True, fetch=True)
            assert len(res.frames) >= 8 * fpp - 50 and len(res.frames) % 37888 not in (0,)
            tables.append(res.frames.copy())
            stores.append(hashlib.sha256(b"".join(res.psdu(i) for i in np.nonzero(res.frames["decoded"])[0])).hexdigest())
        differing = [k for k in tables[0].dtype.names if not np.array_equal(tables[0][k], tables[1][k], equal_nan=True)]
        assert not differing, [(k, np.nonzero(tables[0][k] != tables[1][k])[0][:8]) for k in differing]
        assert stores[0] == stores[1]
        ok = tables[0]["crc_ok"].sum()
        assert 0.5 * 8 * fpp < ok < 8 * fpp - 100, ok
    finally:
        h.close()
