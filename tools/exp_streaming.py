"""Throughput of the streaming entry point (wifi_b200_rx_push / rx_pop: what the GNU Radio adapter and the UDP
runner call) for different push sizes, one continuous 64-QAM 3/4 stream from host memory."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench as B
import wifi_b200 as W

B.set_workload("c3")
fpl = 2048
flen = B.frame_samples()
n_samples = B.LEAD + fpl * (flen + B.GAP)
h = W.Handle(device=0, chan_est=1, encoding=7, max_samples=n_samples + 1024, max_frames=fpl + 1024)
cap, link_off, psdus = B.build_capture(h, W, torch, 1, fpl, seed=7)
x = cap.cpu().numpy().view(np.complex64)
h.close()
for chunk, batch in ((4096, 0), (16384, 0), (65536, 0), (262144, 0), (1048576, 0), (4096, 65536), (4096, 131072), (4096, 262144), (8192, 1048576)):
    hs = W.Handle(device=0, chan_est=1, encoding=7, max_samples=2 * max(chunk, batch) + 262144, max_frames=4096)
    hs.set_param(W.wifi_b200.P_STREAM_BATCH, batch)
    n_pdu = 0
    t0 = time.perf_counter()
    lim = min(x.size, max(chunk, batch) * 24 if batch else chunk * 200)
    for p in range(0, lim, chunk):
        hs.rx_push(x[p:p + chunk], flush=(p + chunk >= lim))
        while True:
            got = hs.rx_pop()
            if not got:
                break
            n_pdu += len(got)
    dt = time.perf_counter() - t0
    print("push %8d samples, batch %8d: %8.1f Msamples/s  (%d pushes, %d PDUs, %.1f us per push)" % (chunk, batch, lim / dt / 1e6, -(-lim // chunk), n_pdu, 1e6 * dt / -(-lim // chunk)))
    hs.close()

# ---- many live links in one handle (wifi_b200_rx_push_links), samples handed over in pinned memory ----
for n_links, chunk in ((4, 131072), (16, 131072), (32, 131072), (64, 131072), (100, 131072), (100, 262144), (256, 131072)):
    per = min(x.size // chunk - 1, 12)
    hs = W.Handle(device=0, chan_est=1, encoding=7, max_samples=n_links * (chunk + 65536) + 1024, max_frames=n_links * (chunk // 6000 + 16))
    pin = torch.empty(2 * n_links * chunk, dtype=torch.float32, pin_memory=True)
    blob = pin.numpy().view(np.complex64)
    off = (np.arange(n_links + 1) * chunk).astype(np.uint64)
    n_pdu = 0
    t_fill = t_push = 0.0
    t0 = time.perf_counter()
    for k in range(per):
        tf = time.perf_counter()
        for l in range(n_links):      # every link carries the same capture, shifted so that frames do not line up (the "SDR" writing its buffer)
            blob[l * chunk:(l + 1) * chunk] = x[k * chunk + 37 * (l % 64):k * chunk + 37 * (l % 64) + chunk]
        t_fill += time.perf_counter() - tf
        tp = time.perf_counter()
        hs.rx_push_links_blob(blob, off, flush=(k == per - 1))
        if k >= 2:                     # the first pushes allocate the arena: not timed
            t_push += time.perf_counter() - tp
        else:
            t0 = time.perf_counter(); t_fill = 0.0
        while True:
            meta, pdus = hs.rx_pop_arrays(cap=8192)
            if not len(meta):
                break
            n_pdu += len(meta)
    dt = time.perf_counter() - t0 - t_fill
    per -= 2
    print("%3d links x %7d samples per push: %8.1f Msamples/s aggregate = %5.2f x real time per 20 Msps link (%d PDUs, %.1f ms per push, of which %.1f ms in wifi_b200_rx_push_links)"
          % (n_links, chunk, n_links * per * chunk / dt / 1e6, per * chunk / dt / 20e6, n_pdu, 1e3 * dt / per, 1e3 * t_push / per))
    hs.close()
