"""Experiment: do the ALU-bound Viterbi of one link group and the FMA/issue-bound demod of another overlap when
two handles (two streams) run concurrently?  Compares one handle over all links with G handles over link groups,
free-running from G host threads (optionally staggered).

Result on a B200 (round 1): no.  One handle over 74 links: 6.85 ms per step; 2 / 3 / 4 handles over link groups:
8.0 / 7.6 / 7.7 ms.  A Viterbi launch takes about 3 ms however few frames it holds (one serial trellis per thread),
so splitting the batch only adds latency-bound tails; the batch stays whole."""
import sys, time, threading
import numpy as np, torch
sys.path.insert(0, ".")
import bench as B
import wifi_b200 as W

B.set_workload("c3")
n_links, fpl = 74, 512
flen = B.frame_samples()
n_samples = n_links * (B.LEAD + fpl * (flen + B.GAP))
h = W.Handle(device=0, chan_est=1, encoding=7, max_samples=n_samples + 1024, max_frames=n_links * fpl + n_links + 1024)
cap, link_off, psdus = B.build_capture(h, W, torch, n_links, fpl, seed=1000)
torch.cuda.synchronize()
def run_single(steps):
    t0 = time.perf_counter()
    for _ in range(steps):
        h.rx_batch_dev(cap.data_ptr(), link_off, final=True, fetch=False)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps
run_single(3)
print("single handle: %.3f ms/step" % (1e3 * run_single(10)))
for G in (2, 3, 4):
    bounds = [n_links * i // G for i in range(G + 1)]
    hs = [W.Handle(device=0, chan_est=1, encoding=7, max_samples=int(link_off[bounds[i + 1]] - link_off[bounds[i]]) + 1024,
                   max_frames=(bounds[i + 1] - bounds[i]) * (fpl + 1) + 1024) for i in range(G)]
    def worker(i, steps, delay):
        time.sleep(delay)
        lo = link_off[bounds[i]:bounds[i + 1] + 1]
        for _ in range(steps):
            hs[i].rx_batch_dev(cap.data_ptr(), lo, final=True, fetch=False)
    for stagger in (0.0, 1.0):
        def go(steps):
            th = [threading.Thread(target=worker, args=(i, steps, stagger * i * 0.007 / G)) for i in range(G)]
            t0 = time.perf_counter()
            for t in th: t.start()
            for t in th: t.join()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0 - stagger * (G - 1) * 0.007 / G) / steps
        go(3)
        print("G=%d handles, stagger %.0f: %.3f ms per full step" % (G, stagger, 1e3 * go(20)))
    ok = sum(int(x.results().frames["crc_ok"].sum()) for x in hs)
    print("   crc_ok", ok)
    for x in hs: x.close()
