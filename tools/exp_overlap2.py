"""Experiment (round 2): two handles, each with a WHOLE batch (74 links x 512 frames), free-running from two host
threads on one GPU.  If the ALU-bound Viterbi of one batch overlaps the HBM / FMA-bound front-end of the other, two
batches finish in less than twice the single-batch time -- the case for an asynchronous rx_batch that pipelines
consecutive batches.  Prints ms per batch for 1 handle and for 2 concurrent handles (with and without a half-step stagger)."""
import sys, time, threading
import numpy as np, torch
sys.path.insert(0, ".")
import bench as B
import wifi_b200 as W

B.set_workload("c3")
n_links, fpl = 74, 512
flen = B.frame_samples()
n_samples = n_links * (B.LEAD + fpl * (flen + B.GAP))
hs = [W.Handle(device=0, chan_est=1, encoding=7, max_samples=n_samples + 1024, max_frames=n_links * fpl + n_links + 1024) for _ in range(2)]
cap, link_off, psdus = B.build_capture(hs[0], W, torch, n_links, fpl, seed=1000)
torch.cuda.synchronize()


def run(i, steps, delay=0.0):
    time.sleep(delay)
    for _ in range(steps):
        hs[i].rx_batch_dev(cap.data_ptr(), link_off, final=True, fetch=False)


run(0, 3); run(1, 3)
t0 = time.perf_counter(); run(0, 20); torch.cuda.synchronize()
one = (time.perf_counter() - t0) / 20
print("one handle: %.3f ms per batch" % (1e3 * one))
for stagger in (0.0, 0.003):
    th = [threading.Thread(target=run, args=(i, 20, stagger * i)) for i in range(2)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    two = (time.perf_counter() - t0 - stagger) / 40
    print("two handles (stagger %.1f ms): %.3f ms per batch = %.2fx the single-handle throughput" % (1e3 * stagger, 1e3 * two, one / two))
print("crc_ok", [int(x.results().frames["crc_ok"].sum()) for x in hs])
