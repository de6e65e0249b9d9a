#!/usr/bin/env python
"""bench.py -- RX throughput of the 802.11a/g baseband on B200 (BASELINE.json metric:
"RX Msamples/s and decoded Mb/s per GPU, 64-QAM 3/4").

Workload (BASELINE.json configs[2], SURVEY.md 8d "C3"): synthetic 20 MHz baseband, 64-QAM 3/4,
1528-byte PSDUs (57 OFDM symbols, 4961 samples) back to back with 1100-sample idle gaps,
74 links x 512 frames per GPU (37888 frames = 148 SMs x 4 blocks x 64: whole waves of the Viterbi kernel),
AWGN at 30 dB, LMS equalizer, hard-decision Viterbi.  The capture is generated on the GPU by
the library's own TX chain and Philox channel and is far larger than L2 (no flush needed).

One "step" = one wifi_b200_rx_batch_dev pass over the whole resident capture (all links).
`value` = complex samples consumed per second, all ranks (device-resident input).
`e2e`   = the same pass through wifi_b200_rx_batch with pinned HOST input and host results.
`--impl reference` times the CPU oracle (a port of the reference algorithm; the real
gr-ieee802-11 flowgraph cannot be built here) on a bounded sample with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENC = 7
PSDU_LEN = 1528
GAP = 1100
SNR_DB = 30.0
ALGO = 1          # LMS
LEAD = 128
TX_INFO = {}
SC16_SCALE = 1.0 / 4096.0   # int16 wire format of the e2e sc16 line: full scale +-8, quantisation noise ~78 dB below the signal


def set_workload(name):
    global ENC, PSDU_LEN, GAP, SNR_DB, N_DBPS, TAPS, WORKLOAD_DESC
    if name == "c2":
        ENC, PSDU_LEN, GAP, SNR_DB, N_DBPS = 4, 1500, 1100, 25.0, 96
        TAPS = True
        WORKLOAD_DESC = "BASELINE configs[1]: one 20 Msps stream, 16-QAM 1/2, 1500-byte PSDUs (126 symbols, 10481 samples) + 1100-sample gaps, per-frame CFO, 3-tap multipath, 25 dB"


N_DBPS = 216
TAPS = False
WORKLOAD_DESC = "BASELINE configs[2]: 54 Mb/s 64-QAM 3/4 batched RX, 1528-byte PSDUs (57 symbols, 4961 samples) + 1100-sample gaps, AWGN 30 dB"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--links", type=int, default=74, help="independent links per GPU (74 x 512 frames = 148 SMs x 4 blocks x 64 trellises: whole waves of k_viterbi)")
    ap.add_argument("--frames-per-link", type=int, default=512)
    ap.add_argument("--algo", type=int, default=ALGO)
    ap.add_argument("--workload", default="c3", choices=["c3", "c2"], help="c3 = BASELINE configs[2] (metric workload); c2 = configs[1]: one 10 s 20 Msps stream, 16-QAM 1/2, CFO + 3-tap multipath")
    ap.add_argument("--soft", action="store_true", help="soft-decision mode (extension, DESIGN.md 9)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=8192, help="frames of the cpu_baseline sample")
    return ap.parse_args()


def frame_samples():
    n_sym = -(-(16 + 8 * PSDU_LEN + 6) // N_DBPS)
    return 80 * (5 + n_sym) + 1


def make_psdus(n, seed):
    import zlib
    rng = np.random.default_rng(seed)
    out = []
    hdr = bytes([0x08, 0, 0, 0]) + b"\x42" * 6 + b"\x23" * 6 + b"\xff" * 6
    pay = rng.integers(0, 256, (n, PSDU_LEN - 28), dtype=np.uint8)
    for i in range(n):
        body = hdr + int((i & 0xfff) << 4).to_bytes(2, "little") + pay[i].tobytes()
        out.append(body + zlib.crc32(body).to_bytes(4, "little"))
    return out


def build_capture(h, W, torch, n_links, fpl, seed):
    """TX + channel on the GPU.  Returns (capture tensor [2*N] f32, link_off, psdus)."""
    n = n_links * fpl
    flen = frame_samples()
    stride = flen + GAP
    link_len = LEAD + fpl * stride
    psdus = make_psdus(n, seed)
    tx = torch.empty(2 * n * flen, dtype=torch.float32, device="cuda")
    tot, off = h.tx_dev(psdus, tx.data_ptr(), n * flen, enc=ENC)
    assert tot == n * flen
    # the TX chain is not on the metric path; its rate is reported beside it (PSDU blob over PCIe, IQ stays on the device)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    h.tx_dev(psdus, tx.data_ptr(), n * flen, enc=ENC)
    torch.cuda.synchronize()
    TX_INFO.update(value=n * flen / (time.perf_counter() - t0) / 1e6, unit="Msamples/s", frames=n,
                   how="wifi_b200_tx_dev: mapper .. cyclic prefixer for every frame of the capture, host PSDUs in, device IQ out")
    cap = torch.zeros(2 * n_links * link_len, dtype=torch.float32, device="cuda")
    rng = np.random.default_rng(seed + 1)
    seg = np.zeros(n + n_links, W.wifi_b200.CHANSEG_DTYPE)
    f = np.arange(n)
    seg["in_off"][:n] = off[:-1]
    seg["in_len"][:n] = flen
    seg["out_off"][:n] = (f // fpl) * link_len + LEAD + (f % fpl) * stride
    seg["n"][:n] = stride
    # noise-only lead-in of every link
    seg["in_len"][n:] = 0
    seg["out_off"][n:] = np.arange(n_links) * link_len
    seg["n"][n:] = LEAD
    seg["n0"] = seg["out_off"]
    seg["gain"] = 0.6
    seg["noise_sigma"] = 0.6 * 10 ** (-SNR_DB / 20)
    seg["cfo"][:n] = rng.uniform(-0.0025 * 2 * np.pi, 0.0025 * 2 * np.pi, n)   # |eps| <= 50 kHz at 20 Msps
    seg["phase0"][:n] = rng.uniform(-np.pi, np.pi, n)
    seg["n_taps"] = 1
    seg["tap_re"][:, 0] = 1.0
    if TAPS:   # h = [1, 0.4 e^{j phi1}, 0, 0.2 e^{j phi2}] / |h| per frame
        p1, p2 = rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 2 * np.pi, n)
        nrm = 1.0 / np.sqrt(1 + 0.16 + 0.04)
        seg["n_taps"][:n] = 3
        seg["delay"][:n, 1], seg["delay"][:n, 2] = 1, 3
        seg["tap_re"][:n, 0] = nrm
        seg["tap_re"][:n, 1], seg["tap_im"][:n, 1] = 0.4 * nrm * np.cos(p1), 0.4 * nrm * np.sin(p1)
        seg["tap_re"][:n, 2], seg["tap_im"][:n, 2] = 0.2 * nrm * np.cos(p2), 0.2 * nrm * np.sin(p2)
    seg["seed"] = seed
    seg["stream"] = 0
    h.channel_dev(tx.data_ptr(), cap.data_ptr(), seg)
    del tx
    link_off = (np.arange(n_links + 1) * link_len).astype(np.uint64)
    return cap, link_off, psdus


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.p:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0]))
                mx.append(float(c[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if c[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def config_dict(args, n_links, fpl):
    return {"workload": WORKLOAD_DESC,
            "equalizer": ["LS", "LMS", "COMB", "STA"][args.algo], "decisions": "soft" if getattr(args, "soft", False) else "hard", "links_per_gpu": n_links, "frames_per_link": fpl,
            "samples_per_gpu": int(n_links * (LEAD + fpl * (frame_samples() + GAP))),
            "l2_policy": "input (%.2f GB per GPU) larger than L2, no flush" % (n_links * (LEAD + fpl * (frame_samples() + GAP)) * 8 / 1e9),
            "parallelism": "links sharded across GPUs, no data-path collective"}


def run_reference(args, out_fd):
    """CPU arm: the oracle (port of the reference algorithm) on all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    threads = os.cpu_count() or 1
    flen = frame_samples()
    fpl = 96
    n_links = max(threads, 1)
    rng = np.random.default_rng(0)
    # one link built with the oracle TX + channel, replicated with different noise per link
    psdus = make_psdus(fpl, 0)
    parts = [np.zeros(LEAD, np.complex64)]
    for i, p in enumerate(psdus):
        parts += [O.tx_frame(p, ENC, 1 + i % 127), np.zeros(GAP, np.complex64)]
    clean = np.concatenate(parts).astype(np.complex64)
    links = [O.channel(clean, gain=0.6, cfo=float(rng.uniform(-0.015, 0.015)), noise_sigma=0.6 * 10 ** (-SNR_DB / 20), seed=l) for l in range(n_links)]
    x = np.concatenate(links)
    off = np.arange(n_links) * clean.size
    ln = np.full(n_links, clean.size)
    times = []
    ok = 0
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = O.rx_links(x, off, ln, n_threads=threads, algo=args.algo, want_carrier=False, soft=args.soft)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
            ok = int(r.frames["crc_ok"].sum())
    t = float(np.sum(times))
    msps = x.size * len(times) / t / 1e6
    line = {"impl": "reference", "metric": "rx_msamples_per_s", "value": msps, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
            "config": config_dict(args, args.links, args.frames_per_link),
            "cpu_baseline": {"value": msps, "unit": "Msamples/s", "cores": threads, "kind": "port",
                             "sample": "%d links x %d frames (%d samples) per step, one link per host thread" % (n_links, fpl, x.size)},
            "e2e": {"value": msps, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    line["decoded_mbps"] = ok * (PSDU_LEN - 4) * 8 / (t / len(times)) / 1e6
    _emit(out_fd, line)


def _claim_stdout():
    """Libraries write banners to file descriptor 1 (NCCL prints its version there); the contract is ONE
    JSON line on stdout.  Everything else goes to stderr, the line goes to the saved descriptor."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def _emit(real_fd, line):
    sys.stdout.flush()
    os.write(real_fd, (json.dumps(line) + "\n").encode())


def main():
    args = parse()
    out_fd = _claim_stdout()
    set_workload(args.workload)
    if args.workload == "c2" and args.links == 74 and args.frames_per_link == 512:
        args.links, args.frames_per_link = 1, 17270          # 10 s at 20 Msps
    if args.impl == "reference":
        run_reference(args, out_fd)
        return
    import torch
    import torch.distributed as dist
    import wifi_b200 as pkg
    W = pkg
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_links, fpl = args.links, args.frames_per_link
    n = n_links * fpl
    flen = frame_samples()
    n_samples = n_links * (LEAD + fpl * (flen + GAP))
    h = W.Handle(device=local, chan_est=args.algo, encoding=ENC, max_samples=n_samples + 1024, max_frames=n + n_links + 1024,
                 soft_decision=args.soft)
    cap, link_off, psdus = build_capture(h, W, torch, n_links, fpl, seed=1000 + rank)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        h.rx_batch_dev(cap.data_ptr(), link_off, final=True, fetch=False)

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    barrier()
    stage_acc = {}
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()                      # synchronises its own stream before returning
        for k, v in h.stage_times().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    # the library times every stage with CUDA events on its own stream; the step time on the device is
    # their sum, the host wall clock adds launch/sync gaps -- report the larger (honest) one
    dev_ms = sum(stage_acc.values())
    elapsed = max(wall, dev_ms / 1e3)
    clocks = sampler.stop() if sampler else None
    res = h.results()
    st_frames = len(res.frames)
    st_ok = int(res.frames["crc_ok"].sum())
    # parity property at full size: every CRC-ok PDU is one of the PSDUs that were sent
    sent = {p[:-4] for p in psdus}
    pd = res.pdus()
    assert all(p in sent for p in pd[:2000]), "decoded PDU not among the transmitted ones"
    tvec = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([n_samples, st_frames, st_ok, st_ok * (PSDU_LEN - 4)], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(tvec, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)     # the only collective: counters over NCCL
    elapsed_max = float(tvec.item())
    tot_samples, tot_frames, tot_ok, tot_bytes = [int(v) for v in cnt.tolist()]
    value = tot_samples * args.steps / elapsed_max / 1e6
    mbps = tot_bytes * 8 * args.steps / elapsed_max / 1e6

    # ---- e2e: host buffers through the C ABI (H2D + D2H inside the timed region) ----
    # Two handles driven from two host threads, each taking alternate quarters of the links: the
    # library serialises capture copies per GPU, so one handle's H2D overlaps the other's kernels.
    e2e = None
    if not args.no_e2e:
        import threading
        host = torch.empty(cap.numel(), dtype=torch.float32, pin_memory=True)
        host.copy_(cap)
        torch.cuda.synchronize()
        hn = host.numpy().view(np.complex64)
        parts = 4 if n_links >= 4 else 1
        bounds = [n_links * i // parts for i in range(parts + 1)]
        part_samples = max(int(link_off[bounds[i + 1]] - link_off[bounds[i]]) for i in range(parts))
        part_frames = (max(bounds[i + 1] - bounds[i] for i in range(parts))) * (fpl + 1) + 1024
        hs = [W.Handle(device=local, chan_est=args.algo, encoding=ENC, max_samples=part_samples + 1024, max_frames=part_frames, soft_decision=args.soft) for _ in range(2)]
        tot = {"frames": 0, "store": 0, "ok": 0}
        lock = threading.Lock()

        mode = {"sc16": None}

        def worker(k, count):
            for i in range(k, parts, 2):
                if mode["sc16"] is not None:
                    hs[k].rx_batch_sc16(mode["sc16"], SC16_SCALE, link_off[bounds[i]:bounds[i + 1] + 1], final=True, fetch=False)
                else:
                    hs[k].rx_batch(hn, link_off[bounds[i]:bounds[i + 1] + 1], final=True, fetch=False)   # frames + PSDU store land in pinned host memory
                if count:
                    c = hs[k].counts()
                    with lock:
                        tot["frames"] += c["n_frames"]
                        tot["store"] += c["psdu_store_bytes"]
                        tot["ok"] += c["n_pdus"]

        def e2e_step(count=False):
            th = [threading.Thread(target=worker, args=(k, count)) for k in range(2 if parts > 1 else 1)]
            for t_ in th:
                t_.start()
            for t_ in th:
                t_.join()

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        te = time.perf_counter() - t0
        e2e_step(count=True)
        assert tot["ok"] == st_ok, (tot, st_ok)          # same answers through the host path
        # the plain single-call form, for reference
        t0 = time.perf_counter()
        for _ in range(max(1, args.steps // 2)):
            h.rx_batch(hn, link_off, final=True, fetch=False)
        t1 = (time.perf_counter() - t0) / max(1, args.steps // 2)
        tv = torch.tensor([te, t1], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        e2e = {"value": tot_samples * args.steps / float(tv[0].item()) / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(n_samples * 8), "d2h_bytes_per_step": int(tot["frames"] * 96 + tot["store"]),
               "decoded_mbps": tot_bytes * 8 * args.steps / float(tv[0].item()) / 1e6,
               "how": "wifi_b200_rx_batch on pinned host IQ, 2 handles x 2 host threads over 4 link groups (H2D of one overlaps kernels of the other); results copied to host",
               "single_call_value": tot_samples / float(tv[1].item()) / 1e6}
        # the same capture in the radio's wire format (int16 I/Q, converted on the GPU): half the PCIe bytes.
        # Reported beside the headline, not as it: the reference's samp_in port is complex float.
        try:
            h16 = torch.empty(cap.numel(), dtype=torch.int16, pin_memory=True)
            h16.copy_(torch.clamp(torch.round(cap / SC16_SCALE), -32768, 32767).to(torch.int16))
            torch.cuda.synchronize()
            mode["sc16"] = h16.numpy()
            tot.update(frames=0, store=0, ok=0)
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(max(1, args.steps // 2)):
                e2e_step()
            barrier()
            t16 = (time.perf_counter() - t0) / max(1, args.steps // 2)
            e2e_step(count=True)
            tv16 = torch.tensor([t16], dtype=torch.float64, device="cuda")
            ok16 = torch.tensor([tot["ok"]], dtype=torch.int64, device="cuda")
            if world > 1:
                dist.all_reduce(tv16, op=dist.ReduceOp.MAX)
                dist.all_reduce(ok16, op=dist.ReduceOp.SUM)
            e2e["sc16_ingest"] = {"value": tot_samples / float(tv16.item()) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(n_samples * 4),
                                  "crc_ok_per_step": int(ok16.item()),
                                  "how": "wifi_b200_rx_batch_sc16: int16 I/Q over PCIe, x = float(i16) * scale on the GPU (extension; not the headline)"}
            del h16
        except Exception as ex:     # the headline must not depend on the extension
            e2e["sc16_ingest"] = {"error": repr(ex)}
        mode["sc16"] = None
        for x_ in hs:
            x_.close()
        # live form: every link of the capture as a continuous stream, pushed chunk by chunk into ONE handle
        # (wifi_b200_rx_push_links, pinned input, bulk pop) -- what many SDR front-ends feeding one GPU look like
        try:
            chunk, pushes = 262144, 6
            link_len = int(link_off[1] - link_off[0])
            if link_len >= chunk * pushes and world >= 1:
                hl = W.Handle(device=local, chan_est=args.algo, encoding=ENC, max_samples=n_links * (chunk + 131072) + 1024,
                              max_frames=n_links * (chunk // 4096 + 16), soft_decision=args.soft)
                pin = torch.empty(2 * n_links * chunk, dtype=torch.float32, pin_memory=True)
                blob = pin.numpy().view(np.complex64)
                off = (np.arange(n_links + 1) * chunk).astype(np.uint64)
                hv = hn.reshape(n_links, link_len)
                # untimed warm-up run: the arena, the scratch buffer and the pinned result mirrors are allocated on first use
                # (and CUDA loads a kernel on its first launch: the tail-compaction kernel only runs after a push that does not flush)
                blob.reshape(n_links, chunk)[:] = hv[:, :chunk]
                for fl in (False, True):
                    hl.rx_push_links_blob(blob, off, flush=fl)
                    while len(hl.rx_pop_arrays(cap=8192)[0]):
                        pass
                hl.rx_reset()
                t_lib, n_pdu, push_ms = 0.0, 0, []
                for k in range(pushes):
                    blob.reshape(n_links, chunk)[:] = hv[:, k * chunk:(k + 1) * chunk]      # the radios filling their buffers: not timed
                    t0 = time.perf_counter()
                    hl.rx_push_links_blob(blob, off, flush=(k == pushes - 1))
                    while True:
                        meta, _pd = hl.rx_pop_arrays(cap=8192)
                        if not len(meta):
                            break
                        n_pdu += len(meta)
                    t_lib += time.perf_counter() - t0
                    push_ms.append(round(1e3 * (time.perf_counter() - t0), 2))
                e2e["streaming"] = {"value": n_links * chunk * pushes / t_lib / 1e6, "push_ms": push_ms, "unit": "Msamples/s", "links": n_links, "samples_per_push_per_link": chunk,
                                    "pushes": pushes, "pdus": n_pdu, "realtime_factor_per_20Msps_link": chunk * pushes / t_lib / 20e6,
                                    "how": "wifi_b200_rx_push_links + rx_pop per push, pinned host chunks, one handle, per-rank figure"}
                hl.close()
                del pin
        except Exception as ex:
            e2e["streaming"] = {"error": repr(ex)}
        del host

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    stage_ms = {k: v / args.steps for k, v in stage_acc.items()}
    n_jobs = st_ok if st_ok else n
    # algorithmic bytes per launch of the dominant kernel (Viterbi): N_CBPS coded bits in + PSDU bytes out per frame
    n_sym_w = -(-(16 + 8 * PSDU_LEN + 6) // N_DBPS)
    vit_alg = n * (n_sym_w * {216: 288, 96: 192}[N_DBPS] / 8 + PSDU_LEN)
    vit_ms = stage_ms.get("viterbi", 0.0)
    achieved = vit_alg / (vit_ms * 1e-3) / 1e9 if vit_ms > 0 else 0.0
    # integer work of the kernel: 64 ACS x 4 int-ops per decoded bit
    dec_bits = n * (PSDU_LEN + 2) * 8
    traffic, alu_pct = None, None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of k_viterbi from the committed ncu --set full capture (same workload)
        prof = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_full_top_kernels.json")))
        for kk in prof["kernels"]:
            if kk["kernel"] == "k_viterbi":
                u = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                traffic = sum(float(kk[m]["value"]) * u[kk[m]["unit"]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                alu_pct = float(kk["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]["value"])
    except Exception:
        pass
    roof = {"kernel": "k_viterbi", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": vit_alg, "ms_per_launch": vit_ms,
            "note": "integer-ALU-bound kernel (ncu, profiles/r01_ncu_full_top_kernels.json: sm__inst_executed_pipe_alu %s%% of peak, DRAM about 1%%): %.2f Tint-op/s algorithmic (256 int-op per decoded bit)" % ("%.0f" % alu_pct if alu_pct is not None else "?", dec_bits * 256 / (vit_ms * 1e-3) / 1e12 if vit_ms else 0.0)}
    # the streaming front-end (sync_short autocorrelation + plateau flags) is the path's HBM-side kernel: it reads the
    # whole capture once (8 B per sample) and writes one flag bit per sample
    det_ms = stage_ms.get("detect", 0.0)
    det_alg = n_samples * (8 + 1.0 / 8)
    roof_det = {"kernel": "k_detect", "bound": "hbm", "achieved": det_alg / (det_ms * 1e-3) / 1e9 if det_ms else 0.0, "peak": hbm_peak, "unit": "GB/s",
                "frac": (det_alg / (det_ms * 1e-3) / 1e9 / hbm_peak) if det_ms else 0.0, "algorithmic_bytes_per_launch": det_alg, "ms_per_launch": det_ms,
                "note": "second-largest streaming stage; issue-bound by the oracle's sequential running sums (DESIGN.md 4)"}
    path_alg = n_samples * 8 + n * PSDU_LEN
    step_ms = 1e3 * elapsed_max / args.steps
    line = {"metric": "rx_msamples_per_s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
            "decoded_mbps": mbps, "frames_per_step": tot_frames, "crc_ok_per_step": tot_ok,
            "config": config_dict(args, n_links, fpl), "clocks": clocks, "e2e": e2e, "gpu_launches": 13 * args.steps,   # detect, select_spec, select, reserve, frames_init, sync_long, demod x2, signal, plan_fast, plan, pack, viterbi
            "roofline": roof, "roofline_frontend": roof_det, "stage_ms": stage_ms, "tx": dict(TX_INFO),
            "path_hbm": {"algorithmic_GBps": path_alg / (step_ms * 1e-3) / 1e9, "frac_of_peak": path_alg / (step_ms * 1e-3) / 1e9 / hbm_peak}}
    if not args.no_cpu and world == 1:
        from oracle import oracle as O
        nf = min(args.cpu_frames, fpl * n_links)
        nl = max(1, nf // fpl)
        sl = int(link_off[nl])
        xs = cap[:2 * sl].cpu().numpy().view(np.complex64)
        t0 = time.perf_counter()
        r = O.rx_links(xs, link_off[:nl].astype(np.int64), np.diff(link_off[:nl + 1]).astype(np.int64), n_threads=1, algo=args.algo, want_carrier=False,
                       soft=args.soft)
        dt = time.perf_counter() - t0
        # same inputs, same answers: the GPU frame table of these links equals the oracle's
        g = res.frames[np.isin(res.frames["link"], np.arange(nl))]
        g = g[np.lexsort((g["trigger"], g["link"]))]
        same = len(g) == len(r.frames) and all(np.array_equal(g[k], r.frames[k]) for k in ("trigger", "frame_start", "encoding", "length", "crc_ok"))
        line["cpu_baseline"] = {"value": sl / dt / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "port",
                                "sample": "first %d links (%d frames, %d samples) of the same capture, oracle single thread, %.1f s" % (nl, nl * fpl, sl, dt),
                                "matches_gpu_frame_table": bool(same)}
    _emit(out_fd, line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
