"""Multi-rank host logic on CPU (gloo, world_size 2): link sharding, overlapping-segment sharding
with core-region dedup, and the counter all-reduce.  The per-rank compute here is the oracle --
this tests the plumbing, the GPU path is covered by the -m gpu tests."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, kind, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    import torch.distributed as dist
    from oracle import oracle as O
    from util import make_capture
    S = importlib.import_module("gnuradio-wifi-imagetransfer_b200.sharding")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    if kind == "links":
        caps = [make_capture(O, rng, [(int(rng.integers(0, 8)), 80)] * 3, snr_db=30, seed=l)[0] for l in range(5)]
        mine = S.shard_links(len(caps), world, rank)
        fr = [O.rx(caps[l], link=l, algo=0, want_carrier=False).frames for l in mine]
        frames = np.concatenate(fr)
        n_samp = sum(caps[l].size for l in mine)
        total = np.concatenate([O.rx(c, link=l, algo=0, want_carrier=False).frames for l, c in enumerate(caps)])
        ref = S.stats_vector(total, sum(c.size for c in caps))
    else:
        from util import adversarial_stream, oracle_segment_decoder
        tails = kind.startswith("tail")          # the form bench.py runs: only the last owned records travel, in fixed-size buffers
        if kind == "stream":
            y, _ = make_capture(O, rng, [(2, 400)] * 40, snr_db=30, seed=9, gap=900)
        else:
            y = adversarial_stream(O, trunc=int(kind[4:] if tails else kind[3:]))
        seg = S.shard_stream(y.size, world, overlap=S.OVERLAP)[rank]
        full = O.rx(y, algo=0, want_carrier=False).frames
        # every rank decodes its segment, the frame records are all-gathered (gloo here, NCCL on the GPUs), every rank
        # checks that it joined the sequential receiver's state and decodes again from a known state if it did not
        trace = {}
        own, owned_all, rounds = S.reconcile(oracle_segment_decoder(O, y, algo=0), S.shard_stream(y.size, world), rank, y.size,
                                             tail_rows=64 if tails else None, trace=trace)
        if tails:
            assert all(len(t) <= 64 for t in owned_all) and {"decode", "exchange", "check"} <= set(trace)
            owned_all = S.gather_owned(own)
        assert S.same(np.concatenate(owned_all), S.records(full.copy(), 0)), (kind, rank)
        assert (rounds >= 1) == (kind != "stream"), (kind, rounds)      # idle gaps: joined at once; adversarial: one more pass
        frames = np.zeros(len(own), full.dtype)
        for k, col in (("sig_ok", S.SIG), ("decoded", S.DEC), ("crc_ok", S.CRC), ("length", S.LEN), ("encoding", S.ENC)):
            frames[k] = own[:, col]
        n_samp = seg["core_end"] - seg["core_start"]
        ref = S.stats_vector(full, y.size)
        np.save(os.path.join(out_dir, "trig%d.npy" % rank), own[:, S.TRIG])
        if rank == 0:
            np.save(os.path.join(out_dir, "trig_full.npy"), full["trigger"])
    got = S.allreduce_stats(S.stats_vector(frames, n_samp))
    assert np.array_equal(got, ref), (rank, got, ref)
    dist.destroy_process_group()


def _run(kind, tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, kind, str(tmp_path)), nprocs=2, join=True)


def test_link_sharding_counters_agree(tmp_path):
    _run("links", tmp_path)


def test_overlapping_segments_dedup_to_the_sequential_frame_set(tmp_path):
    _run("stream", tmp_path)
    t = np.sort(np.concatenate([np.load(tmp_path / "trig0.npy"), np.load(tmp_path / "trig1.npy")]))
    assert np.array_equal(t, np.load(tmp_path / "trig_full.npy"))


def test_back_to_back_traffic_without_idle_gaps_still_reconciles(tmp_path):
    """Trigger chains that never merge (adv500) and a decode_mac tag pending across more than the overlap (adv641)."""
    for kind in ("adv500", "adv641", "tail500", "tail641"):
        _run(kind, tmp_path)
        t = np.sort(np.concatenate([np.load(tmp_path / "trig0.npy"), np.load(tmp_path / "trig1.npy")]))
        assert np.array_equal(t, np.load(tmp_path / "trig_full.npy"))


def test_reconcile_cascades_over_many_ranks():
    """Ranks 1..4 each start inside the adversarial zone: the lowest rank that has not joined decodes again, the next
    one is checked against the corrected table, and so on (all ranks in one process, a barrier where the all-gather is)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    from oracle import oracle as O
    from util import adversarial_stream, make_capture, oracle_segment_decoder
    S = importlib.import_module("gnuradio-wifi-imagetransfer_b200.sharding")
    for trunc in (500, 641):
        y = adversarial_stream(O, trunc=trunc)
        truth = S.records(O.rx(y, algo=1, want_carrier=False).frames, 0)
        for world in (3, 5):
            owned, rounds = S.simulate_ranks(oracle_segment_decoder(O, y, algo=1), S.shard_stream(y.size, world), y.size)
            assert S.same(np.concatenate(owned), truth) and rounds >= 1, (trunc, world, rounds)
            # the same with only the last 64 owned records of every rank exchanged (fewer than the zone's 230 frames:
            # a rank whose look-back shows no closed state decodes from the start of the capture)
            owned, rounds = S.simulate_ranks(oracle_segment_decoder(O, y, algo=1), S.shard_stream(y.size, world), y.size, tail_rows=64)
            assert S.same(np.concatenate(owned), truth) and rounds >= 1, (trunc, world, rounds, "tails")
    y, _ = make_capture(O, np.random.default_rng(8), [(4, 600)] * 30, snr_db=28, seed=3, gap=1100, cfo=0.004)
    owned, rounds = S.simulate_ranks(oracle_segment_decoder(O, y, algo=0), S.shard_stream(y.size, 4), y.size)
    assert S.same(np.concatenate(owned), S.records(O.rx(y, algo=0, want_carrier=False).frames, 0)) and rounds == 0


def test_reconcile_fuzz_random_traffic():
    """Random traffic -- gaps from none at all to a long silence, frames cut short by the next one, short captures against
    many ranks (cores shorter than the overlap) -- cut across 2..6 ranks: always the sequential table."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    from oracle import oracle as O
    from util import make_psdu, oracle_segment_decoder
    S = importlib.import_module("gnuradio-wifi-imagetransfer_b200.sharding")
    for seed in range(5):
        rng = np.random.default_rng(500 + seed)
        parts = [np.zeros(int(rng.integers(0, 400)), np.complex64)]
        for i in range(int(rng.integers(25, 60))):
            f = O.tx_frame(make_psdu(O, rng, int(rng.integers(30, 500)), seq=i), int(rng.integers(0, 8)), seed=1 + i % 127)
            if rng.random() < 0.2:
                f = f[:int(rng.integers(350, f.size))]                  # cut short: the next preamble re-triggers inside it
            parts += [f, np.zeros(int(rng.choice([0, 0, 40, 300, 900, 2500, 60000 if i % 17 == 5 else 700])), np.complex64)]
        x = np.concatenate(parts).astype(np.complex64)
        y = O.channel(x, gain=0.6, cfo=float(rng.uniform(-0.01, 0.01)), noise_sigma=0.6 * 10 ** (-float(rng.uniform(14, 30)) / 20), seed=seed)
        algo = int(rng.integers(0, 4))
        truth = S.records(O.rx(y, algo=algo, want_carrier=False).frames, 0)
        for world in (2, 3, 6):
            owned, rounds = S.simulate_ranks(oracle_segment_decoder(O, y, algo=algo), S.shard_stream(y.size, world), y.size)
            assert S.same(np.concatenate(owned), truth), (seed, world, rounds)
            owned, rounds = S.simulate_ranks(oracle_segment_decoder(O, y, algo=algo), S.shard_stream(y.size, world), y.size, tail_rows=16)
            assert S.same(np.concatenate(owned), truth), (seed, world, rounds, "tails")


def test_resumed_stream_state_reproduces_the_sequential_receiver():
    """(min_pos, fo_carry, hist) -- wifi_b200_link_state -- is all a decode needs to continue behind two regular frames."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    from oracle import oracle as O
    from util import make_capture
    S = importlib.import_module("gnuradio-wifi-imagetransfer_b200.sharding")
    y, _ = make_capture(O, np.random.default_rng(4), [(int(e), 200 + 50 * int(e)) for e in (0, 3, 5, 7, 2, 6, 1, 4)], snr_db=30, seed=6, gap=300, cfo=0.01)
    full = O.rx(y, algo=3, want_carrier=False)
    truth = S.records(full.frames, 0)
    for k in (2, 4, 6):
        lo, st, front = S.resume_point(truth, 0, int(truth[k, S.TRIG]) + 1)
        assert len(front) == 1 and S.same(front, truth[k - 1:k]) and lo + st["min_pos"] == truth[k, S.TRIG]
        r = O.rx(y[lo - st["hist"]:], algo=3, want_carrier=False, hist=st["hist"], min_pos=st["min_pos"], fo_carry=st["fo_carry"])
        assert S.same(S.records(r.frames, lo), truth[k:])
        assert r.pdus() == full.pdus()[k:]


def test_shard_geometry():
    sys.path.insert(0, ROOT)
    import importlib
    S = importlib.import_module("gnuradio-wifi-imagetransfer_b200.sharding")
    segs = S.shard_stream(10_000_000, 8)
    assert segs[0]["start"] == 0 and segs[-1]["end"] == 10_000_000
    assert all(s["start"] % 64 == 0 for s in segs)
    assert all(a["core_end"] == b["core_start"] for a, b in zip(segs, segs[1:]))
    assert all(s["core_start"] - s["start"] >= min(S.OVERLAP, s["core_start"]) - 64 for s in segs)
    assert sorted(sum((S.shard_links(11, 4, r) for r in range(4)), [])) == list(range(11))
