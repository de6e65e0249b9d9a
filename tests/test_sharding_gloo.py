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
        y, _ = make_capture(O, rng, [(2, 400)] * 40, snr_db=30, seed=9, gap=900)
        seg = S.shard_stream(y.size, world, overlap=S.OVERLAP)[rank]
        r = O.rx(y[seg["start"]:seg["end"]], algo=0, want_carrier=False, final=(seg["end"] == y.size))
        keep = S.owned(r.frames, seg)
        frames = r.frames[keep]
        n_samp = seg["core_end"] - seg["core_start"]
        full = O.rx(y, algo=0, want_carrier=False).frames
        ref = S.stats_vector(full, y.size)
        trig = np.sort(frames["trigger"].astype(np.int64) + seg["start"])
        np.save(os.path.join(out_dir, "trig%d.npy" % rank), trig)
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
