"""Multi-GPU host logic: the path shards into independent units, so there is no
data-path collective.  Links are dealt round-robin to ranks; one long capture is cut into
overlapping segments, each frame is owned by the segment whose core region holds its trigger.
Only counters cross ranks (one all-reduce over NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np

# sync_short MAX_SAMPLES + sync_long window + FFT + MIN_GAP look-back + window warm-up
OVERLAP = 43200 + 320 + 64 + 480 + 64
FE_CHUNK = 64

STAT_KEYS = ("samples", "frames_detected", "signal_ok", "decoded", "crc_ok", "pdu_bytes")


def shard_links(n_links, world, rank):
    """Link ids owned by `rank` (round-robin)."""
    return list(range(rank, n_links, world))


def shard_stream(n_samples, world, overlap=OVERLAP):
    """Cut [0, n_samples) into `world` segments.  Returns a list of dicts with the samples to
    process [start, end) and the core region [core_start, core_end) whose triggers the segment owns.
    Starts are aligned to the front-end chunk so the running sums re-seed on the same grid."""
    core = -(-n_samples // world)
    core = -(-core // FE_CHUNK) * FE_CHUNK
    segs = []
    for r in range(world):
        cs, ce = min(r * core, n_samples), min((r + 1) * core, n_samples)
        start = max(0, (cs - overlap) // FE_CHUNK * FE_CHUNK)
        end = min(n_samples, ce + overlap)
        segs.append({"rank": r, "start": start, "end": end, "core_start": cs, "core_end": ce})
    return segs


def owned(frames, seg):
    """Mask of the frames (segment-relative triggers) this segment owns after dedup."""
    t = frames["trigger"].astype(np.int64) + seg["start"]
    return (t >= seg["core_start"]) & (t < seg["core_end"])


def stats_vector(frames, n_samples):
    ok = frames["crc_ok"] == 1
    v = [n_samples, len(frames), int(frames["sig_ok"].sum()), int(frames["decoded"].sum()), int(ok.sum()),
         int((frames["length"][ok] - 4).sum())]
    v += [int((ok & (frames["encoding"] == e)).sum()) for e in range(8)]
    return np.array(v, np.int64)


def allreduce_stats(vec, device=None):
    """Sum of the counter vector over all ranks (NCCL over NVLink on GPUs).  Returns numpy int64."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.asarray(vec, np.int64).copy())
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
