"""Multi-GPU host logic: the path shards into independent units, so there is no
data-path collective.  Links are dealt round-robin to ranks; one long capture is cut into
overlapping segments, each frame is owned by the segment whose core region holds its trigger.
What crosses ranks is small: one all-reduce of a counter vector and, for time sharding, an
all-gather of the per-segment frame records with which every rank checks that its segment joined
the sequential receiver's state before its core region began (`reconcile`); a rank whose check
fails decodes again from a point where that state is known (NCCL over NVLink on GPUs, gloo in the
CPU tests)."""
import bisect
import time

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


# ---------------------------------------------------------------------------------------------------------------
# Time sharding, exactness.  A segment decoded from `start` (OVERLAP before its core) has the sequential receiver's
# frames in its core region iff its state at `core_start` is the sequential one: sync_short's last trigger (MIN_GAP
# rule), sync_long's carried frequency offset, decode_mac's collection.  With idle air between frames all three join
# within a frame or two; back-to-back traffic without a single idle gap can keep two trigger chains apart for ever.
# `boundary_check` decides from the frame records alone whether the segment joined; `resume_point` gives the place
# and the state from which to decode again when it did not.
# A frame record is the library's own 96-byte frame record (wifi_b200_frame / orc_frame, include/wifi_b200.h) seen as 24
# int32 words -- no conversion, the all-gather moves the table as it is -- with the trigger made absolute in place (a
# capture is one link: fewer than 2^31 samples, so the low word holds it).  Word of each field the checks look at:
TRIG, BURST, FREQ, FOUND, SIG, ENC, LEN, FSYM, NROWS, ACC, DEC, CRC = 0, 3, 5, 6, 9, 10, 11, 12, 13, 14, 15, 16
_CMP = [TRIG, BURST, FREQ, FOUND, SIG, ENC, LEN, FSYM, NROWS, ACC, DEC, CRC]      # what must agree between two decodes of the same frame
REC_FIELDS = 24                                                                    # (row / PSDU offsets, words 20-23, are per decode)
SS_MIN_GAP = 480
HIST = 128          # front-end history a resumed segment is given (two FE_CHUNKs: the running sums re-seed inside it)


def records(frames, offset):
    """Frame table -> int32 [n, 24] view of its records, ordered by absolute trigger position (changes `frames` in place)."""
    frames = np.ascontiguousarray(frames)
    assert frames.dtype.itemsize == 96
    r = frames.view(np.int32).reshape(len(frames), REC_FIELDS)
    if offset:
        r[:, TRIG] += offset
    if len(r) > 1 and np.any(np.diff(r[:, TRIG]) < 0):
        r = r[np.argsort(r[:, TRIG], kind="stable")]
    return r


def same(a, b):
    """Two record tables describe the same frames (every field that does not depend on where a decode started)."""
    return len(a) == len(b) and bool(np.array_equal(a[:, _CMP], b[:, _CMP]))


def _regular(r):
    """decode_mac's state is closed behind such a frame: own tag accepted, all of its symbols delivered by its own burst."""
    return (r[:, SIG] == 1) & (r[:, DEC] == 1) & (r[:, ACC] == 1) & (r[:, NROWS] >= r[:, FSYM]) & (r[:, FSYM] > 0)


def open_state(rec):
    """decode_mac's state behind the last record (the machine of decode_mac.cc / k_plan, replayed from the last closed
    point): trigger of the frame whose tag is still pending or whose symbol collection is still open, or None."""
    reg = _regular(rec)
    idx = np.nonzero(reg[:-1] & reg[1:])[0] if len(rec) >= 2 else np.zeros(0, int)
    start = int(idx[-1]) + 2 if len(idx) else 0
    cur, copied, need, pending = -1, 0, 0, -1
    for i in range(start, len(rec)):
        r = rec[i]
        if r[SIG] != 1:
            continue
        if r[NROWS] == 0:
            if pending < 0:
                pending = i
            continue
        tag = pending if pending >= 0 else i
        pending = -1
        if rec[tag, FSYM] <= 511 and rec[tag, LEN] <= 1528:
            cur, copied, need = tag, 0, int(rec[tag, FSYM])
        if cur < 0 or copied >= need:
            continue
        copied += min(int(r[NROWS]), need - copied)
    if pending >= 0:
        return int(rec[pending, TRIG])
    if cur >= 0 and copied < need:
        return int(rec[cur, TRIG])
    return None


def _f32_bits(x):
    return int(np.array([x], np.float32).view(np.int32)[0])


def _bits_f32(b):
    return float(np.array([b], np.int32).view(np.float32)[0])


def boundary_check(truth, local, carry_bits, seg_start, core_start):
    """truth: records of the sequential receiver's frames with trigger < core_start (what the lower ranks own);
    local: this segment's records; carry_bits: the sync_long frequency offset the segment's decode started with.
    True iff the segment's receiver is in the sequential state at core_start:
      * the two tables end (before core_start) with the same frames and that common tail holds two consecutive regular
        frames (behind the first no tag is pending, behind the second the collection is closed), or
      * the air was silent over the whole look-back, decode_mac was closed before it and the segment started with the
        frequency offset the sequential receiver carried into it."""
    t = truth[truth[:, TRIG] < core_start]
    l = local[local[:, TRIG] < core_start]
    if len(l) == 0 and (len(t) == 0 or t[-1, TRIG] < seg_start):
        if len(t) == 0:
            return carry_bits == _f32_bits(0.0)
        last_sig = t[t[:, SIG] == 1]
        closed = len(last_sig) == 0 or bool(_regular(last_sig[-1:])[0])
        return closed and carry_bits == int(t[-1, FREQ])
    m = min(len(t), len(l))
    if m < 2:
        return False
    same = np.all(t[len(t) - m:][:, _CMP] == l[len(l) - m:][:, _CMP], axis=1)[::-1]       # common tail, newest frame first
    n = m if same.all() else int(np.argmin(same))
    if n < 2:
        return False
    reg = _regular(t[len(t) - n:])
    return bool(np.any(reg[:-1] & reg[1:]))


def resume_point(truth, seg_start, core_start):
    """Where to decode again from, and with which state.  Returns (sample index `lo` the decode starts at -- its history
    lies in front of it --, state fields, records to put in front of the new table) or None.
      * silent look-back: the same start, carrying the sequential receiver's frequency offset;
      * else the last two consecutive regular frames (F1, F2) of the sequential table before core_start: start at
        F2's chunk, first possible trigger F2 itself, frequency offset as left by F1."""
    t = truth[truth[:, TRIG] < core_start]
    if len(t) and t[-1, TRIG] < seg_start:
        last_sig = t[t[:, SIG] == 1]
        if len(last_sig) == 0 or bool(_regular(last_sig[-1:])[0]):
            return seg_start, {"min_pos": max(0, int(t[-1, TRIG]) + SS_MIN_GAP + 1 - seg_start), "fo_carry": _bits_f32(t[-1, FREQ]),
                               "hist": min(HIST, seg_start)}, t[:0]
    reg = _regular(t)
    idx = np.nonzero(reg[:-1] & reg[1:])[0] if len(t) >= 2 else np.zeros(0, int)
    if len(idx) == 0:
        return None
    f1, f2 = t[idx[-1]], t[idx[-1] + 1]
    start = int(f2[TRIG]) // FE_CHUNK * FE_CHUNK
    return start, {"min_pos": int(f2[TRIG]) - start, "fo_carry": _bits_f32(f1[FREQ]), "hist": min(HIST, start)}, t[idx[-1]:idx[-1] + 1]


def gather_records(hdr, rec, device=None, fixed_rows=None):
    """All ranks' (header, record table) pairs (NCCL all_gather of the padded int32 tables on GPUs, gloo on CPU); list
    indexed by rank.  The header row says where a rank's decode started and the frequency offset it started with.
    fixed_rows = K: every rank sends at most K rows, so the buffers have a known size and the row count travels in the
    header (column FOUND) -- ONE collective and one copy back instead of a count exchange first."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [(hdr, rec)]
    world = dist.get_world_size()
    rec = np.ascontiguousarray(rec, np.int32)
    if fixed_rows is not None:
        assert len(rec) <= fixed_rows
        mine = np.zeros((fixed_rows + 1, REC_FIELDS), np.int32)
        mine[0] = hdr
        mine[0, FOUND] = len(rec)
        mine[1:1 + len(rec)] = rec
        buf = torch.from_numpy(mine).to(device) if device is not None else torch.from_numpy(mine)
        everyone = torch.empty((world * buf.shape[0], buf.shape[1]), dtype=torch.int32, device=buf.device)    # rank r's rows at [r * (K + 1), ...): the layout gloo accepts too
        dist.all_gather_into_tensor(everyone, buf)
        a = everyone.cpu().numpy().reshape(world, buf.shape[0], buf.shape[1])
        out = []
        for r in range(world):
            c = int(a[r, 0, FOUND])
            h = a[r, 0].copy()
            h[FOUND] = 0
            out.append((h, a[r, 1:1 + c]))
        return out
    n = torch.tensor([len(rec)], dtype=torch.int64, device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    cap = max(counts) + 1
    buf = torch.zeros((cap, REC_FIELDS), dtype=torch.int32, device=device)
    buf[0] = torch.from_numpy(hdr).to(buf.device)
    if len(rec):
        buf[1:1 + len(rec)] = torch.from_numpy(rec).to(buf.device)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    out = []
    for b, c in zip(bufs, counts):
        a = b[:1 + c].cpu().numpy()
        out.append((a[0], a[1:]))
    return out


def _header(lo, carry_bits):
    h = np.zeros(REC_FIELDS, np.int32)
    h[TRIG], h[BURST], h[FREQ] = -1, lo, carry_bits
    return h


def _owned(rec, seg):
    """The rows of a trigger-ordered table that lie in the segment's core region (a view)."""
    col = rec[:, TRIG]                    # strided view: bisect touches ~2 log2(n) elements, np.searchsorted would copy the column
    return rec[bisect.bisect_left(col, seg["core_start"]):bisect.bisect_left(col, seg["core_end"])]


def reconcile(decode, segs, rank, n_samples, device=None, max_rounds=None, gather=None, tail_rows=None, trace=None):
    """Time-sharded receive with an exact result.  `decode(lo, end, state, final)` runs this rank's receiver over
    samples [lo, end) of the capture -- state None: a stream start at lo; else the wifi_b200_link_state fields, with
    state["hist"] samples of history in front of lo (the callee reads them from lo - hist) -- and returns its frame
    table (triggers relative to lo).
    Every rank decodes its segment; the records are all-gathered; every rank walks the same check from rank 0 upwards;
    the lowest rank that has not joined the sequential state decodes again from a known one; repeated until all have.
    Returns (records this rank owns, every rank's owned records, number of re-decoding rounds).

    tail_rows = K: the ranks exchange only the last K records they own (a few tens of kB instead of the whole tables) and
    every rank checks ITSELF against the tails of the ranks below it, then the verdicts are gathered; what comes back as
    "every rank's owned records" are those tails.  Ordinary traffic joins within a frame or two, so K = 512 decides every
    check; a look-back with more than K irregular frames falls back to decoding from the start of the capture."""
    seg = segs[rank]
    do_gather = gather or gather_records

    def lap(key, t0):
        # trace: a dict that receives the wall time (s) this rank spent per phase (decode, exchange, check); tail mode only
        if trace is not None:
            trace[key] = trace.get(key, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    def decode_closed(lo, st):
        """Decode [lo, end); while a frame this rank owns leaves decode_mac open at the end of the segment (a tag pending,
        a collection short of symbols: state the sequential receiver keeps for as long as it takes), decode further."""
        end, grow = seg["end"], OVERLAP
        while True:
            rec = records(decode(lo, end, st, end == n_samples), lo)
            opener = None if end == n_samples else open_state(rec)
            if opener is None or opener >= seg["core_end"]:
                return rec
            end, grow = min(n_samples, end + grow), 2 * grow

    lo, carry = seg["start"], _f32_bits(0.0)
    t_ = time.perf_counter()
    local = decode_closed(lo, None)
    t_ = lap("decode", t_)
    rounds = 0
    while tail_rows is not None:
        own = _owned(local, seg)
        tails = do_gather(_header(lo, carry), own[-tail_rows:], device, tail_rows)
        t_ = lap("exchange", t_)
        truth = np.concatenate([t for _, t in tails[:rank]]) if rank else np.zeros((0, REC_FIELDS), np.int32)
        truth = truth[truth[:, TRIG] < seg["core_start"]]
        ok = lo == 0 or boundary_check(truth, local, carry, lo, seg["core_start"])     # a decode from sample 0 IS the sequential receiver
        t_ = lap("check", t_)
        verdicts = do_gather(_header(int(ok), 0), np.zeros((0, REC_FIELDS), np.int32), device, 0)
        t_ = lap("exchange", t_)
        bad = next((r for r, (h, _) in enumerate(verdicts) if int(h[BURST]) == 0), None)
        if bad is None:                      # rank 0 is right; rank r agreed with the tails of ranks it just saw agree: all are
            return own, [t for _, t in tails], rounds
        rounds += 1
        if max_rounds is not None and rounds > max_rounds:
            raise RuntimeError("time sharding did not reconcile in %d rounds" % max_rounds)
        if bad == rank:
            rp = resume_point(truth, seg["start"], seg["core_start"])
            if rp is None:                   # no closed state among the records at hand: only the start of the capture is known
                lo, carry = 0, _f32_bits(0.0)
                local = decode_closed(0, None)
            else:
                lo, st, front = rp
                carry = _f32_bits(st["fo_carry"])
                local = np.concatenate([front, decode_closed(lo, st)])
    while True:
        allrec = do_gather(_header(lo, carry), local, device)
        owned_all, bad = [], None

        def truth_before(core_start, rows):
            """The lower ranks' owned records in front of core_start: the last `rows` of them (None: all)."""
            parts, have = [], 0
            for own in reversed(owned_all):
                parts.append(own)
                have += len(own)
                if rows is not None and have >= rows:
                    break
            t = np.concatenate(parts[::-1]) if parts else np.zeros((0, REC_FIELDS), np.int32)
            return t[t[:, TRIG] < core_start]

        for r, sr in enumerate(segs):
            hdr, rec = allrec[r]
            if r > 0 and int(hdr[BURST]) != 0:           # (a decode that started at sample 0 is the sequential receiver)
                n_front = int(np.searchsorted(rec[:, TRIG], sr["core_start"]))
                if not boundary_check(truth_before(sr["core_start"], n_front + 64), rec, int(hdr[FREQ]), int(hdr[BURST]), sr["core_start"]):
                    bad = r
                    break
            owned_all.append(_owned(rec, sr))
        if bad is None:
            return owned_all[rank], owned_all, rounds
        rounds += 1
        if max_rounds is not None and rounds > max_rounds:
            raise RuntimeError("time sharding did not reconcile in %d rounds" % max_rounds)
        if bad == rank:
            rp = resume_point(truth_before(seg["core_start"], None), seg["start"], seg["core_start"])
            if rp is None:                                # nothing is known but the start of the capture
                lo, carry = 0, _f32_bits(0.0)
                local = decode_closed(0, None)
            else:
                lo, st, front = rp
                carry = _f32_bits(st["fo_carry"])
                local = np.concatenate([front, decode_closed(lo, st)])


def simulate_ranks(decode, segs, n_samples, tail_rows=None):
    """All ranks of `reconcile` in one process (threads and a barrier where the all-gather is): what the single-GPU
    tests use.  Returns (every rank's owned records, rounds)."""
    import threading
    world = len(segs)
    slots, out, bar, errs = [None] * world, [None] * world, threading.Barrier(world), []

    def run(rank):
        def gather(hdr, rec, device=None, fixed_rows=None):
            slots[rank] = (hdr, rec)
            bar.wait()
            res = list(slots)
            bar.wait()
            return res
        try:
            out[rank] = reconcile(decode, segs, rank, n_samples, gather=gather, max_rounds=world + 1, tail_rows=tail_rows)
        except Exception as e:        # a failing rank must not leave the others at the barrier
            errs.append(e)
            bar.abort()

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errs:
        raise errs[0]
    if tail_rows is not None:            # every rank's own records (the exchanged tails are only what the checks needed)
        return [o[0] for o in out], out[0][2]
    return out[0][1], out[0][2]


def gather_owned(own, device=None):
    """Every rank's owned records (for a check of the union against a sequential decode; not part of the receive path)."""
    return [t for _, t in gather_records(_header(0, 0), own, device)]
