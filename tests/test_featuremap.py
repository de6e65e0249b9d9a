"""The feature-map packetiser around the PHY (featuremap.py) against golden digests made from the reference's own
image_detach_rebuild.py (tests/golden/make_featuremap_fixture.py), and its wire format."""
import hashlib
import json
import os
import pickle
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _digest(pieces):
    h = hashlib.sha256()
    for pos, piece in sorted(pieces, key=lambda p: p[0]):
        h.update(repr(tuple(int(v) for v in pos)).encode())
        h.update(np.ascontiguousarray(piece).tobytes())
        h.update(repr(piece.shape).encode())
    return h.hexdigest()


def test_detach_matches_the_references_packetiser(W):
    gold = json.load(open(os.path.join(HERE, "golden", "featuremap_fixture.json")))
    fm = W.featuremap
    for name, g in gold.items():
        rng = np.random.default_rng(g["seed"])
        dt = np.dtype(g["dtype"])
        a = (rng.standard_normal(g["shape"]) if dt == np.float32 else rng.integers(0, 256, g["shape"])).astype(dt)
        pieces = fm.detach(a, random_state=3)
        assert len(pieces) == g["n_pieces"] and _digest(pieces) == g["sorted_pieces_sha256"], name
        assert [p[0] for p in pieces] != sorted(p[0] for p in pieces)                         # shuffled, as the sender does
        first = sorted(pieces, key=lambda p: p[0])[0]
        d = fm.to_datagram(first)
        assert len(d) == g["datagram_bytes_of_first_sorted_piece"] and struct.unpack("=L", d[:4])[0] == len(d) - 4
        pos, piece = fm.from_datagram(d[4:])
        assert pos == first[0] and np.array_equal(piece, first[1])
        assert np.array_equal(fm.rebuild(pieces, a.shape, dt), a)
        assert np.count_nonzero(fm.rebuild(pieces[: len(pieces) // 2], a.shape, dt)) < np.count_nonzero(a)   # lost pieces stay zero
