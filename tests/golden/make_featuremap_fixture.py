"""Golden digests of the reference's own packetiser: imports /root/reference/image_detach_rebuild.py (pure numpy +
sklearn), cuts two seeded arrays with detach_image and round-trips them through rebuild_image.  The reference shuffles
without a seed, so the digest is taken over the pieces in sorted order.   python tests/golden/make_featuremap_fixture.py"""
import hashlib
import json
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")


def digest(pieces):
    h = hashlib.sha256()
    for pos, piece in sorted(pieces, key=lambda p: p[0]):
        h.update(repr(tuple(int(v) for v in pos)).encode())
        h.update(np.ascontiguousarray(piece).tobytes())
        h.update(repr(piece.shape).encode())
    return h.hexdigest()


def main():
    import image_detach_rebuild as R
    out = {}
    for name, shape, dtype in (("latent_30x30x128_f32", (30, 30, 128), np.float32), ("image_300x300x3_u8", (300, 300, 3), np.uint8)):
        rng = np.random.default_rng(7)
        a = (rng.standard_normal(shape) if dtype == np.float32 else rng.integers(0, 256, shape)).astype(dtype)
        pieces = R.detach_image(a)
        out[name] = {"shape": list(shape), "dtype": np.dtype(dtype).name, "seed": 7, "n_pieces": len(pieces), "sorted_pieces_sha256": digest(pieces),
                     "datagram_bytes_of_first_sorted_piece": 4 + len(pickle.dumps(sorted(pieces, key=lambda p: p[0])[0]))}
        if dtype == np.uint8:
            assert np.array_equal(R.rebuild_image(pieces, shape), a)
    json.dump(out, open(os.path.join(HERE, "featuremap_fixture.json"), "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()
