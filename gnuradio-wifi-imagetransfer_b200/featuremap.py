"""Feature-map mode around the PHY: what the reference's sender and viewer do on either side of the
flowgraph, without the neural codec (its weights are absent from the reference; any (H, W, C) array stands for the latent).

  upload_featuremap_udp.py:32-48     latent -> image_detach_rebuild.detach_image -> shuffled ((y, x, c), piece) tuples ->
                                     `=L` length + pickle, one datagram per piece, to UDP 50010
  download_featuremap_udp.py:53-69   datagram -> strip the length -> pickle.loads -> rebuild_image

`detach` / `rebuild` restate image_detach_rebuild.py:6-55 (10 x 10 x 1 pieces walked y, x, c; sklearn's shuffle, seedable
here so that tests repeat).  Pure host code: the payload format (pickle) is the reference's wire contract."""
import pickle
import struct

import numpy as np

PIECE_SIZE = (10, 10)


def detach(latent, piece_size=PIECE_SIZE, random_state=None):
    """List of ((y, x, c), piece[h, w, 1]) covering `latent`, shuffled (image_detach_rebuild.detach_image)."""
    from sklearn.utils import shuffle
    h, w, ch = latent.shape
    pieces = []
    for y in range(0, h, piece_size[1]):
        for x in range(0, w, piece_size[0]):
            for c in range(ch):
                pieces.append(((y, x, c), latent[y:y + piece_size[1], x:x + piece_size[0], c:c + 1]))
    return shuffle(pieces, random_state=random_state)


def to_datagram(piece):
    """One piece as the sender puts it on the wire: native-endian 32-bit length, then the pickle."""
    d = pickle.dumps(piece)
    return struct.pack("=L", len(d)) + d


def from_datagram(data):
    """What the viewer does with a received datagram (the flowgraph's "Extract Pics" block already removed the MAC header
    and the 4-byte length: pass `data[4:]` of the original datagram, i.e. the PDU's [24:][4:] slice)."""
    return pickle.loads(data)


def rebuild(pieces, shape, dtype=np.float32, piece_size=PIECE_SIZE):
    """image_detach_rebuild.rebuild_image for any dtype; positions never received stay zero."""
    out = np.zeros(shape, dtype)
    for (y, x, c), piece in pieces:
        out[y:y + piece_size[1], x:x + piece_size[0], c:c + 1] = piece
    return out
