"""Writes tests/golden/kodim01_payload.bin: the bytes of the reference's images/kodim01.png, the payload of BASELINE.json
configs[0] ("kodim01.png packetized into 1500-byte PSDUs"), and its digest into kodim01_payload.json.  The GPU box has no
/root/reference, so the payload travels as a fixture.   python tests/golden/make_kodim01_fixture.py"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/images/kodim01.png"


def main():
    dst = os.path.join(HERE, "kodim01_payload.bin")
    shutil.copyfile(SRC, dst)
    os.chmod(dst, 0o644)
    data = open(dst, "rb").read()
    meta = {"source": "images/kodim01.png of OedonLestrange42/GNURadio-WiFI-ImageTransfer", "bytes": len(data),
            "sha256": hashlib.sha256(data).hexdigest(), "payload_bytes_per_frame": 1472, "psdu_bytes": 1500,
            "frames": -(-len(data) // 1472)}
    json.dump(meta, open(os.path.join(HERE, "kodim01_payload.json"), "w"), indent=1)
    print(meta)


if __name__ == "__main__":
    main()
