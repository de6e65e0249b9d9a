#!/bin/bash
# A/B harness: run bench.py once per prebuilt library variant under tools/variants/ (GPU box).
# usage: tools/ab.sh [bench args...]   -> gpurun_out/ab_<variant>.json
cd "$(dirname "$0")/.."
PKG=gnuradio-wifi-imagetransfer_b200
python $PKG/build.py >/dev/null
cp $PKG/libwifi_b200.so /tmp/orig.so
mkdir -p gpurun_out
for so in tools/variants/*.so; do
  v=$(basename $so .so)
  cp $so $PKG/libwifi_b200.so
  python bench.py --no-e2e --no-cpu "$@" > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.load(open("gpurun_out/ab_%s.json" % v))
    st = d["stage_ms"]
    print("%-12s value %9.1f  crc_ok %d/%d  " % (v, d["value"], d["crc_ok_per_step"], d["frames_per_step"]) + " ".join("%s %.3f" % (k, st[k]) for k in ("detect", "sync_long", "demod_data", "viterbi")))
except Exception as e:
    print(v, "FAILED", e)
PY
done
cp /tmp/orig.so $PKG/libwifi_b200.so
