#!/bin/bash
# Everything profiles/ holds for one round, taken on the GPU box in the order the profiling recipe asks for:
# plain bench first (numbers), then the ncu launch list of the same command, then one --set full capture.
# usage: tools/profile_round.sh <tag>      (outputs under gpurun_out/)
set -u
cd "$(dirname "$0")/.."
TAG=${1:-rXX}
mkdir -p gpurun_out
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || { tail -5 gpurun_out/${TAG}_bench.err; exit 1; }
python bench.py --impl reference > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_ref.err
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_ncu_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu1.log 2>&1
# second step's launches of the four large kernels (+ both demod phases): skip the warm-up step's five
ncu --set full --clock-control none --import-source on -k 'regex:k_viterbi|k_demod|k_detect|k_sync_long' -s 5 -c 5 -f \
    -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu2.log 2>&1
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'stage', d['stage_ms'])"
