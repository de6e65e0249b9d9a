#!/bin/bash
# Everything profiles/ holds for one round, taken on the GPU box in the order the profiling recipe asks for:
# plain run first (must exit 0), then the ncu launch list of the same command, the pipe-count metrics pass and one
# --set full capture of the large kernels.   usage: tools/profile_round.sh <tag>      (outputs under gpurun_out/)
set -u
cd "$(dirname "$0")/.."
TAG=${1:-rXX}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-time-shard"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_ncu_launches.csv \
    $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --metrics smsp__inst_executed.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_lsu.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active \
    --clock-control none -k 'regex:k_viterbi|k_demod|k_detect|k_sync_long' --csv --log-file gpurun_out/${TAG}_ncu_pipes.csv \
    $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
# last step's launches of the four large kernels (+ both demod phases)
ncu --set full --clock-control none --import-source on -k 'regex:k_viterbi|k_demod|k_detect|k_sync_long' -s 10 -c 5 -f \
    -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out/${TAG}_*
# the middle-sized-call Viterbi (one trellis per four lanes): 4736 frames in one call
CMDQ="python bench.py --links 32 --frames-per-link 148 --steps 2 --warmup 1 --no-e2e --no-cpu --no-time-shard"
$CMDQ > gpurun_out/${TAG}_quad_plain.json 2> gpurun_out/${TAG}_quad_plain.err && \
ncu --set full --clock-control none --import-source on -k 'regex:k_viterbi_quad' -c 1 -s 2 -f \
    -o gpurun_out/${TAG}_full_quad $CMDQ > gpurun_out/${TAG}_ncu4.log 2>&1
ls -la gpurun_out/${TAG}_*
