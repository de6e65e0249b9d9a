#!/bin/bash
# compute-sanitizer memcheck over a handful of small GPU tests (bulk-copy front-end, link groups, time sharding state,
# asynchronous pushes, gathered collections): one tool, smallest cases, as the profiling guide asks.
# usage (GPU box): tools/sanitize.sh   -> gpurun_out/r02_sanitize.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL='test_flags_match_frontend or test_rx_truncated_and_back_to_back or test_rx_multi_link or test_collection_through_more_than_four_bursts or test_resumed_stream_state_entry_point or test_host_batch_link_groups or test_asynchronous_pushes or test_rx_streaming_equals_batch or test_tx_bit_exact'
python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" > gpurun_out/r02_sanitize_plain.log 2>&1 || { tail -5 gpurun_out/r02_sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r02_sanitize_memcheck.txt \
    python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" > gpurun_out/r02_sanitize.log 2>&1
echo "exit $?" >> gpurun_out/r02_sanitize.log
tail -3 gpurun_out/r02_sanitize.log; tail -5 gpurun_out/r02_sanitize_memcheck.txt
