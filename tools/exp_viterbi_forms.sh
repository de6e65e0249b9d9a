#!/bin/bash
# Viterbi latency by kernel form and frame count (one call, 64-QAM 3/4, 1528-byte frames): which form serves which call size.
# usage (on a B200): bash tools/exp_viterbi_forms.sh > gpurun_out/viterbi_forms.txt
for frames in 592 1184 2368 4736 9472 18944 37888; do
  links=$(( frames / 148 )); [ $links -lt 1 ] && links=1
  fpl=$(( frames / links ))
  for form in 1 2 3; do
    if [ $form -eq 1 ] && [ $frames -gt 9472 ]; then continue; fi
    python bench.py --links $links --frames-per-link $fpl --viterbi-form $form --steps 6 --warmup 3 --no-e2e --no-cpu --no-time-shard 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('frames %6d form %d viterbi %.3f ms  step %.3f ms crc_ok %s' % ($frames, $form, d['stage_ms']['viterbi'], d['ms_per_step'], d.get('crc_ok_per_step')))"
  done
done
