#!/bin/bash
# A/B runs of one batch-4 window call under environment switches: tools/ab.sh "VAR=1" "VAR=2 OTHER=3" ...  (each twice, interleaved)
for rep in 1 2; do
  for cfg in "$@"; do
    printf "%-50s " "[$cfg]"
    env $cfg timeout 300 python tools/one_window.py --batch ${AB_BATCH:-4} --reps ${AB_REPS:-12} 2>&1 | tail -1 | cut -c1-90
  done
done
