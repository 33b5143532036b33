#!/bin/bash
# GPU call: wave structure of the sliding-window operator — one exact wave, two waves, a wave of half-units, a wave of quarter-units
O=gpurun_out/r04n; mkdir -p $O
for a in "2 74 80" "4 74 80" "1 74 80" "1 37 80" "3 69 80" "6 74 80"; do
  timeout 200 python tools/gn_operator_time.py $a 30 54 5 3 2>&1 | tail -1
done | tee $O/waves.txt
