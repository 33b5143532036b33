#!/bin/bash
O=gpurun_out/r03l; mkdir -p $O
timeout 600 python tools/init_hostprof.py 2 > $O/hostprof2.txt 2>&1; grep -A 40 "cumulative" $O/hostprof2.txt | cut -c1-150 | head -45
