#!/bin/bash
O=gpurun_out/r03x; mkdir -p $O
timeout 600 python tools/e2e_probe.py 2 > $O/probe.txt 2>&1; tail -9 $O/probe.txt
