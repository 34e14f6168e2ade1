#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tools/r2_setup_trace.py > gpurun_out/setup_trace.txt 2>&1
REPS=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:unfold_tiled --launch-count 2 -o gpurun_out/unfold_full -f python tools/r2_setup_trace.py > gpurun_out/unfold_ncu.log 2>&1
grep -v "^\[itcpd setup\] \(csr\|cand\|gather\|pivot\)" gpurun_out/setup_trace.txt | tail -40
tail -2 gpurun_out/unfold_ncu.log
