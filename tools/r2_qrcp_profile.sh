#!/bin/bash
# per-step behaviour of the wide QRCP: step scaling, launch list of one factorisation, one full capture of a mid-run apply
mkdir -p gpurun_out
timeout 120 python tools/r2_qrcp_scaling.py > gpurun_out/qrcp_scaling.txt 2>&1
QR_STEPS=1024 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:qrw_ --csv --log-file gpurun_out/qrcp_launches.csv python tools/r2_qrcp_scaling.py > gpurun_out/qrcp_ncu.log 2>&1
QR_STEPS=256 timeout 200 ncu --set full --clock-control none --import-source on -k regex:qrw_apply --launch-skip 130 --launch-count 1 -o gpurun_out/qrcp_apply_full -f python tools/r2_qrcp_scaling.py > gpurun_out/qrcp_ncu2.log 2>&1
cat gpurun_out/qrcp_scaling.txt
tail -3 gpurun_out/qrcp_ncu.log
