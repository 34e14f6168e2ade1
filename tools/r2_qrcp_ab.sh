#!/bin/bash
# wide QRCP: register-resident apply kernel (two CTAs/SM with small spills, or one CTA/SM without) against the two-pass kernel
mkdir -p gpurun_out
{
timeout 200 python -m pytest tests/test_gpu_sampled.py -m gpu -x -q -k "qrcp or seqrcs" 2>&1 | tail -3
echo "== register-resident (default)"; QR_CHECK=1 timeout 100 python tools/r2_qrcp_scaling.py 2>&1 | tail -7
echo "== two-pass kernel"; ITCPD_QRCP_TWO_PASS=1 timeout 100 python tools/r2_qrcp_scaling.py 2>&1 | tail -6
} > gpurun_out/qrcp_ab.txt 2>&1
cat gpurun_out/qrcp_ab.txt
