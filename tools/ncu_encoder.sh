#!/usr/bin/env bash
# Developer tool: `ncu --set full` of the encoder kernels of the SECOND layer of the second forward (B=256, N=1000).
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:tc_chain|tc_attention' -s 52 -c 4 -f -o gpurun_out/r2_encoder \
    python tools/one_forward.py 256 1000 2 > gpurun_out/r2_encoder_ncu.log 2>&1
tail -2 gpurun_out/r2_encoder_ncu.log
