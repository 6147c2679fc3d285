#!/usr/bin/env bash
# Developer tool: one `ncu --set full` capture of every tail kernel (second forward of a short bench run) plus the launch list.
#   gpurun --timeout 900 -- 'bash tools/ncu_tail.sh'
set -u
mkdir -p gpurun_out
PAT='regex:nsm_power|knn_select|head_kernel|seed_hypotheses|nms_key|sc_matrix_tiled|knn_dist_tc|seed_sort|select_refine'
timeout 600 ncu --set full --clock-control none --import-source on -k "$PAT" -s 9 -c 9 -f -o gpurun_out/r2_tail \
    python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2_tail_ncu.log 2>&1
tail -3 gpurun_out/r2_tail_ncu.log
