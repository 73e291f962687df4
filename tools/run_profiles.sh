#!/bin/bash
# ncu launch list + full-set capture of one seg step (tools/one_step.py) -> gpurun_out/<tag>_launches.csv, <tag>_full_raw.csv
# (the .ncu-rep stays in /tmp on the box: with --import-source it exceeds what gpurun copies back)
tag=${1:-r2k}
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
    python tools/one_step.py > gpurun_out/${tag}_ncu_l.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k 'regex:(linear_tma|mlp2|n2p_attend|edge_mlp_tc|knn_select|knn_tc|xgemm|knn_xyz2|ds_attend|ds_edge_partial|interpolate3_kernel)' \
    -o /tmp/${tag}_full python tools/one_step.py > gpurun_out/${tag}_ncu_f.log 2>&1
ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>> gpurun_out/${tag}_ncu_f.log
ls -la gpurun_out/${tag}_*
python tools/summarize_ncu.py launches gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches.md; head -30 gpurun_out/${tag}_launches.md
