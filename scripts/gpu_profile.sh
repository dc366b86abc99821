#!/bin/bash
# ncu evidence for profiles/: (1) launch list with device times of two decoder steps, (2) full-set capture of the
# dominant kernel (conv_rows_tc, 12 launches of the second step), (3) full-set capture of the HBM-bound kernels of the
# IDWT chain and the layout moves (head_idwt x4, nchw_to_rows x3, gather_rows_list x2 of the second step).  Run under gpurun (1 GPU).
# Outputs -> gpurun_out/; scripts/summarise_profiles.py <tag> turns them into the tracked summaries.
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python scripts/profile_step.py 2 > gpurun_out/launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-conv_rows} -s ${KSKIP:-12} -c ${KCOUNT:-12} \
    -f -o gpurun_out/prof_conv python scripts/profile_step.py 2 > gpurun_out/prof_conv.log 2>&1; echo "full capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'head_idwt|nchw_to_rows|gather_rows_list' -s 9 -c 9 \
    -f -o gpurun_out/prof_other python scripts/profile_step.py 2 > gpurun_out/prof_other.log 2>&1; echo "second capture rc=$?"
ls -la gpurun_out | head -30
