#!/bin/bash
# ncu evidence for profiles/: (1) launch list with device times of two decoder steps, (2) full-set capture of the
# dominant kernel.  Run under gpurun (1 GPU).  Outputs -> gpurun_out/.
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python scripts/profile_step.py 2 > gpurun_out/launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-conv_rows} -s ${KSKIP:-12} -c ${KCOUNT:-12} \
    -f -o gpurun_out/prof_conv python scripts/profile_step.py 2 > gpurun_out/prof_conv.log 2>&1; echo "full capture rc=$?"
ls -la gpurun_out
