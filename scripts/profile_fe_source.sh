#!/bin/bash
# Source-level stall profile of the fused front-end kernel (one launch of scripts/fe_attrib.py's workload).
AMIRA_FE_DEBUG=0 ncu --set full --clock-control none --import-source on -k regex:fe_fused -s 2 -c 1 -o gpurun_out/prof_fe_src python scripts/fe_attrib.py > gpurun_out/fe_src_ncu.log 2>&1
ncu -i gpurun_out/prof_fe_src.ncu-rep --page raw --csv > gpurun_out/fe_src_raw.csv
ncu -i gpurun_out/prof_fe_src.ncu-rep --page source --csv > gpurun_out/fe_src_source.csv 2>/dev/null
ls -la gpurun_out/fe_src*
