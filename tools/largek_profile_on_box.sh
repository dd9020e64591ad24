#!/bin/bash
# cfg-5-shape profile on one GPU (a 100k-document slice of the large-K corpus): launch list of one
# E-step and an ncu --set full capture of the cluster warp-group kernel's launches.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/probe.py --workload largek --docs 100000 --steps 2 2>&1 | tail -1 | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
  --log-file gpurun_out/r2_largek_launches.csv \
  python tools/probe.py --workload largek --docs 100000 --steps 1 > gpurun_out/r2_largek_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:estep3c_kernel -s 4 -c 4 \
  -o gpurun_out/r2_estep3c_full -f \
  python tools/probe.py --workload largek --docs 100000 --steps 1 > gpurun_out/r2_largek_ncu_full.log 2>&1
ncu -i gpurun_out/r2_estep3c_full.ncu-rep --page raw --csv > gpurun_out/r2_estep3c_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_estep3c_full.ncu-rep --page details > gpurun_out/r2_estep3c_full_details.txt 2>/dev/null
ls -la gpurun_out | tail -8
