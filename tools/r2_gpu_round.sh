#!/bin/bash
# One GPU-box call of the round: GPU test suite, bench (both arms), launch list, one ncu capture of
# the dominant E-step launch.  Everything lands in gpurun_out/<tag>_*; a number printed by a run
# under ncu is never a bench value.    usage: tools/r2_gpu_round.sh <tag> [tests|notests] [ncu|noncu]
set -u
cd "$(dirname "$0")/.."
tag=${1:-r2}
mkdir -p gpurun_out
if [ "${2:-tests}" = tests ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gputests.log 2>&1
  tail -3 gpurun_out/${tag}_gputests.log
fi
timeout 600 python bench.py > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err || exit 1
cut -c1-600 gpurun_out/${tag}_bench_1gpu.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
cut -c1-300 gpurun_out/${tag}_bench_reference.json
[ "${3:-ncu}" = ncu ] || exit 0
# launch list (cold-cache, serialised: shares of the step, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${tag}_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_launches.log 2>&1
# full capture of the longest E-step launch (the 4-warps-per-document class of estep3_kernel)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:estep3_kernel -s 4 -c 4 \
  -o gpurun_out/${tag}_estep3_full -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_full.log 2>&1
ncu -i gpurun_out/${tag}_estep3_full.ncu-rep --page raw --csv > gpurun_out/${tag}_estep3_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_estep3_full.ncu-rep --page details > gpurun_out/${tag}_estep3_full_details.txt 2>/dev/null
ls -la gpurun_out | tail -12
