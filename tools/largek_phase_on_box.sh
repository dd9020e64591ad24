#!/bin/bash
# Per-phase cycle counters of the cluster E-step kernel on cfg-5-shape slices (a
# -DMRLDA_PHASE_TIMING build shipped as scratch_so/phase.so).  Each line of SPECS:
# "<MRLDA_ESTEP2_CLASSES> <min rows> <max rows> <documents generated>"
set -u
cd "$(dirname "$0")/.."
cp mrlda_b200/libmrlda_b200.so /tmp/keep.so
cp scratch_so/phase.so mrlda_b200/libmrlda_b200.so || exit 1
export MRLDA_ESTEP_KERNEL=2
SPECS="${SPECS:-0 97 160 60000;1 97 160 60000}"
IFS=';' read -ra specs <<<"$SPECS"
for spec in "${specs[@]}"; do
  read -r cls lo hi docs <<<"$spec"
  echo "== classes=$cls rows $lo..$hi"
  MRLDA_ESTEP2_CLASSES=$cls timeout 200 python tools/probe.py --workload largek --docs "$docs" --steps 2 --min-rows "$lo" --max-rows "$hi" 2>&1 | tail -3 | cut -c1-260
done
cp /tmp/keep.so mrlda_b200/libmrlda_b200.so
