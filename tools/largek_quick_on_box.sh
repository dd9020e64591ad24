#!/bin/bash
# quick cfg-5 check of a kernel edit: parity (mixed lengths, slow path), three length slices, the mix
set -u
cd "$(dirname "$0")/.."
timeout 300 python tools/check3.py 1000 2>&1 | tail -4 | cut -c1-200
timeout 120 python tools/check_sparse.py 1000 2>&1 | tail -2 | cut -c1-200
for spec in "33 96 60000" "97 160 60000" "161 320 60000"; do
  read -r lo hi docs <<<"$spec"
  timeout 200 python tools/probe.py --workload largek --docs "$docs" --steps 2 --min-rows "$lo" --max-rows "$hi" 2>&1 | tail -1 | cut -c1-200
done
timeout 200 python tools/probe.py --workload largek --docs 100000 --steps 2 2>&1 | tail -1 | cut -c1-200
