#!/bin/bash
# cfg-5 (K = 1000) kernels on one GPU box: parity on a mixed-length corpus and on the slow path, then
# the probe per document-length slice and on the whole mix, cluster warp-group kernel (default)
# against the register-stationary cluster kernel (MRLDA_ESTEP_KERNEL=2).
#   gpurun --timeout 900 -- tools/largek_on_box.sh
set -u
cd "$(dirname "$0")/.."
export MRLDA_DEBUG=${MRLDA_DEBUG:-0}
[ "$MRLDA_DEBUG" = 0 ] && unset MRLDA_DEBUG
timeout 300 python tools/check3.py 1000 2>&1 | tail -5 | cut -c1-200
timeout 120 python tools/check_sparse.py 1000 2>&1 | tail -2 | cut -c1-200
for spec in "33 96 60000" "97 160 60000" "161 320 60000"; do
  read -r lo hi docs <<<"$spec"
  for k in 3 2; do
    echo "== rows $lo..$hi, MRLDA_ESTEP_KERNEL=$k"
    MRLDA_ESTEP_KERNEL=$k timeout 200 python tools/probe.py --workload largek --docs "$docs" --steps 2 --min-rows "$lo" --max-rows "$hi" 2>&1 | tail -1 | cut -c1-200
  done
done
for k in 3 2; do
  echo "== all documents, MRLDA_ESTEP_KERNEL=$k"
  MRLDA_ESTEP_KERNEL=$k timeout 200 python tools/probe.py --workload largek --docs 100000 --steps 2 2>&1 | tail -1 | cut -c1-200
done
