#!/bin/bash
set -u
cd "$(dirname "$0")/.."
export MRLDA_DEBUG=1
for spec in "0 0" "1 0" "1 1"; do
  read -r cls noclamp <<<"$spec"
  echo "== classes=$cls noclamp=$noclamp rows 97..160"
  if [ "$noclamp" = 1 ]; then export MRLDA_NO_CLUSTER_CLAMP=1; fi
  MRLDA_ESTEP2_CLASSES=$cls timeout 200 python tools/probe.py --workload largek --docs 60000 --steps 1 --min-rows 97 --max-rows 160 2>&1 | tail -2 | cut -c1-260
done
