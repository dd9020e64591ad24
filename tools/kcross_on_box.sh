#!/bin/bash
# Which kernel family wins at which topic count: warp-group (estep3 / estep3c) against the
# register-stationary kernel (MRLDA_ESTEP_KERNEL=2), Wikipedia-shape documents, V = 20000.
set -u
cd "$(dirname "$0")/.."
for K in 20 64 100; do
  for m in 105 1; do
    echo "== K=$K MRLDA_ESTEP3_MIN_TOPICS=$m"
    MRLDA_ESTEP3_MIN_TOPICS=$m timeout 100 python tools/probe.py --topics $K --terms 20000 --docs 40000 --steps 2 2>&1 | tail -1 | cut -c60-200
  done
done
for K in 112 128 160 224 256 320 400; do
  for k in 3 2; do
    echo "== K=$K MRLDA_ESTEP_KERNEL=$k"
    MRLDA_ESTEP_KERNEL=$k timeout 100 python tools/probe.py --topics $K --terms 20000 --docs 40000 --steps 2 2>&1 | tail -1 | cut -c60-200
  done
done
