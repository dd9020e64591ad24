#!/bin/bash
# A/B of two (or more) builds of libmrlda_b200.so on one GPU box, same process environment.
#
# Here:   build each candidate, copy it to scratch_so/<name>.so (scratch_so/ is git-ignored and
#         travels with the snapshot), then
#         gpurun --timeout 600 -- 'tools/ab_on_box.sh good new'
# On the box this script puts each library in place in turn and runs the probe (K=200, 200k
# documents, ~290 ms per E-step, repeatable to about 0.2 %) plus three document-length slices.
# It finishes by running the GPU test suite against the last library named.
set -u
cd "$(dirname "$0")/.."
for name in "$@"; do
  cp "scratch_so/$name.so" mrlda_b200/libmrlda_b200.so || exit 1
  echo "== $name"
  timeout 200 python tools/check3.py 200 2>&1 | tail -4 | cut -c1-150
  for r in "33 64" "97 128" "193 224"; do
    read -r lo hi <<<"$r"
    timeout 200 python tools/probe.py --docs 40000 --steps 3 --min-rows "$lo" --max-rows "$hi" 2>&1 | tail -1 | cut -c1-150
  done
  timeout 200 python tools/probe.py --docs 200000 --steps 3 2>&1 | tail -1 | cut -c1-150
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
