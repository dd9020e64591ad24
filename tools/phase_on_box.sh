#!/bin/bash
# Per-phase cycle counters of the warp-group E-step kernel (a -DMRLDA_PHASE_TIMING build shipped as
# scratch_so/<name>.so), one document-length slice per line.  usage: tools/phase_on_box.sh <name>
set -u
cd "$(dirname "$0")/.."
cp mrlda_b200/libmrlda_b200.so /tmp/keep.so
cp "scratch_so/$1.so" mrlda_b200/libmrlda_b200.so || exit 1
for r in "40 72" "73 96" "97 128" "129 144" "145 160" "161 192" "193 224" "225 256" "257 288"; do
  read -r lo hi <<<"$r"
  echo "== rows $lo..$hi"
  timeout 200 python tools/probe.py --docs 40000 --steps 2 --min-rows "$lo" --max-rows "$hi" 2>&1 | tail -2 | cut -c1-260
done
cp /tmp/keep.so mrlda_b200/libmrlda_b200.so
