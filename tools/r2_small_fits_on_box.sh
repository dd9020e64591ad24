#!/bin/bash
# GPU test suite, then the full 50-iteration fits of cfg 1-3 on one GPU (tools/r2_fit.py).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; tail -3 gpurun_out/r2_gputests.log
timeout 300 python tools/r2_fit.py --workload ap --iterations 50 --sample-docs 2246 --sample-at 1,2,10,25,50 2>&1 | grep -E "SUMMARY|rror" | cut -c1-700
timeout 300 python tools/r2_fit.py --workload 20news --iterations 50 --symmetric --heldout 0.1 --tag _symmetric --sample-at 1,2,10,25,50 2>&1 | grep -E "SUMMARY|HELDOUT|rror" | cut -c1-700
timeout 300 python tools/r2_fit.py --workload 20news --iterations 50 --heldout 0.1 --tag _vector --sample-at 1,2,10,25,50 2>&1 | grep -E "SUMMARY|HELDOUT|rror" | cut -c1-700
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
