#!/usr/bin/env python
"""Writes a Wikipedia-shape corpus as a ParseCorpus-style SequenceFile directory and runs the native
mrlda-vi driver on it (fresh training, then resume), printing the log tail and wall times."""
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrlda_b200 import corpus, seqfile  # noqa: E402

docs = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c = corpus.generate(docs, 100000, 200, 200, seed=4, device="cuda")
root = tempfile.mkdtemp(prefix="mrlda_e2e_")
inp, out = os.path.join(root, "input"), os.path.join(root, "model") + "/"
os.makedirs(inp)
t0 = time.time()
n_parts = 4
b = corpus.shard_bounds(c.row_ptr, n_parts)
for p in range(n_parts):
    s = c.slice_docs(b[p], b[p + 1])
    seqfile.write_corpus(os.path.join(inp, f"part-{p:05d}"), s.row_ptr, s.term_id, s.count,
                         doc_id=list(range(b[p] + 1, b[p + 1] + 1)))
print(f"wrote {docs} documents ({c.nnz} nnz) in {time.time() - t0:.1f} s")
drv = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mrlda_b200", "host", "mrlda-vi")
for extra in ([], ["-modelindex", str(iters), "-iteration", str(iters + 2)]):
    t0 = time.time()
    args = [drv, "-input", inp, "-output", out, "-term", "100000", "-topic", "200"]
    args += extra if extra else ["-iteration", str(iters)]
    p = subprocess.run(args, capture_output=True, text=True)
    print(f"--- {' '.join(extra) or 'fresh'}: exit {p.returncode}, {time.time() - t0:.1f} s")
    print("\n".join(l for l in p.stdout.splitlines() if "Log likelihood after" in l or "E-step" in l))
    if p.returncode:
        print(p.stderr[-2000:])
print(sorted(os.listdir(out)))
subprocess.run(["rm", "-rf", root])
