#!/usr/bin/env python
"""Throughput of the SequenceFile / Document codecs the native driver uses (CPU only): writes and
reads a Wikipedia-shape document directory with and without stored gamma through
libmrlda_seqfile.so and prints MB/s.   usage: tools/io_bench.py [docs] [part files]"""
import os
import shutil
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrlda_b200 import corpus, seqfile  # noqa: E402

docs = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 1
K, V = 200, 100000
c = corpus.generate(docs, V, K, 200, seed=4)
g = np.random.default_rng(1).random((c.n_docs, K))
root = tempfile.mkdtemp(prefix="mrlda_io_", dir="/tmp")


def size_of(d):
    return sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))


def timed(label, fn, nbytes=None):
    t = time.perf_counter()
    out = fn()
    dt = time.perf_counter() - t
    nb = nbytes() if callable(nbytes) else nbytes
    print(f"{label:34s} {dt:6.2f} s  {nb / dt / 1e6:7.0f} MB/s  ({nb / 1e6:.0f} MB)")
    return out


try:
    for with_gamma in (False, True):
        d = os.path.join(root, "gamma" if with_gamma else "input")
        os.makedirs(d)
        b = corpus.shard_bounds(c.row_ptr, parts)

        def write():
            for p in range(parts):
                s = c.slice_docs(b[p], b[p + 1])
                seqfile.write_corpus(os.path.join(d, f"part-{p:05d}"), s.row_ptr, s.term_id, s.count,
                                     gamma=g[b[p]:b[p + 1]] if with_gamma else None, n_topics=K if with_gamma else 0)
        tag = "with gamma" if with_gamma else "no gamma"
        timed(f"write {parts} part(s), {tag}", write, lambda: size_of(d))
        r = timed(f"read  {parts} part(s), {tag}", lambda: seqfile.read_corpus(d, K, V), size_of(d))
        assert np.array_equal(r["row_ptr"], c.row_ptr) and np.array_equal(r["term_id"], c.term_id)
        assert np.array_equal(r["count"], c.count) and (not with_gamma or np.array_equal(r["gamma"], g))
        timed(f"read  shard 1 of 2, {tag}", lambda: seqfile.read_corpus(d, K, V, rank=1, world=2), size_of(d) / 2)
    # the driver's own gamma output path: concurrent part files of about 64 MB
    d = os.path.join(root, "gamma-parts")
    os.makedirs(d)
    n = timed("write_gamma_parts (driver path)",
              lambda: seqfile.write_gamma_parts(d, 0, c.row_ptr, c.term_id, c.count, g), lambda: size_of(d))
    print(f"  -> {n} part file(s), {os.cpu_count()} host threads visible")
    r = seqfile.read_corpus(d, K, V)
    assert np.array_equal(r["gamma"], g) and np.array_equal(r["term_id"], c.term_id)
    # the model file of one iteration: beta-N (K records of V values each)
    dl = np.random.default_rng(2).standard_normal((V, K))
    nm, pr = np.ones(K, np.float32), np.ones(V, np.uint8)
    bp = os.path.join(root, "beta-1")
    timed("write_beta", lambda: seqfile.write_beta(bp, dl, nm, pr), lambda: os.path.getsize(bp))
    rb = timed("read_beta", lambda: seqfile.read_beta(bp, K, V), os.path.getsize(bp))
    assert np.array_equal(rb[0], dl)
finally:
    shutil.rmtree(root)
