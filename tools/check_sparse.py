#!/usr/bin/env python
"""Development aid: the exact log-space slow path (rows whose linear-space norm underflows) against
the oracle for a given topic count.  usage: check_sparse.py [K]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrlda_b200 import capi, corpus  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    V, D = 400, 24
    rng = np.random.default_rng(5)
    lb = np.full((V, K), -1.0e12)
    owner = rng.integers(0, K, size=V)
    lb[np.arange(V), owner] = np.log(rng.random(V) + 0.1)
    docs = []
    for d in range(D):
        terms = rng.choice(V, size=int(rng.integers(3, 120)), replace=False) + 1
        docs.append({int(t): int(rng.integers(1, 4)) for t in terms})
    c = corpus.from_docs(docs, V, K)
    alpha = np.full(K, 1e-3)
    g0 = np.full((D, K), 1e-3)
    g0[:, 0] = 30.0
    with capi.Context(K, V, sweep_cap=6) as ctx:
        ctx.set_corpus_csr(c.row_ptr, c.term_id, c.count, gamma0=g0)
        ctx.set_alpha(alpha)
        ctx.set_logbeta(lb)
        r = ctx.iterate()
        gamma = ctx.get_gamma()
        stats = ctx.get_stats()
    ref = O.iterate(K, V, c.row_ptr, c.term_id, c.count, lb, alpha, gamma=g0, sweeps=5)
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
    print("slow rows", r["slow_path_rows"], "launches", r["estep_launches"], "ll", r["log_likelihood"],
          ref["log_likelihood"], "finite", bool(np.all(np.isfinite(gamma))))
    print("gamma rel", rel(gamma, ref["gamma"]), "stats sum", stats.sum(), "tokens", c.n_tokens)


if __name__ == "__main__":
    main()
