#!/usr/bin/env python
"""Quick parity check of the warp-group E-step kernel against the oracle on a mixed-length K=200
corpus (development aid; the committed parity tests live in tests/)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrlda_b200 import capi, corpus  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    V = 5000
    parts = [corpus.generate(40, V, K, m, seed=20 + i) for i, m in enumerate((20, 60, 120, 200, 330, 700))]
    n = np.concatenate([np.diff(p.row_ptr) for p in parts])
    rp = np.zeros(len(n) + 1, dtype=np.int64)
    rp[1:] = np.cumsum(n)
    c = corpus.Corpus(rp, np.concatenate([p.term_id for p in parts]),
                      np.concatenate([p.count for p in parts]), V, K)
    print("rows per doc: min %d max %d mean %.1f" % (n.min(), n.max(), n.mean()))
    alpha = np.full(K, 1e-3)
    lb = corpus.init_logbeta(V, K, seed=7)
    with capi.Context(K, V, sweep_cap=30) as ctx:
        ctx.set_corpus_csr(c.row_ptr, c.term_id, c.count)
        ctx.set_alpha(alpha)
        ctx.set_logbeta(lb)
        r = ctx.iterate()
        gamma = ctx.get_gamma()
        ss = ctx.get_alpha_ss()
        stats = ctx.get_stats()
        nlb = ctx.get_logbeta()
    ref = O.iterate(K, V, c.row_ptr, c.term_id, c.count, lb, alpha, sweeps=29)
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
    seen = ref["present"].astype(bool)
    print("launches", r["estep_launches"], "ll", r["log_likelihood"], ref["log_likelihood"],
          "rel", abs(r["log_likelihood"] - ref["log_likelihood"]) / abs(ref["log_likelihood"]))
    print("gamma rel", rel(gamma, ref["gamma"]), "alpha_ss rel", rel(ss, ref["alpha_ss"]))
    print("stats sum", stats.sum(), "tokens", c.n_tokens, "logbeta rel", rel(nlb[seen], ref["logbeta"][seen]))
    bad = np.abs(gamma - ref["gamma"]).max(axis=1) / np.abs(ref["gamma"]).max(axis=1) > 1e-9
    print("bad docs:", np.nonzero(bad)[0][:20], n[bad][:20])


if __name__ == "__main__":
    main()
