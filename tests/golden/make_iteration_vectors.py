#!/usr/bin/env python
"""Writes tests/golden/iteration_vectors.json: small, fixed inputs of one EM iteration and the outputs
the reference's algorithm produces for them, for four cases (vector alpha cold start, symmetric
alpha, stored gamma, held-out mode).

The reference is Java on Hadoop and cannot run in this image, so the outputs are NOT reference
output: they are what the C restatement (oracle/mrlda_oracle.c) computes, accepted into the fixture
only where the second, independently written restatement (oracle/mirror.py: NumPy in the
reference's fold order, written from DocumentMapper.java:175-259,402-429 and
TermReducer.java:134-238 alone) agrees to 1e-12 — the script refuses to write otherwise.  What the
fixture pins is therefore: neither restatement, nor the CUDA path, drifts from the values both
restatements agreed on when this was generated.  Floats are stored as Python reprs (shortest
round-trip form: bit-exact).   usage: python tests/golden/make_iteration_vectors.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from mrlda_b200 import corpus  # noqa: E402
from oracle import mirror as M  # noqa: E402
from oracle import oracle as O  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def case(name, D, V, K, mean_len, seed, symmetric=False, learning=True, stored_gamma=False):
    c = corpus.generate(D, V, K, mean_len, seed=seed)
    lb = corpus.init_logbeta(V, K, seed=seed + 100)
    rng = np.random.default_rng(seed)
    alpha = np.full(K, 1e-3) if not stored_gamma else rng.uniform(0.05, 0.6, K)
    if symmetric:
        alpha = np.full(K, 0.05)
    gamma0 = rng.uniform(0.1, 5.0, (D, K)) if stored_gamma else None
    ref = O.iterate(K, V, c.row_ptr, c.term_id, c.count, lb, alpha, gamma=gamma0, learning=learning,
                    symmetric=symmetric)
    # the independent restatement must agree before anything is written
    m = M.map_phase(K, c.row_ptr, c.term_id, c.count, lb, alpha, gamma=gamma0)
    assert abs(m["counter"] - ref["counter"]) <= D, (name, m["counter"], ref["counter"])
    assert rel(m["gamma"], ref["gamma"]) < 1e-12 and rel(m["alpha_ss"], ref["alpha_ss"]) < 1e-12, name
    out = dict(name=name, K=K, V=V, symmetric=symmetric, learning=learning,
               row_ptr=c.row_ptr.tolist(), term_id=c.term_id.tolist(), count=c.count.tolist(),
               logbeta=lb.ravel().tolist(), alpha=alpha.tolist(),
               gamma0=None if gamma0 is None else gamma0.ravel().tolist(),
               expect=dict(counter=ref["counter"], gamma=ref["gamma"].ravel().tolist(),
                           alpha_ss=ref["alpha_ss"].tolist()))
    if learning:
        dl, nrm, present, lam = M.reduce_terms(K, V, c.row_ptr, c.term_id, m["log_phi"])
        seen = present.astype(bool)
        assert np.array_equal(present, ref["present"]) and np.array_equal(nrm.view(np.uint32), ref["normaliser"].view(np.uint32)), name
        ok = lam[seen] >= 1e-3
        assert rel(dl[seen][ok], ref["digamma_lambda"][seen][ok]) < 1e-12, name
        out["expect"].update(n_terms_seen=ref["n_terms_seen"], present=ref["present"].tolist(),
                             normaliser=[float(x) for x in ref["normaliser"]],
                             digamma_lambda=ref["digamma_lambda"].ravel().tolist(),
                             logbeta=ref["logbeta"].ravel().tolist(), alpha=ref["alpha"].tolist())
    return out


if __name__ == "__main__":
    # K = 8 throughout: a topic count the GPU suite already exercises in every one of these modes
    cases = [case("vector_alpha_cold_start", 24, 60, 8, 30, 1),
             case("symmetric_alpha", 20, 50, 8, 25, 2, symmetric=True),
             case("stored_gamma", 16, 40, 8, 20, 3, stored_gamma=True),
             case("heldout_mode", 18, 45, 8, 20, 4, learning=False)]
    path = os.path.join(HERE, "iteration_vectors.json")
    with open(path, "w") as f:
        json.dump(dict(note="see make_iteration_vectors.py: restatement-agreed values, not reference output",
                       cases=cases), f)
    print(path, os.path.getsize(path), "bytes")
