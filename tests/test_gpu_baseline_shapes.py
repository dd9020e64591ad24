"""GPU parity at the shapes BASELINE.json names (configs 1-5), through the C ABI and the native
driver, against the CPU oracle on the same inputs.

Tolerance, stated literally (north_star: "per-iteration ELBO and logbeta within 1e-9 relative"):
  * ELBO (decoded likelihood), gamma, alpha, alpha statistics: |gpu - ref| <= 1e-9 |ref|;
  * counts, `present`, the (float) topic normalisers: bit-exact;
  * logbeta, every cell with lambda >= 1e-3: |gpu - ref| <= 1e-9 |ref|, no floor, no extra term;
  * logbeta, cells with lambda < 1e-3: 1e-9 + 2e-15 |ref| relative.  Deviation from north_star's
    wording, caused by the reference's own digamma series (it evaluates 1/((x+6)-6), so a last-bit
    difference in lambda moves the value by up to ulp(6)/lambda relative).
Each run appends its measured worst cases to gpurun_out/r2_parity_shapes.jsonl (when that directory
exists) — the numbers DESIGN.md quotes per config.
"""
import json
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import rel_err
from mrlda_b200 import corpus, seqfile

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "mrlda_b200", "host", "mrlda-vi")
TOL = 1e-9
ETA = 1e-12


def _report(name, **kw):
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "r2_parity_shapes.jsonl"), "a") as f:
            f.write(json.dumps(dict(case=name, **kw)) + "\n")


def _check_iteration(name, it, ctx, r, ref, c, learning=True):
    """Literal checks of one EM iteration; returns the measured worst cases."""
    D = c.n_docs
    m = {}
    assert r["n_docs"] == D and r["n_tokens"] == c.n_tokens
    assert abs(r["ll_counter"] - ref["counter"]) <= D
    m["elbo"] = abs(r["log_likelihood"] - ref["log_likelihood"]) / abs(ref["log_likelihood"])
    nz = np.diff(c.row_ptr) > 0
    m["gamma"] = rel_err(ctx.get_gamma()[nz], ref["gamma"][nz])
    m["alpha_ss"] = rel_err(ctx.get_alpha_ss(), ref["alpha_ss"])
    if learning:
        assert r["n_terms_seen"] == ref["n_terms_seen"]
        dl, nm, pr = ctx.get_beta()
        assert np.array_equal(pr, ref["present"])
        assert np.array_equal(nm.view(np.uint32), ref["normaliser"].view(np.uint32))
        seen = pr.astype(bool)
        lam = ctx.get_stats()[seen] + ETA
        got, want = ctx.get_logbeta()[seen], ref["logbeta"][seen]
        err = np.abs(got - want) / np.abs(want)
        big = lam >= 1e-3
        m["logbeta_lambda_ge_1e-3"] = float(err[big].max()) if big.any() else 0.0
        small = ~big
        if small.any():
            bound = TOL + 2e-15 * np.abs(want[small])
            m["logbeta_small_lambda_over_bound"] = float((err[small] / bound).max())
            assert m["logbeta_small_lambda_over_bound"] < 1.0, (name, it, m)
        m["logbeta_all_cells"] = float(err.max())
        m["alpha"] = rel_err(ctx.get_alpha(), ref["alpha"])
        assert m["logbeta_lambda_ge_1e-3"] < TOL, (name, it, m)
        assert m["alpha"] < TOL, (name, it, m)
    assert m["elbo"] < TOL and m["gamma"] < TOL and m["alpha_ss"] < TOL, (name, it, m)
    _report(name, iteration=it + 1, launches=r["estep_launches"], e_step_ms=r["e_step_ms"], **m)
    return m


def _fit(name, capi, oracle, c, iterations, **cfg):
    """`iterations` consecutive EM iterations on the GPU; the oracle redoes each one from the state
    the GPU started it with (per-iteration parity, not a free-running second trajectory)."""
    K, V = c.n_topics, c.n_terms
    alpha = np.full(K, 1e-3)  # VariationalInference.java:160
    lb = corpus.init_logbeta(V, K, seed=K + 7)
    gamma = None
    with capi.Context(K, V, **cfg) as ctx:
        ctx.set_corpus_csr(c.row_ptr, c.term_id, c.count)
        ctx.set_alpha(alpha)
        ctx.set_logbeta(lb)
        for it in range(iterations):
            r = ctx.iterate()
            ref = oracle.iterate(K, V, c.row_ptr, c.term_id, c.count, lb, alpha, gamma=gamma,
                                 symmetric=bool(cfg.get("symmetric_alpha", 0)))
            _check_iteration(name, it, ctx, r, ref, c)
            if it + 1 < iterations:
                lb, alpha, gamma = ctx.get_logbeta(), ctx.get_alpha(), ctx.get_gamma()
    return r


def test_cfg1_ap_full_shape_three_iterations(capi, oracle):
    """configs[0]: D = 2 246, V = 10 473, K = 20, vector alpha, three consecutive iterations."""
    D, V, K, mean_len, _ = corpus.SHAPES["ap"]
    c = corpus.generate(D, V, K, mean_len, seed=1, device="cuda")
    assert c.n_docs == 2246 and c.n_terms == 10473
    _fit("cfg1_ap", capi, oracle, c, 3)


@pytest.mark.parametrize("symmetric", [1, 0], ids=["cfg2_symmetric", "cfg3_vector_alpha"])
def test_cfg2_cfg3_20news_shape(capi, oracle, symmetric):
    """configs[1] / configs[2]: V = 8 000, K = 100, 2 200 documents, two iterations, -symmetricalpha
    and the Newton-Raphson vector update."""
    c = corpus.generate(2200, 8000, 100, 120, seed=2, device="cuda")
    _fit("cfg2_symmetric" if symmetric else "cfg3_vector", capi, oracle, c, 2, symmetric_alpha=symmetric)


def test_cfg4_wikipedia_shape_slice(capi, oracle):
    """configs[3]: K = 200, V = 100 000, 2 000 documents of the bench corpus's distribution plus
    long ones, so that every warp-group class and the long-document kernel take part."""
    a = corpus.generate(2000, 100_000, 200, 200, seed=4, device="cuda")
    b = corpus.generate(6, 100_000, 200, 900, seed=5, device="cuda")  # several hundred rows each
    n = np.concatenate([np.diff(a.row_ptr), np.diff(b.row_ptr)])
    assert n.max() > 320 and n.min() < 40
    rp = np.zeros(len(n) + 1, dtype=np.int64)
    rp[1:] = np.cumsum(n)
    c = corpus.Corpus(rp, np.concatenate([a.term_id, b.term_id]), np.concatenate([a.count, b.count]),
                      100_000, 200)
    r = _fit("cfg4_wikipedia_slice", capi, oracle, c, 1)
    assert r["estep_launches"] >= 3  # several document classes ran


def test_cfg5_largek_shape_slice(capi, oracle):
    """configs[4]: K = 1 000, V = 200 000, heavy-tailed lengths: the cluster kernel (documents up
    to 256 rows) and the long-document kernel both run."""
    c = corpus.generate(260, 200_000, 1000, 200, seed=5, heavy_tail=True, device="cuda")
    n = np.diff(c.row_ptr)
    assert n.max() > 256 and (n <= 256).sum() > 100
    r = _fit("cfg5_largek_slice", capi, oracle, c, 1)
    assert r["estep_launches"] >= 2


def _run_driver(*args):
    p = subprocess.run([DRIVER, *map(str, args)], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


def test_cfg2_train_then_heldout_through_driver(tmp_path, oracle):
    """configs[1] as the reference runs it: train on 90 %, then `-test <model> -modelindex N` on the
    held-out 10 % (VariationalInferenceOptions.java:166-178); both likelihood logs against the
    oracle.  The injected initial beta needs `-modelindex 1`, and the reference refuses
    `-symmetricalpha` on a resumed run (VariationalInferenceOptions.java:180-186), so the training
    part uses the vector alpha update; the symmetric update at this shape is covered through the
    ABI by test_cfg2_cfg3_20news_shape."""
    K, V = 100, 8000
    full = corpus.generate(2200, V, K, 120, seed=2, device="cuda")
    n_train = 1980
    train, held = full.slice_docs(0, n_train), full.slice_docs(n_train, full.n_docs)
    # keep the held-out entries whose term occurs in the training split: a term without a row in the
    # model would be initialised from Math.random() (DocumentMapper.java:446-463), unreproducible
    in_train = np.zeros(V + 1, dtype=bool)
    in_train[train.term_id] = True
    keep = in_train[held.term_id]
    doc_of = np.repeat(np.arange(held.n_docs), np.diff(held.row_ptr))
    hrp = np.zeros(held.n_docs + 1, dtype=np.int64)
    hrp[1:] = np.cumsum(np.bincount(doc_of[keep], minlength=held.n_docs))
    held = corpus.Corpus(hrp, held.term_id[keep], held.count[keep], V, K)
    out = str(tmp_path / "model") + "/"
    os.makedirs(out + "gamma-1")
    alpha1 = np.full(K, 1e-3)
    lb1 = corpus.init_logbeta(V, K, seed=3)
    seqfile.write_alpha(out + "alpha-1", alpha1)
    seqfile.write_beta(out + "beta-1", lb1, np.zeros(K, dtype=np.float32), np.ones(V, dtype=np.uint8))
    seqfile.write_corpus(out + "gamma-1/gamma_gamma-m-00000", train.row_ptr, train.term_id, train.count)
    log = _run_driver("-input", "unused-on-resume", "-output", out, "-term", V, "-topic", K,
                      "-iteration", 3, "-modelindex", 1)
    lls = [float(x) for x in re.findall(r"Log likelihood after iteration \d+ is (\S+)", log)]
    assert len(lls) == 2
    a, lb, g = alpha1, lb1, None
    for it in range(2):
        ref = oracle.iterate(K, V, train.row_ptr, train.term_id, train.count, lb, a, gamma=g)
        assert abs(lls[it] - ref["log_likelihood"]) <= TOL * abs(ref["log_likelihood"]), it
        seen = ref["present"].astype(bool)
        a, g = ref["alpha"], ref["gamma"]
        lb = np.where(seen[:, None], ref["logbeta"], lb)
    a3 = seqfile.read_alpha(out + "alpha-3", K)
    assert rel_err(a3, ref["alpha"]) < 1e-8  # two iterations of compounding on top of 1e-9 each
    # held-out pass
    tst = str(tmp_path / "heldout-in")
    seqfile.write_corpus(tst, held.row_ptr, held.term_id, held.count)
    tout = str(tmp_path / "heldout-out") + "/"
    tlog = _run_driver("-input", tst, "-output", tout, "-term", V, "-topic", K, "-test", out,
                       "-modelindex", 3, "-iteration", 5)
    tll = [float(x) for x in re.findall(r"Log likelihood after iteration \d+ is (\S+)", tlog)]
    assert " - training mode: false" in tlog and len(tll) >= 1
    assert os.path.exists(tst), "the -test input must survive the gamma rotation"
    dl, nm, pr = seqfile.read_beta(out + "beta-3", K, V)
    model_lb = oracle.import_beta(K, V, dl, nm, pr)
    known = pr[held.term_id - 1] == 1
    if known.all():  # every held-out term has a row in the model: reproducible without the RNG
        href = oracle.iterate(K, V, held.row_ptr, held.term_id, held.count, model_lb, a3, learning=False)
        err = abs(tll[0] - href["log_likelihood"]) / abs(href["log_likelihood"])
        _report("cfg2_heldout_driver", elbo=err, heldout_docs=held.n_docs, train_elbo=lls)
        assert err <= TOL
    else:
        _report("cfg2_heldout_driver", skipped_unseen_terms=int((~known).sum()))
