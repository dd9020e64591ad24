"""End-to-end through the native driver (mrlda_b200/host/mrlda-vi): the reference's CLI options,
model files, gamma rotation and log lines around the CUDA path, checked against the oracle.

Injecting the same initial beta uses the reference's own mechanism: hand-written alpha-1 / beta-1 /
gamma-1 and `-modelindex 1` (VariationalInference.java:169-174,191-194)."""
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


def run_driver(*args):
    p = subprocess.run([DRIVER, *map(str, args)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


def likelihoods(log):
    return [float(x) for x in re.findall(r"Log likelihood after iteration \d+ is (\S+)", log)]


def test_resume_two_iterations_match_oracle(tmp_path, oracle):
    K, V = 12, 400
    c = corpus.generate(150, V, K, 50, seed=21)
    out = str(tmp_path / "model") + "/"
    os.makedirs(out + "gamma-1")
    alpha1 = np.full(K, 0.02)
    lb1 = corpus.init_logbeta(V, K, seed=5)
    rng = np.random.default_rng(1)
    gamma1 = rng.gamma(2.0, 1.0, size=(c.n_docs, K)) + 0.05
    ids = np.arange(1, c.n_docs + 1, dtype=np.int32)
    seqfile.write_alpha(out + "alpha-1", alpha1)
    seqfile.write_beta(out + "beta-1", lb1, np.zeros(K, dtype=np.float32), np.ones(V, dtype=np.uint8))
    seqfile.write_corpus(out + "gamma-1/gamma_gamma-m-00000", c.row_ptr, c.term_id, c.count, ids,
                         gamma1, K)
    log = run_driver("-input", "ignored-on-resume", "-output", out, "-term", V, "-topic", K,
                     "-iteration", 3, "-modelindex", 1, "-mapper", 7, "-reducer", 3, "-directemit")
    lls = likelihoods(log)
    assert len(lls) == 2 and "Log likelihood after iteration 2 is" in log
    assert "Total number of documents is: 150" in log

    a, lb, g = alpha1, lb1, gamma1
    for it in range(2):
        ref = oracle.iterate(K, V, c.row_ptr, c.term_id, c.count, lb, a, gamma=g)
        assert abs(lls[it] - ref["log_likelihood"]) <= 1e-9 * abs(ref["log_likelihood"])
        seen = ref["present"].astype(bool)
        a, g = ref["alpha"], ref["gamma"]
        lb = np.where(seen[:, None], ref["logbeta"], lb)
    # files of the last iteration, and the rotation: gamma-1 and gamma-2 are gone (VI.java:375-378)
    assert sorted(os.listdir(out)) == ["alpha-1", "alpha-2", "alpha-3", "beta-1", "beta-2", "beta-3",
                                       "gamma-3"]
    assert rel_err(seqfile.read_alpha(out + "alpha-3", K), ref["alpha"]) < 1e-8
    dl, nm, pr = seqfile.read_beta(out + "beta-3", K, V)
    assert np.array_equal(pr, ref["present"]) and np.array_equal(nm, ref["normaliser"])
    assert rel_err(dl[seen], ref["digamma_lambda"][seen]) < 1e-8
    got = seqfile.read_corpus(out + "gamma-3", K, V)
    assert np.array_equal(got["doc_id"], ids) and np.array_equal(got["term_id"], c.term_id)
    assert rel_err(got["gamma"], ref["gamma"]) < 1e-8
    assert f"Total number of terms is: {ref['n_terms_seen']}" in log


def test_fresh_training_then_test_mode(tmp_path, oracle):
    K, V = 8, 300
    full = corpus.generate(220, V, K, 40, seed=33)
    train, held = full.slice_docs(0, 180), full.slice_docs(180, 220)
    inp, tst, out = str(tmp_path / "train"), str(tmp_path / "test"), str(tmp_path / "out") + "/"
    os.makedirs(inp)
    seqfile.write_corpus(inp + "/part-00000", train.row_ptr, train.term_id, train.count)
    seqfile.write_corpus(tst, held.row_ptr, held.term_id, held.count)
    log = run_driver("-input", inp, "-output", out, "-term", V, "-topic", K, "-iteration", 4,
                     "-symmetricalpha", "-seed", 11)
    lls = likelihoods(log)
    assert len(lls) == 4
    assert lls[2] > lls[1] or abs(lls[2] - lls[1]) < 1e-6 * abs(lls[1])  # ELBO rises from iteration 2
    assert np.array_equal(seqfile.read_alpha(out + "alpha-0", K), np.full(K, 1e-3))  # VI.java:160
    a4 = seqfile.read_alpha(out + "alpha-4", K)
    assert np.all(a4 == a4[0]) and a4[0] > 0
    assert sorted(os.listdir(out)) == ["alpha-" + str(i) for i in range(5)] + \
        ["beta-" + str(i) for i in range(1, 5)] + ["gamma-4"]

    # held-out likelihood: -test <modelDir> -modelindex 4 (VariationalInferenceOptions.java:166-178)
    tout = str(tmp_path / "heldout") + "/"
    tlog = run_driver("-input", tst, "-output", tout, "-term", V, "-topic", K, "-test", out,
                      "-modelindex", 4, "-iteration", 6)
    tll = likelihoods(tlog)
    assert " - training mode: false" in tlog and len(tll) >= 1
    dl, nm, pr = seqfile.read_beta(out + "beta-4", K, V)
    unseen = np.unique(held.term_id[pr[held.term_id - 1] == 0])
    if len(unseen) == 0:  # every held-out term has a row in the model: fully reproducible
        lb = oracle.import_beta(K, V, dl, nm, pr)
        ref = oracle.iterate(K, V, held.row_ptr, held.term_id, held.count, lb, a4, learning=False)
        assert abs(tll[0] - ref["log_likelihood"]) <= 1e-9 * abs(ref["log_likelihood"])
    assert not any(f.startswith(("alpha-", "beta-")) for f in os.listdir(tout))  # no model written
    # the gamma rotation only ever removes gamma-N directories the driver wrote under -output; the
    # reference's fs.delete(gammaDir) would remove the user's -input here (VI.java:359-379)
    assert os.path.isfile(tst) and os.path.getsize(tst) > 0


def test_informed_prior_option(tmp_path, oracle):
    K, V = 6, 150
    c = corpus.generate(80, V, K, 30, seed=8)
    out = str(tmp_path / "m") + "/"
    os.makedirs(out + "gamma-1")
    alpha1, lb1 = np.full(K, 0.1), corpus.init_logbeta(V, K, seed=9)
    seqfile.write_alpha(out + "alpha-1", alpha1)
    seqfile.write_beta(out + "beta-1", lb1, np.zeros(K, dtype=np.float32), np.ones(V, dtype=np.uint8))
    seqfile.write_corpus(out + "gamma-1/gamma_gamma-m-00000", c.row_ptr, c.term_id, c.count)
    topics, terms = np.array([1, 1, 2, 6]), np.array([3, 7, 9, 150])
    eta = str(tmp_path / "eta-in")
    seqfile.write_eta(eta, K, topics, terms)
    log = run_driver("-input", "x", "-output", out, "-term", V, "-topic", K, "-iteration", 2,
                     "-modelindex", 1, "-informedprior", eta)
    assert "4 seed terms" in log and os.path.exists(out + "eta")
    seed = np.zeros((V, K), dtype=np.uint8)
    seed[terms - 1, topics - 1] = 1
    ref = oracle.iterate(K, V, c.row_ptr, c.term_id, c.count, lb1, alpha1, seed=seed)
    dl, nm, pr = seqfile.read_beta(out + "beta-2", K, V)
    s = ref["present"].astype(bool)
    assert np.array_equal(nm, ref["normaliser"]) and rel_err(dl[s], ref["digamma_lambda"][s]) < 1e-9
