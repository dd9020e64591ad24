"""The native driver's option handling, which needs no GPU: the reference logs the parse error,
prints the usage and exits 0 (VariationalInferenceOptions.java:57-266); a run that gets past the
options fails loudly without a CUDA device (no CPU fallback)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "mrlda_b200", "host", "mrlda-vi")


def test_option_errors_print_help_and_exit_zero():
    for args in (["-output", "x", "-term", "5", "-topic", "2"],                      # -input missing
                 ["-input", "a", "-output", "b", "-term", "5", "-topic", "0"],      # topics <= 0
                 ["-input", "a", "-output", "b", "-term", "5", "-topic", "2", "-test", "m"],  # no index
                 ["-input", "a", "-output", "b", "-term", "5", "-topic", "2", "-modelindex", "3",
                  "-iteration", "3"],                                                 # index >= iteration
                 ["-input", "a", "-output", "b", "-term", "5", "-topic", "2", "-modelindex", "1",
                  "-symmetricalpha"]):                                               # symmetric + resume
        p = subprocess.run([DRIVER, *args], capture_output=True, text=True, timeout=60)
        assert p.returncode == 0 and "usage: cc.mrlda.VariationalInference" in p.stderr


def test_help_lists_every_reference_option():
    p = subprocess.run([DRIVER, "-help"], capture_output=True, text=True, timeout=60)
    text = p.stdout + p.stderr
    assert p.returncode == 0
    for opt in ("input", "output", "term", "topic", "iteration", "mapper", "reducer", "test", "modelindex",
                "randomstart", "symmetricalpha", "directemit", "truncatebeta", "informedprior", "help"):
        assert f"-{opt}" in text, opt


def test_no_cpu_fallback_without_a_cuda_device(tmp_path):
    """Past the options the driver needs the CUDA library and a B200: without a device it stops with
    an error and produces no iteration output (the product path has no CPU implementation)."""
    import numpy as np
    import pytest
    import torch

    from mrlda_b200 import seqfile

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    inp = tmp_path / "corpus"
    inp.mkdir()
    seqfile.write_corpus(str(inp / "part-00000"), np.array([0, 2, 3]), np.array([1, 2, 3]), np.array([2, 1, 4]))
    out = tmp_path / "out"
    p = subprocess.run([DRIVER, "-input", str(inp), "-output", str(out), "-term", "3", "-topic", "2",
                        "-iteration", "1"], capture_output=True, text=True, timeout=120)
    assert p.returncode != 0, p.stdout + p.stderr
    assert "mrlda" in (p.stdout + p.stderr).lower()
    # alpha-0 may exist (the reference writes the initial alpha before its first job,
    # VariationalInference.java:160-165); nothing an iteration produces may
    made = os.listdir(out) if out.exists() else []
    assert not any(f in ("alpha-1", "beta-1", "gamma-1") for f in made), made
