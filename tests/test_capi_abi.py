"""The C-ABI library loads and exports every symbol include/mrlda_b200.h declares.
No compute calls: this runs without a GPU."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mrlda_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mrlda_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(capi):
    lib = capi.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mrlda_b200.h but not exported"
    assert sorted(capi.EXPORTS) == names


def test_struct_layouts_match_header(capi):
    assert ctypes.sizeof(capi.MrldaConfig) == 72
    assert ctypes.sizeof(capi.MrldaIterResult) == 88
    cfg = capi.default_config()
    assert (cfg.sweep_cap, cfg.learning, cfg.world_size, cfg.update_alpha) == (100, 1, 1, 1)
    assert cfg.eta == 1e-12 and cfg.random_start == 0 and cfg.symmetric_alpha == 0 and cfg.fp32_mode == 0 and cfg.deterministic == 0
    assert capi.load().mrlda_version().decode().startswith("mrlda_b200")


def test_no_cpu_fallback(capi):
    """Without a GPU mrlda_create must fail loudly (MRLDA_ECUDA), never compute on the host."""
    import torch
    if torch.cuda.is_available():
        return
    try:
        capi.Context(4, 10)
    except capi.MrldaError as e:
        assert e.code == -2 and "no CPU fallback" in str(e)
    else:
        raise AssertionError("mrlda_create succeeded without a CUDA device")


def test_product_package_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under mrlda_b200/ may import, load or link it."""
    pkg = os.path.join(ROOT, "mrlda_b200")
    bad = re.compile(r"import\s+oracle|from\s+oracle|libmrlda_oracle|mrlda_oracle_[a-z]|oracle/_ref")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not bad.search(text), f"{f} uses the oracle"


def test_argument_errors_are_codes_not_crashes(capi):
    """Argument validation happens before any CUDA call and answers with MRLDA_EINVAL plus a message
    (the reference's Preconditions.checkArgument); null handles are tolerated everywhere."""
    lib = capi.load()
    lib.mrlda_last_error.restype = ctypes.c_char_p
    lib.mrlda_last_error.argtypes = [ctypes.c_void_p]
    h = ctypes.c_void_p()
    assert lib.mrlda_create(None, ctypes.byref(h)) == -1 and not h.value
    for field, bad in (("n_topics", 0), ("n_terms", -3), ("sweep_cap", 1), ("world_size", 0), ("rank", 5),
                       ("eta", 0.0)):
        cfg = capi.default_config()
        cfg.n_topics, cfg.n_terms = 4, 10
        setattr(cfg, field, bad)
        assert lib.mrlda_create(ctypes.byref(cfg), ctypes.byref(h)) == -1, field
        assert not h.value and b"invalid config" in lib.mrlda_last_error(None)
    # a null context: every entry point returns a code (or 0 launches), none dereferences it
    assert lib.mrlda_e_step(None, None) == -1
    assert lib.mrlda_m_step(None, None) == -1
    assert lib.mrlda_iterate(None, None) == -1
    assert lib.mrlda_kernel_launches(None) == 0
    lib.mrlda_destroy(None)


def test_kernel_plan_by_topic_count(capi):
    """mrlda_describe_kernels (pure host code): which E-step kernel family a topic count gets.  The
    bounds are measured crossovers (tools/kcross_on_box.sh): the
    register-stationary kernel up to 104 topics and from 209 to 256, single-CTA warp groups from 105 to
    208, warp groups with the topics split across a cluster above 256."""
    first = lambda K: capi.describe_kernels(K).split(" | ")[0]
    for K in (2, 20, 100, 104):
        assert first(K).startswith("estep2_kernel<") and " x 1 CTA " in first(K), (K, first(K))
    for K in (105, 160, 200, 208):
        assert first(K).startswith("estep3_kernel<8,") and "<= " in first(K), (K, first(K))
    assert first(200).startswith("estep3_kernel<8,25,2,8,G>") and "<= 320 rows" in first(200)
    for K in (209, 256):
        assert first(K).startswith("estep2_kernel<8,") and " x 1 CTA " in first(K), (K, first(K))
    for K, cl in ((257, 2), (400, 2), (513, 3), (800, 4), (1000, 5), (1024, 5)):
        assert first(K).startswith("estep3c_kernel<8,") and f" x {cl} CTAs " in first(K), (K, first(K))
    plan = capi.describe_kernels(1000)
    assert plan.startswith("estep3c_kernel<8,25,2,8,G,CL> x 5 CTAs per document")  # K = 1000 is 5 x 200
    assert "estep2_kernel<8,16,4,8,8> x 8 CTAs per document, <= 512 rows" in plan
    assert plan.endswith("estep_kernel<32,32> longer documents and deterministic mode")
    import pytest
    with pytest.raises(capi.MrldaError):
        capi.describe_kernels(0)
    with pytest.raises(capi.MrldaError):
        capi.describe_kernels(100000)


def test_binding_checks_buffer_extents_before_the_call(capi):
    """The ABI takes plain pointers and trusts their extents, so the ctypes binding refuses short or
    ill-typed buffers itself (ValueError) before any pointer reaches the library.  Runs without a
    GPU: the checks come before the C call, on a Context that owns no handle."""
    import numpy as np
    import pytest
    ctx = capi.Context.__new__(capi.Context)
    ctx.K, ctx.V, ctx.D, ctx._h = 3, 5, 2, None
    rp = np.array([0, 2, 3], dtype=np.int64)
    with pytest.raises(ValueError, match="row_ptr"):
        ctx.set_corpus_csr(np.zeros(0, np.int64), [], [])
    with pytest.raises(ValueError, match="shorter"):
        ctx.set_corpus_csr(rp, [1, 2], [1, 1, 1])
    with pytest.raises(ValueError, match="shorter"):
        ctx.set_corpus_csr(rp, [1, 2, 3], [1, 1])
    with pytest.raises(ValueError, match="doc_id"):
        ctx.set_corpus_csr(rp, [1, 2, 3], [1, 1, 1], doc_id=[7])
    with pytest.raises(ValueError, match="gamma0"):
        ctx.set_corpus_csr(rp, [1, 2, 3], [1, 1, 1], gamma0=np.ones(5))
    with pytest.raises(ValueError):
        ctx.set_alpha(np.ones(2))
    with pytest.raises(ValueError):
        ctx.set_logbeta(np.zeros((5, 2)))
    with pytest.raises(ValueError):
        ctx.set_beta(np.zeros((5, 3)), np.zeros(3, np.float32), np.zeros(4, np.uint8))
    with pytest.raises(ValueError, match="same length"):
        ctx.set_informed_prior([1, 2], [1])
    for bad in (np.empty((2, 2)), np.empty((2, 3), np.float32), np.empty((3, 2)).T, [0.0] * 6):
        with pytest.raises(ValueError, match="out must be"):
            ctx.get_gamma(out=bad)
    with pytest.raises(ValueError):
        ctx.update_vector_alpha(3, 10, np.ones(3), np.ones(2))
    ctx._h = capi._ctxp()  # so that __del__ finds nothing to destroy
