"""ctypes binding of libmrlda_b200.so (include/mrlda_b200.h).

This is the same C ABI a Java shim binds over JNI/Panama.  There is no fallback:
if the shared library is missing, or no B200 is visible, the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmrlda_b200.so")

MRLDA_NCCL_ID_BYTES = 128

EXPORTS = [
    "mrlda_default_config", "mrlda_create", "mrlda_destroy", "mrlda_last_error",
    "mrlda_set_corpus_csr", "mrlda_set_alpha", "mrlda_get_alpha", "mrlda_set_beta",
    "mrlda_get_beta", "mrlda_set_informed_prior", "mrlda_set_logbeta", "mrlda_get_logbeta", "mrlda_init_beta_random",
    "mrlda_iterate", "mrlda_e_step", "mrlda_m_step", "mrlda_get_gamma", "mrlda_get_alpha_ss",
    "mrlda_get_stats", "mrlda_update_vector_alpha", "mrlda_update_scalar_alpha",
    "mrlda_eval_special", "mrlda_measure_fp64_peak", "mrlda_debug_phase_cycles", "mrlda_nccl_unique_id", "mrlda_comm_init", "mrlda_kernel_launches",
    "mrlda_version", "mrlda_describe_kernels",
]


class MrldaConfig(C.Structure):
    _fields_ = [
        ("n_topics", C.c_int32), ("n_terms", C.c_int32), ("sweep_cap", C.c_int32),
        ("learning", C.c_int32), ("random_start", C.c_int32), ("symmetric_alpha", C.c_int32),
        ("device", C.c_int32), ("rank", C.c_int32), ("world_size", C.c_int32),
        ("update_alpha", C.c_int32), ("fp32_mode", C.c_int32), ("eta", C.c_double),
        ("seed", C.c_uint64), ("deterministic", C.c_int32), ("reserved", C.c_int32),
    ]


class MrldaIterResult(C.Structure):
    _fields_ = [
        ("ll_counter", C.c_int64), ("log_likelihood", C.c_double), ("n_docs", C.c_int64),
        ("n_terms_seen", C.c_int64), ("n_tokens", C.c_int64), ("e_step_ms", C.c_double),
        ("m_step_ms", C.c_double), ("alpha_ms", C.c_double), ("allreduce_ms", C.c_double),
        ("estep_launches", C.c_int32), ("total_launches", C.c_int32),
        ("slow_path_rows", C.c_int32), ("reserved", C.c_int32),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_ if name != "reserved"}


class MrldaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mrlda error {code}: {msg}")
        self.code = code


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_fp = C.POINTER(C.c_float)
_bp = C.POINTER(C.c_uint8)
_ctxp = C.c_void_p


def load():
    """Loads the shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} not found: build it with `make -C mrlda_b200/csrc` (or "
            f"`python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.mrlda_default_config.argtypes = [C.POINTER(MrldaConfig)]
    L.mrlda_default_config.restype = None
    L.mrlda_create.argtypes = [C.POINTER(MrldaConfig), C.POINTER(_ctxp)]
    L.mrlda_destroy.argtypes = [_ctxp]
    L.mrlda_destroy.restype = None
    L.mrlda_last_error.argtypes = [_ctxp]
    L.mrlda_last_error.restype = C.c_char_p
    L.mrlda_set_corpus_csr.argtypes = [_ctxp, C.c_int64, _lp, _ip, _ip, _ip, _dp]
    L.mrlda_set_alpha.argtypes = [_ctxp, _dp]
    L.mrlda_get_alpha.argtypes = [_ctxp, _dp]
    L.mrlda_set_beta.argtypes = [_ctxp, _dp, _fp, _bp]
    L.mrlda_get_beta.argtypes = [_ctxp, _dp, _fp, _bp]
    L.mrlda_set_informed_prior.argtypes = [_ctxp, C.c_int64, _ip, _ip]
    L.mrlda_set_logbeta.argtypes = [_ctxp, _dp]
    L.mrlda_get_logbeta.argtypes = [_ctxp, _dp]
    L.mrlda_init_beta_random.argtypes = [_ctxp]
    L.mrlda_iterate.argtypes = [_ctxp, C.POINTER(MrldaIterResult)]
    L.mrlda_e_step.argtypes = [_ctxp, C.POINTER(MrldaIterResult)]
    L.mrlda_m_step.argtypes = [_ctxp, C.POINTER(MrldaIterResult)]
    L.mrlda_get_gamma.argtypes = [_ctxp, _dp]
    L.mrlda_get_alpha_ss.argtypes = [_ctxp, _dp]
    L.mrlda_get_stats.argtypes = [_ctxp, _dp]
    L.mrlda_update_vector_alpha.argtypes = [_ctxp, C.c_int32, C.c_int32, _dp, _dp, _dp]
    L.mrlda_update_scalar_alpha.argtypes = [_ctxp, C.c_int32, C.c_int32, C.c_double, C.c_double, _dp]
    L.mrlda_eval_special.argtypes = [_ctxp, C.c_int32, C.c_int64, _dp, _dp]
    L.mrlda_measure_fp64_peak.argtypes = [_ctxp, _dp]
    L.mrlda_debug_phase_cycles.argtypes = [_ctxp, _lp]
    L.mrlda_nccl_unique_id.argtypes = [_bp]
    L.mrlda_comm_init.argtypes = [_ctxp, _bp]
    L.mrlda_kernel_launches.argtypes = [_ctxp]
    L.mrlda_kernel_launches.restype = C.c_int64
    L.mrlda_version.restype = C.c_char_p
    L.mrlda_describe_kernels.argtypes = [C.c_int, C.c_char_p, C.c_size_t]
    L.mrlda_describe_kernels.restype = C.c_int
    _lib = L
    return L


def default_config():
    cfg = MrldaConfig()
    load().mrlda_default_config(C.byref(cfg))
    return cfg


def describe_kernels(n_topics):
    """Which E-step kernels a context with this many topics uses (no GPU needed)."""
    buf = C.create_string_buffer(1024)
    rc = load().mrlda_describe_kernels(int(n_topics), buf, len(buf))
    if rc < 0:
        raise MrldaError(rc, f"mrlda_describe_kernels({n_topics})")
    return buf.value.decode()


def nccl_unique_id():
    buf = (C.c_uint8 * MRLDA_NCCL_ID_BYTES)()
    rc = load().mrlda_nccl_unique_id(buf)
    if rc:
        raise MrldaError(rc, load().mrlda_last_error(None).decode())
    return bytes(buf)


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Context:
    """Owns one mrlda_ctx.  Thin: every method is one C-ABI call with host numpy buffers."""

    def __init__(self, n_topics, n_terms, **kw):
        L = load()
        cfg = default_config()
        cfg.n_topics = int(n_topics)
        cfg.n_terms = int(n_terms)
        for k, v in kw.items():
            if not hasattr(cfg, k):
                raise TypeError(f"unknown config field {k}")
            setattr(cfg, k, v)
        self.cfg = cfg
        self.K, self.V = int(n_topics), int(n_terms)
        self.D = 0
        self._h = _ctxp()
        rc = L.mrlda_create(C.byref(cfg), C.byref(self._h))
        if rc:
            raise MrldaError(rc, L.mrlda_last_error(None).decode())

    def _ck(self, rc):
        if rc:
            raise MrldaError(rc, load().mrlda_last_error(self._h).decode())

    def close(self):
        if self._h:
            load().mrlda_destroy(self._h)
            self._h = _ctxp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def comm_init(self, unique_id: bytes):
        buf = (C.c_uint8 * MRLDA_NCCL_ID_BYTES).from_buffer_copy(unique_id)
        self._ck(load().mrlda_comm_init(self._h, buf))

    def set_corpus_csr(self, row_ptr, term_id, count, doc_id=None, gamma0=None):
        # the ABI takes plain pointers and trusts their extents: check every length here
        row_ptr = _arr(row_ptr, np.int64).ravel()
        term_id = _arr(term_id, np.int32).ravel()
        count = _arr(count, np.int32).ravel()
        if row_ptr.size < 1:
            raise ValueError("row_ptr must hold D+1 values")
        D = row_ptr.size - 1
        Z = int(row_ptr[D])
        if Z < 0 or term_id.size < Z or count.size < Z:
            raise ValueError("term_id / count are shorter than row_ptr[D]")
        did = _arr(doc_id, np.int32) if doc_id is not None else None
        if did is not None and did.size != D:
            raise ValueError("doc_id must hold D values")
        g0 = _arr(gamma0, np.float64) if gamma0 is not None else None
        if g0 is not None and g0.size != D * self.K:
            raise ValueError("gamma0 must hold D*K values")
        self._ck(load().mrlda_set_corpus_csr(
            self._h, D, row_ptr.ctypes.data_as(_lp), term_id.ctypes.data_as(_ip),
            count.ctypes.data_as(_ip), did.ctypes.data_as(_ip) if did is not None else None,
            g0.ctypes.data_as(_dp) if g0 is not None else None))
        self.D = D

    def set_alpha(self, alpha):
        a = _arr(alpha, np.float64)
        if a.size != self.K:
            raise ValueError("alpha must hold K values")
        self._ck(load().mrlda_set_alpha(self._h, a.ctypes.data_as(_dp)))

    def get_alpha(self):
        a = np.empty(self.K, dtype=np.float64)
        self._ck(load().mrlda_get_alpha(self._h, a.ctypes.data_as(_dp)))
        return a

    def set_logbeta(self, logbeta):
        b = _arr(logbeta, np.float64)
        if b.size != self.V * self.K:
            raise ValueError("logbeta must hold V*K values (term-major)")
        self._ck(load().mrlda_set_logbeta(self._h, b.ctypes.data_as(_dp)))

    def get_logbeta(self):
        b = np.empty((self.V, self.K), dtype=np.float64)
        self._ck(load().mrlda_get_logbeta(self._h, b.ctypes.data_as(_dp)))
        return b

    def init_beta_random(self):
        self._ck(load().mrlda_init_beta_random(self._h))

    def set_beta(self, digamma_lambda, normaliser, present):
        dl = _arr(digamma_lambda, np.float64)
        nm = _arr(normaliser, np.float32)
        pr = _arr(present, np.uint8)
        if dl.size != self.V * self.K or nm.size != self.K or pr.size != self.V:
            raise ValueError("bad beta shapes")
        self._ck(load().mrlda_set_beta(self._h, dl.ctypes.data_as(_dp), nm.ctypes.data_as(_fp),
                                       pr.ctypes.data_as(_bp)))

    def get_beta(self):
        dl = np.empty((self.V, self.K), dtype=np.float64)
        nm = np.empty(self.K, dtype=np.float32)
        pr = np.empty(self.V, dtype=np.uint8)
        self._ck(load().mrlda_get_beta(self._h, dl.ctypes.data_as(_dp), nm.ctypes.data_as(_fp),
                                       pr.ctypes.data_as(_bp)))
        return dl, nm, pr

    def set_informed_prior(self, topics, terms):
        """(topic, term) seed pairs, both 1-based; empty lists remove the prior."""
        t = _arr(topics, np.int32).ravel()
        v = _arr(terms, np.int32).ravel()
        if t.size != v.size:
            raise ValueError("topics and terms must have the same length")
        self._ck(load().mrlda_set_informed_prior(self._h, len(t), t.ctypes.data_as(_ip),
                                                 v.ctypes.data_as(_ip)))

    def iterate(self):
        r = MrldaIterResult()
        self._ck(load().mrlda_iterate(self._h, C.byref(r)))
        return r.as_dict()

    def e_step(self):
        r = MrldaIterResult()
        self._ck(load().mrlda_e_step(self._h, C.byref(r)))
        return r.as_dict()

    def m_step(self):
        r = MrldaIterResult()
        self._ck(load().mrlda_m_step(self._h, C.byref(r)))
        return r.as_dict()

    def get_gamma(self, out=None):
        g = out if out is not None else np.empty((self.D, self.K), dtype=np.float64)
        if (not isinstance(g, np.ndarray) or g.dtype != np.float64 or not g.flags.c_contiguous
                or not g.flags.writeable or g.size != self.D * self.K):
            raise ValueError("out must be a writable C-contiguous float64 array of D*K values")
        self._ck(load().mrlda_get_gamma(self._h, g.ctypes.data_as(_dp)))
        return g

    def get_alpha_ss(self):
        a = np.empty(self.K, dtype=np.float64)
        self._ck(load().mrlda_get_alpha_ss(self._h, a.ctypes.data_as(_dp)))
        return a

    def get_stats(self):
        s = np.empty((self.V, self.K), dtype=np.float64)
        self._ck(load().mrlda_get_stats(self._h, s.ctypes.data_as(_dp)))
        return s

    def update_vector_alpha(self, K, n_docs, alpha, ss):
        a = _arr(alpha, np.float64)
        s = _arr(ss, np.float64)
        if K < 1 or a.size != K or s.size != K:
            raise ValueError("alpha and ss must hold K values")
        out = np.empty(K, dtype=np.float64)
        self._ck(load().mrlda_update_vector_alpha(self._h, K, n_docs, a.ctypes.data_as(_dp),
                                                  s.ctypes.data_as(_dp), out.ctypes.data_as(_dp)))
        return out

    def update_scalar_alpha(self, K, n_docs, alpha_init, ss):
        out = C.c_double(0)
        self._ck(load().mrlda_update_scalar_alpha(self._h, K, n_docs, float(alpha_init), float(ss),
                                                  C.byref(out)))
        return out.value

    def eval_special(self, which, x):
        names = {"digamma": 0, "trigamma": 1, "lngamma": 2, "exp_digamma": 3, "exp_digamma_n": 4}
        x = _arr(x, np.float64)
        y = np.empty_like(x)
        self._ck(load().mrlda_eval_special(self._h, names[which], x.size, x.ctypes.data_as(_dp),
                                           y.ctypes.data_as(_dp)))
        return y

    def measure_fp64_peak(self):
        out = C.c_double(0)
        self._ck(load().mrlda_measure_fp64_peak(self._h, C.byref(out)))
        return out.value

    def debug_phase_cycles(self):
        out = np.zeros(16, dtype=np.int64)
        self._ck(load().mrlda_debug_phase_cycles(self._h, out.ctypes.data_as(_lp)))
        return out

    def kernel_launches(self):
        return int(load().mrlda_kernel_launches(self._h))
