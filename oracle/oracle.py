"""ctypes binding of oracle/libmrlda_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module (see the header of mrlda_oracle.c).  The product package mrlda_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmrlda_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "mrlda_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmrlda_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_fp = C.POINTER(C.c_float)
_bp = C.POINTER(C.c_uint8)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.mrlda_oracle_digamma.restype = C.c_double
        L.mrlda_oracle_digamma.argtypes = [C.c_double]
        L.mrlda_oracle_trigamma.restype = C.c_double
        L.mrlda_oracle_trigamma.argtypes = [C.c_double]
        L.mrlda_oracle_lngamma.restype = C.c_double
        L.mrlda_oracle_lngamma.argtypes = [C.c_double]
        L.mrlda_oracle_logadd.restype = C.c_double
        L.mrlda_oracle_logadd.argtypes = [C.c_double, C.c_double]
        L.mrlda_oracle_set_logadd_cutoff.argtypes = [C.c_double]
        L.mrlda_oracle_num_threads.restype = C.c_int
        L.mrlda_oracle_set_num_threads.restype = C.c_int
        L.mrlda_oracle_set_num_threads.argtypes = [C.c_int]
        L.mrlda_oracle_update_scalar_alpha.restype = C.c_double
        L.mrlda_oracle_update_scalar_alpha.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
        L.mrlda_oracle_update_vector_alpha.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp]
        L.mrlda_oracle_update_alpha.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]
        L.mrlda_oracle_estep.restype = C.c_int64
        L.mrlda_oracle_estep.argtypes = [C.c_int, C.c_int, C.c_int64, _lp, _ip, _ip, _dp, _dp,
                                         C.c_int, C.c_int, _dp, _dp, _dp, _lp]
        L.mrlda_oracle_mstep.restype = C.c_int64
        L.mrlda_oracle_mstep.argtypes = [C.c_int, C.c_int, C.c_int64, _lp, _ip, _dp, _dp, _fp, _bp, _bp]
        L.mrlda_oracle_import_beta.argtypes = [C.c_int, C.c_int, _dp, _fp, _bp, _dp]
        L.mrlda_oracle_iterate.restype = C.c_int64
        L.mrlda_oracle_iterate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, _lp, _ip, _ip, _dp,
                                           _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _fp,
                                           _bp, _dp, _dp, _lp, _bp]
        L.mrlda_oracle_converged.restype = C.c_int
        L.mrlda_oracle_converged.argtypes = [C.c_double, C.c_double]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def digamma(x):
    return lib().mrlda_oracle_digamma(float(x))


def trigamma(x):
    return lib().mrlda_oracle_trigamma(float(x))


def lngamma(x):
    return lib().mrlda_oracle_lngamma(float(x))


def logadd(a, b):
    return lib().mrlda_oracle_logadd(float(a), float(b))


def set_logadd_cutoff(nats):
    lib().mrlda_oracle_set_logadd_cutoff(float(nats))


def num_threads():
    return lib().mrlda_oracle_num_threads()


def set_num_threads(n=None):
    """OpenMP team size; default = the cores this process may run on (ignores OMP_NUM_THREADS,
    which torchrun sets to 1)."""
    if n is None:
        n = len(os.sched_getaffinity(0))
    return lib().mrlda_oracle_set_num_threads(int(n))


def update_vector_alpha(K, n_docs, alpha, ss):
    a = np.ascontiguousarray(alpha, dtype=np.float64)
    s = np.ascontiguousarray(ss, dtype=np.float64)
    out = np.empty(K, dtype=np.float64)
    lib().mrlda_oracle_update_vector_alpha(K, n_docs, _p(a, _dp), _p(s, _dp), _p(out, _dp))
    return out


def update_scalar_alpha(K, n_docs, alpha_init, ss):
    return lib().mrlda_oracle_update_scalar_alpha(K, n_docs, float(alpha_init), float(ss))


def update_alpha(K, n_docs, symmetric, alpha, ss):
    a = np.ascontiguousarray(alpha, dtype=np.float64)
    s = np.ascontiguousarray(ss, dtype=np.float64)
    out = np.empty(K, dtype=np.float64)
    lib().mrlda_oracle_update_alpha(K, n_docs, int(symmetric), _p(a, _dp), _p(s, _dp), _p(out, _dp))
    return out


def estep(K, row_ptr, term_id, count, logbeta, alpha, gamma=None, random_start=False, sweeps=99,
          want_log_phi=False):
    """Whole map phase.  Returns dict(gamma, alpha_ss, counter, doc_counter, log_phi)."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    term_id = np.ascontiguousarray(term_id, dtype=np.int32)
    count = np.ascontiguousarray(count, dtype=np.int32)
    logbeta = np.ascontiguousarray(logbeta, dtype=np.float64)
    alpha = np.ascontiguousarray(alpha, dtype=np.float64)
    D = len(row_ptr) - 1
    has_gamma = gamma is not None
    g = (np.array(gamma, dtype=np.float64, order="C", copy=True) if has_gamma
         else np.zeros((D, K), dtype=np.float64))
    Z = int(row_ptr[-1])
    lp = np.empty((Z, K), dtype=np.float64) if want_log_phi else None
    ss = np.zeros(K, dtype=np.float64)
    dc = np.zeros(D, dtype=np.int64)
    counter = lib().mrlda_oracle_estep(K, sweeps, D, _p(row_ptr, _lp), _p(term_id, _ip),
                                       _p(count, _ip), _p(logbeta, _dp), _p(alpha, _dp),
                                       int(has_gamma), int(random_start), _p(g, _dp), _p(lp, _dp),
                                       _p(ss, _dp), _p(dc, _lp))
    return dict(gamma=g.reshape(D, K), alpha_ss=ss, counter=int(counter), doc_counter=dc, log_phi=lp)


def import_beta(K, V, digamma_lambda, normaliser, present):
    dl = np.ascontiguousarray(digamma_lambda, dtype=np.float64)
    nm = np.ascontiguousarray(normaliser, dtype=np.float32)
    pr = np.ascontiguousarray(present, dtype=np.uint8)
    out = np.empty((V, K), dtype=np.float64)
    lib().mrlda_oracle_import_beta(K, V, _p(dl, _dp), _p(nm, _fp), _p(pr, _bp), _p(out, _dp))
    return out


def iterate(K, V, row_ptr, term_id, count, logbeta, alpha, gamma=None, random_start=False,
            learning=True, symmetric=False, sweeps=99, seed=None):
    """One EM iteration.  Returns dict(counter, log_likelihood, gamma, digamma_lambda, normaliser,
    present, alpha, alpha_ss, n_terms_seen, logbeta) with logbeta = importBeta(new beta)."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    term_id = np.ascontiguousarray(term_id, dtype=np.int32)
    count = np.ascontiguousarray(count, dtype=np.int32)
    logbeta = np.ascontiguousarray(logbeta, dtype=np.float64)
    alpha = np.ascontiguousarray(alpha, dtype=np.float64)
    D = len(row_ptr) - 1
    has_gamma = gamma is not None
    g = (np.array(gamma, dtype=np.float64, order="C", copy=True) if has_gamma
         else np.zeros((D, K), dtype=np.float64))
    dl = np.zeros((V, K), dtype=np.float64)
    nm = np.zeros(K, dtype=np.float32)
    pr = np.zeros(V, dtype=np.uint8)
    a_out = alpha.copy()
    ss = np.zeros(K, dtype=np.float64)
    seen = C.c_int64(0)
    sd = np.ascontiguousarray(seed, dtype=np.uint8) if seed is not None else None
    counter = lib().mrlda_oracle_iterate(K, V, sweeps, D, _p(row_ptr, _lp), _p(term_id, _ip),
                                         _p(count, _ip), _p(logbeta, _dp), _p(alpha, _dp),
                                         int(has_gamma), int(random_start), int(learning),
                                         int(symmetric), _p(g, _dp), _p(dl, _dp), _p(nm, _fp),
                                         _p(pr, _bp), _p(a_out, _dp), _p(ss, _dp), C.byref(seen), _p(sd, _bp))
    res = dict(counter=int(counter), log_likelihood=-int(counter) / 1e6, gamma=g.reshape(D, K),
               alpha_ss=ss, alpha=a_out, n_terms_seen=int(seen.value))
    if learning:
        res.update(digamma_lambda=dl, normaliser=nm, present=pr,
                   logbeta=import_beta(K, V, dl, nm, pr))
    return res
