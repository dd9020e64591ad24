"""ctypes binding of mrlda_b200/host/libmrlda_seqfile.so: the Writable / SequenceFile codecs the
native driver uses (cc.mrlda.Document, alpha-N, beta-N, gamma-N files)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "host", "libmrlda_seqfile.so")
_lib = None
_dp, _ip, _lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
_fp, _bp = C.POINTER(C.c_float), C.POINTER(C.c_uint8)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} not found: run `make -C mrlda_b200/host`")
        L = C.CDLL(LIB_PATH)
        L.mrlda_io_last_error.restype = C.c_char_p
        L.mrlda_doc_serialize.restype = C.c_int64
        L.mrlda_doc_serialize.argtypes = [C.c_int32, _ip, _ip, C.c_int32, _dp, _bp, C.c_int64]
        L.mrlda_doc_deserialize.argtypes = [_bp, C.c_int64, _ip, _ip, C.c_int32, _ip, _dp, C.c_int32,
                                            _ip, _lp]
        L.mrlda_io_write_corpus.argtypes = [C.c_char_p, C.c_int64, _lp, _ip, _ip, _ip, _dp, C.c_int32]
        L.mrlda_io_write_gamma_parts.argtypes = [C.c_char_p, C.c_int32, C.c_int64, _lp, _ip, _ip, _ip, _dp,
                                                 C.c_int32, C.c_int64]
        L.mrlda_io_read_corpus.restype = C.c_void_p
        L.mrlda_io_read_corpus.argtypes = [C.c_char_p, C.c_int32, C.c_int32]
        L.mrlda_io_read_corpus_shard.restype = C.c_void_p
        L.mrlda_io_read_corpus_shard.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                                 C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        for f in ("mrlda_io_corpus_docs", "mrlda_io_corpus_nnz"):
            getattr(L, f).restype = C.c_int64
            getattr(L, f).argtypes = [C.c_void_p]
        L.mrlda_io_corpus_has_gamma.argtypes = [C.c_void_p]
        L.mrlda_io_corpus_copy.argtypes = [C.c_void_p, _lp, _ip, _ip, _ip, _dp]
        L.mrlda_io_corpus_copy.restype = None
        L.mrlda_io_corpus_free.argtypes = [C.c_void_p]
        L.mrlda_io_corpus_free.restype = None
        L.mrlda_io_write_alpha.argtypes = [C.c_char_p, C.c_int32, _dp]
        L.mrlda_io_read_alpha.argtypes = [C.c_char_p, C.c_int32, _dp]
        L.mrlda_io_write_beta.argtypes = [C.c_char_p, C.c_int32, C.c_int32, _dp, _fp, _bp]
        L.mrlda_io_read_beta.argtypes = [C.c_char_p, C.c_int32, C.c_int32, _dp, _fp, _bp]
        L.mrlda_io_java_double.argtypes = [C.c_double, C.c_char_p, C.c_int32]
        L.mrlda_io_write_term_index.argtypes = [C.c_char_p, C.c_int32, C.c_char_p]
        L.mrlda_io_write_eta.argtypes = [C.c_char_p, C.c_int32, C.c_int64, _ip, _ip]
        _lib = L
    return _lib


class SeqFileError(ValueError):
    pass


def _ck(rc):
    if rc is None or (isinstance(rc, int) and rc < 0):
        raise SeqFileError(load().mrlda_io_last_error().decode())
    return rc


def document_bytes(content, gamma=None):
    """Document.write: content = {termID: count} or None, gamma = sequence or None."""
    items = list(content.items()) if content else []
    t = np.array([k for k, _ in items], dtype=np.int32)
    c = np.array([v for _, v in items], dtype=np.int32)
    g = np.asarray(gamma if gamma is not None else [], dtype=np.float64)
    cap = 8 + 8 * len(items) + 8 * len(g)
    out = np.empty(cap, dtype=np.uint8)
    n = _ck(load().mrlda_doc_serialize(len(items), t.ctypes.data_as(_ip), c.ctypes.data_as(_ip), len(g),
                                       g.ctypes.data_as(_dp), out.ctypes.data_as(_bp), cap))
    return out[:n].tobytes()


def parse_document(data):
    """Document.readFields: returns (content dict or None, gamma array or None, numberOfTokens)."""
    buf = np.frombuffer(data, dtype=np.uint8).copy()
    cap = max(1, len(data) // 8 + 1)
    t, c = np.empty(cap, dtype=np.int32), np.empty(cap, dtype=np.int32)
    g = np.empty(cap, dtype=np.float64)
    n, k, tok = C.c_int32(0), C.c_int32(0), C.c_int64(0)
    _ck(load().mrlda_doc_deserialize(buf.ctypes.data_as(_bp), len(data), t.ctypes.data_as(_ip),
                                     c.ctypes.data_as(_ip), cap, C.byref(n), g.ctypes.data_as(_dp), cap,
                                     C.byref(k), C.byref(tok)))
    content = {int(a): int(b) for a, b in zip(t[:n.value], c[:n.value])} if n.value > 0 else None
    gamma = g[:k.value].copy() if k.value > 0 else None
    return content, gamma, int(tok.value)


def write_corpus(path, row_ptr, term_id, count, doc_id=None, gamma=None, n_topics=0):
    rp = np.ascontiguousarray(row_ptr, dtype=np.int64)
    t = np.ascontiguousarray(term_id, dtype=np.int32)
    c = np.ascontiguousarray(count, dtype=np.int32)
    did = np.ascontiguousarray(doc_id, dtype=np.int32) if doc_id is not None else None
    g = np.ascontiguousarray(gamma, dtype=np.float64) if gamma is not None else None
    _ck(load().mrlda_io_write_corpus(os.fsencode(path), len(rp) - 1, rp.ctypes.data_as(_lp),
                                     t.ctypes.data_as(_ip), c.ctypes.data_as(_ip),
                                     did.ctypes.data_as(_ip) if did is not None else None,
                                     g.ctypes.data_as(_dp) if g is not None else None, int(n_topics)))


def write_gamma_parts(directory, rank, row_ptr, term_id, count, gamma, doc_id=None, bytes_per_part=64 << 20):
    """One rank's gamma output as the driver writes it: gamma_gamma-m-NNNNN part files of about
    bytes_per_part bytes each, written concurrently.  Returns the number of part files."""
    rp = np.ascontiguousarray(row_ptr, dtype=np.int64)
    t = np.ascontiguousarray(term_id, dtype=np.int32)
    c = np.ascontiguousarray(count, dtype=np.int32)
    g = np.ascontiguousarray(gamma, dtype=np.float64)
    D = len(rp) - 1
    if g.ndim != 2 or g.shape[0] != D or t.size < rp[D] or c.size < rp[D]:
        raise ValueError("gamma must be D x K and term_id / count must cover row_ptr[D]")
    did = np.ascontiguousarray(doc_id, dtype=np.int32) if doc_id is not None else None
    if did is not None and did.size != D:
        raise ValueError("doc_id must hold D values")
    n = load().mrlda_io_write_gamma_parts(os.fsencode(directory), int(rank), D, rp.ctypes.data_as(_lp),
                                          t.ctypes.data_as(_ip), c.ctypes.data_as(_ip),
                                          did.ctypes.data_as(_ip) if did is not None else None,
                                          g.ctypes.data_as(_dp), int(g.shape[1]), int(bytes_per_part))
    _ck(n)
    return n


def read_corpus(path, n_topics, n_terms, rank=0, world=1):
    """Returns dict(row_ptr, term_id, count, doc_id, gamma or None); with world > 1 the input split of
    `rank` alone (contiguous documents balanced by entries), plus first_doc / total_docs."""
    L = load()
    first, total = C.c_int64(0), C.c_int64(0)
    if world > 1:
        h = L.mrlda_io_read_corpus_shard(os.fsencode(path), n_topics, n_terms, rank, world,
                                         C.byref(first), C.byref(total))
    else:
        h = L.mrlda_io_read_corpus(os.fsencode(path), n_topics, n_terms)
    if not h:
        raise SeqFileError(L.mrlda_io_last_error().decode())
    try:
        D, Z = L.mrlda_io_corpus_docs(h), L.mrlda_io_corpus_nnz(h)
        rp = np.empty(D + 1, dtype=np.int64)
        t, c = np.empty(max(Z, 1), dtype=np.int32), np.empty(max(Z, 1), dtype=np.int32)
        did = np.empty(max(D, 1), dtype=np.int32)
        has_g = bool(L.mrlda_io_corpus_has_gamma(h))
        g = np.empty((D, n_topics), dtype=np.float64) if has_g else None
        L.mrlda_io_corpus_copy(h, rp.ctypes.data_as(_lp), t.ctypes.data_as(_ip), c.ctypes.data_as(_ip),
                               did.ctypes.data_as(_ip), g.ctypes.data_as(_dp) if has_g else None)
    finally:
        L.mrlda_io_corpus_free(h)
    out = dict(row_ptr=rp, term_id=t[:Z], count=c[:Z], doc_id=did[:D], gamma=g)
    if world > 1:
        out.update(first_doc=first.value, total_docs=total.value)
    return out


def write_alpha(path, alpha):
    a = np.ascontiguousarray(alpha, dtype=np.float64)
    _ck(load().mrlda_io_write_alpha(os.fsencode(path), len(a), a.ctypes.data_as(_dp)))


def read_alpha(path, n_topics):
    a = np.empty(n_topics, dtype=np.float64)
    _ck(load().mrlda_io_read_alpha(os.fsencode(path), n_topics, a.ctypes.data_as(_dp)))
    return a


def write_beta(path, digamma_lambda, normaliser, present):
    dl = np.ascontiguousarray(digamma_lambda, dtype=np.float64)
    V, K = dl.shape
    nm = np.ascontiguousarray(normaliser, dtype=np.float32)
    pr = np.ascontiguousarray(present, dtype=np.uint8)
    _ck(load().mrlda_io_write_beta(os.fsencode(path), K, V, dl.ctypes.data_as(_dp),
                                   nm.ctypes.data_as(_fp), pr.ctypes.data_as(_bp)))


def read_beta(path, n_topics, n_terms):
    dl = np.empty((n_terms, n_topics), dtype=np.float64)
    nm = np.empty(n_topics, dtype=np.float32)
    pr = np.empty(n_terms, dtype=np.uint8)
    _ck(load().mrlda_io_read_beta(os.fsencode(path), n_topics, n_terms, dl.ctypes.data_as(_dp),
                                  nm.ctypes.data_as(_fp), pr.ctypes.data_as(_bp)))
    return dl, nm, pr


def java_double(x):
    buf = C.create_string_buffer(64)
    _ck(load().mrlda_io_java_double(float(x), buf, 64))
    return buf.value.decode()


def write_term_index(path, terms):
    """ParseCorpus's term index: SequenceFile<IntWritable, Text>, ids 1..len(terms)."""
    _ck(load().mrlda_io_write_term_index(os.fsencode(path), len(terms), "\n".join(terms).encode()))


def write_eta(path, n_topics, topics, terms):
    """The -informedprior file: SequenceFile<IntWritable topic, ArrayListOfIntsWritable seed terms>."""
    t = np.ascontiguousarray(topics, dtype=np.int32)
    v = np.ascontiguousarray(terms, dtype=np.int32)
    _ck(load().mrlda_io_write_eta(os.fsencode(path), n_topics, len(t), t.ctypes.data_as(_ip),
                                  v.ctypes.data_as(_ip)))
