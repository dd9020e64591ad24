"""The C++ codecs (mrlda_b200/host/, through libmrlda_seqfile.so) against an independent plain-Python
implementation of the same published layouts (tests/pyseq.py): files written by one side are read
by the other, for random corpora, model files and sync-marker placements."""
import struct

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import pyseq
from mrlda_b200 import seqfile

INT = "org.apache.hadoop.io.IntWritable"
DBL = "org.apache.hadoop.io.DoubleWritable"
DOC = "cc.mrlda.Document"
PIF = "edu.umd.cloud9.io.pair.PairOfIntFloat"
HMAP = "edu.umd.cloud9.io.map.HMapIDW"


@settings(deadline=None)
@given(st.integers(min_value=-(2 ** 63), max_value=2 ** 63 - 1))
def test_vint_round_trip(i):
    b = pyseq.write_vint(i)
    v, pos = pyseq.read_vint(b, 0)
    assert v == i and pos == len(b) and 1 <= len(b) <= 9


def test_vint_known_encodings():
    # WritableUtils.writeVLong: one byte in [-112, 127]; otherwise a length/sign byte and big-endian bytes
    assert pyseq.write_vint(0) == b"\x00" and pyseq.write_vint(127) == b"\x7f"
    assert pyseq.write_vint(-112) == b"\x90"
    assert pyseq.write_vint(128) == b"\x8f\x80"
    assert pyseq.write_vint(255) == b"\x8f\xff"
    assert pyseq.write_vint(256) == b"\x8e\x01\x00"
    assert pyseq.write_vint(-113) == b"\x87\x70"


docs_strategy = st.lists(
    st.tuples(
        st.integers(min_value=-5, max_value=2 ** 31 - 1),  # document id
        st.dictionaries(st.integers(min_value=1, max_value=500), st.integers(min_value=1, max_value=10 ** 6),
                        min_size=1, max_size=60),
    ), min_size=1, max_size=40)


@settings(max_examples=25, deadline=None)
@given(docs_strategy, st.integers(min_value=0, max_value=5), st.integers(min_value=0, max_value=7))
def test_corpus_written_by_cpp_is_read_by_python(tmp_path_factory, docs, K, seed):
    path = str(tmp_path_factory.mktemp("c") / "part-00000")
    rng = np.random.default_rng(seed)
    row_ptr = np.cumsum([0] + [len(c) for _, c in docs])
    term = np.concatenate([np.fromiter(c.keys(), dtype=np.int32) for _, c in docs])
    cnt = np.concatenate([np.fromiter(c.values(), dtype=np.int32) for _, c in docs])
    ids = np.array([d for d, _ in docs], dtype=np.int32)
    gamma = rng.gamma(1.0, size=(len(docs), K)) if K else None
    seqfile.write_corpus(path, row_ptr, term, cnt, doc_id=ids, gamma=gamma, n_topics=K)
    kc, vc, _, records, escapes = pyseq.read_file(path)
    assert (kc, vc) == (INT, DOC) and len(records) == len(docs)
    total = sum(8 + len(k) + len(v) for k, v in records)
    assert escapes >= total // 4000  # a sync escape at least every few thousand bytes (SYNC_INTERVAL 2000)
    for i, (k, v) in enumerate(records):
        assert struct.unpack(">i", k)[0] == ids[i]
        content, g = pyseq.parse_document(v)
        assert content == docs[i][1]
        if K:
            assert np.array_equal(np.array(g), gamma[i])
        else:
            assert g is None


@settings(max_examples=25, deadline=None)
@given(docs_strategy, st.integers(min_value=0, max_value=4), st.sampled_from([None, 1, 2, 7]))
def test_corpus_written_by_python_is_read_by_cpp(tmp_path_factory, docs, K, sync_every):
    path = str(tmp_path_factory.mktemp("p") / "part-00000")
    rng = np.random.default_rng(len(docs))
    gamma = rng.gamma(1.0, size=(len(docs), K)) if K else None
    records = [(struct.pack(">i", d), pyseq.document_bytes(c, gamma[i] if K else None))
               for i, (d, c) in enumerate(docs)]
    pyseq.write_file(path, INT, DOC, records, sync=bytes(rng.integers(0, 256, 16, dtype=np.uint8)),
                     sync_every=sync_every)
    got = seqfile.read_corpus(path, max(K, 1), 500)
    assert np.array_equal(got["doc_id"], [d for d, _ in docs])
    assert np.array_equal(np.diff(got["row_ptr"]), [len(c) for _, c in docs])
    for i, (_, c) in enumerate(docs):
        lo, hi = got["row_ptr"][i], got["row_ptr"][i + 1]
        assert dict(zip(got["term_id"][lo:hi].tolist(), got["count"][lo:hi].tolist())) == c
    if K:
        assert np.array_equal(got["gamma"], gamma)
    else:
        assert got["gamma"] is None


@settings(max_examples=25, deadline=None)
@given(docs_strategy, st.integers(min_value=0, max_value=3), st.booleans(),
       st.sampled_from([pyseq.DEFAULT_CODEC, pyseq.GZIP_CODEC]), st.sampled_from([1, 3, 1000]))
def test_compressed_corpus_is_read_by_cpp(tmp_path_factory, docs, K, block, codec, per_block):
    """Record- and block-compressed SequenceFiles (zlib / gzip codec streams), as SequenceFile.Reader
    accepts them for the reference (VariationalInference.java:282, DocumentMapper.java:475-536)."""
    path = str(tmp_path_factory.mktemp("z") / "part-00000")
    rng = np.random.default_rng(len(docs) + 5)
    gamma = rng.gamma(1.0, size=(len(docs), K)) if K else None
    records = [(struct.pack(">i", d), pyseq.document_bytes(c, gamma[i] if K else None))
               for i, (d, c) in enumerate(docs)]
    pyseq.write_compressed_file(path, INT, DOC, records, block, codec=codec, per_block=per_block,
                                sync=bytes(rng.integers(0, 256, 16, dtype=np.uint8)), sync_every=2)
    got = seqfile.read_corpus(path, max(K, 1), 500)
    assert np.array_equal(got["doc_id"], [d for d, _ in docs])
    assert np.array_equal(np.diff(got["row_ptr"]), [len(c) for _, c in docs])
    for i, (_, c) in enumerate(docs):
        lo, hi = got["row_ptr"][i], got["row_ptr"][i + 1]
        assert dict(zip(got["term_id"][lo:hi].tolist(), got["count"][lo:hi].tolist())) == c
    if K:
        assert np.array_equal(got["gamma"], gamma)


def test_input_splits_cover_the_corpus(tmp_path):
    """mrlda-vi ranks read only their own split (read_corpus_shard): the splits are contiguous, balanced
    by entries like corpus.shard_bounds, and together they are the corpus; part files are read in
    name order (one of them block-compressed)."""
    from mrlda_b200 import corpus
    rng = np.random.default_rng(11)
    lens = rng.integers(0, 40, 300)
    docs = [(1000 + i, {int(t): int(rng.integers(1, 5)) for t in rng.choice(400, n, replace=False) + 1})
            for i, n in enumerate(lens)]
    recs = [(struct.pack(">i", d), pyseq.document_bytes(c, None)) for d, c in docs]
    d = tmp_path / "in"
    d.mkdir()
    pyseq.write_file(str(d / "part-00000"), INT, DOC, recs[:120], sync_every=7)
    pyseq.write_compressed_file(str(d / "part-00001"), INT, DOC, recs[120:], True, per_block=16)
    (d / "_SUCCESS").write_bytes(b"")
    whole = seqfile.read_corpus(str(d), 1, 500)
    assert whole["doc_id"].tolist() == [i for i, _ in docs]
    for world in (2, 3, 8):
        bounds = corpus.shard_bounds(whole["row_ptr"], world)
        ids, nnz = [], 0
        for rank in range(world):
            part = seqfile.read_corpus(str(d), 1, 500, rank=rank, world=world)
            assert part["total_docs"] == 300 and part["first_doc"] == bounds[rank]
            assert len(part["doc_id"]) == bounds[rank + 1] - bounds[rank]
            lo = whole["row_ptr"][bounds[rank]]
            assert np.array_equal(part["term_id"], whole["term_id"][lo:lo + len(part["term_id"])])
            assert np.array_equal(part["count"], whole["count"][lo:lo + len(part["count"])])
            ids += part["doc_id"].tolist()
            nnz += len(part["term_id"])
        assert ids == whole["doc_id"].tolist() and nnz == len(whole["term_id"])


def test_compressed_model_files_and_unknown_codecs(tmp_path):
    """alpha files through both compressed layouts; a codec that needs Hadoop's native library is
    refused by name; damaged codec streams raise."""
    alpha = np.array([0.5, 1.25, 3.0])
    recs = [(struct.pack(">i", i + 1), struct.pack(">d", a)) for i, a in enumerate(alpha)]
    for block in (False, True):
        p = str(tmp_path / f"alpha-{int(block)}")
        pyseq.write_compressed_file(p, INT, DBL, recs, block)
        assert np.array_equal(seqfile.read_alpha(p, 3), alpha)
    p = str(tmp_path / "snappy")
    pyseq.write_compressed_file(p, INT, DBL, recs, False, codec="org.apache.hadoop.io.compress.SnappyCodec")
    with pytest.raises(seqfile.SeqFileError, match="SnappyCodec"):
        seqfile.read_alpha(p, 3)
    p = str(tmp_path / "damaged")
    pyseq.write_compressed_file(p, INT, DBL, recs, True)
    raw = bytearray(open(p, "rb").read())
    raw[-6] ^= 0xFF
    open(p, "wb").write(bytes(raw))
    with pytest.raises(seqfile.SeqFileError):
        seqfile.read_alpha(p, 3)
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-9])  # truncated block
    with pytest.raises(seqfile.SeqFileError):
        seqfile.read_alpha(p, 3)


def test_alpha_and_beta_files_cross_check(tmp_path):
    rng = np.random.default_rng(3)
    K, V = 7, 23
    alpha = rng.gamma(1.0, size=K)
    seqfile.write_alpha(str(tmp_path / "alpha"), alpha)
    kc, vc, _, records, _ = pyseq.read_file(str(tmp_path / "alpha"))
    assert (kc, vc) == (INT, DBL)
    assert [struct.unpack(">i", k)[0] for k, _ in records] == list(range(1, K + 1))  # VI.java:549-558
    assert np.array_equal([struct.unpack(">d", v)[0] for _, v in records], alpha)
    # and the other way round, in a shuffled record order (importAlpha indexes by key, VI.java:521-540)
    order = rng.permutation(K)
    pyseq.write_file(str(tmp_path / "alpha2"), INT, DBL,
                     [(struct.pack(">i", int(i) + 1), struct.pack(">d", alpha[i])) for i in order])
    assert np.array_equal(seqfile.read_alpha(str(tmp_path / "alpha2"), K), alpha)

    dl = rng.normal(size=(V, K))
    nm = rng.normal(size=K).astype(np.float32)
    present = (rng.random(V) < 0.7).astype(np.uint8)
    seqfile.write_beta(str(tmp_path / "beta"), dl, nm, present)
    kc, vc, _, records, _ = pyseq.read_file(str(tmp_path / "beta"))
    assert (kc, vc) == (PIF, HMAP) and len(records) == K  # one record per topic (TermReducer.java:232-235)
    seen = np.flatnonzero(present) + 1
    for k, (key, val) in enumerate(records):
        topic, norm = struct.unpack(">if", key)
        assert topic == k + 1 and np.float32(norm) == nm[k]
        (n,) = struct.unpack_from(">i", val, 0)
        assert n == len(seen) and len(val) == 4 + 12 * n
        for j in range(n):
            term, x = struct.unpack_from(">id", val, 4 + 12 * j)
            assert term == seen[j] and x == dl[term - 1, k]
    # python-written beta with the terms of each topic in a different order
    recs = []
    for k in range(K):
        terms = rng.permutation(seen)
        val = struct.pack(">i", len(terms)) + b"".join(struct.pack(">id", int(t), dl[t - 1, k]) for t in terms)
        recs.append((struct.pack(">if", k + 1, float(nm[k])), val))
    pyseq.write_file(str(tmp_path / "beta2"), PIF, HMAP, recs, sync_every=3)
    dl2, nm2, pr2 = seqfile.read_beta(str(tmp_path / "beta2"), K, V)
    assert np.array_equal(pr2, present) and np.array_equal(nm2, nm)
    assert np.array_equal(dl2[present.astype(bool)], dl[present.astype(bool)])


def test_cpp_reader_rejects_damaged_files(tmp_path):
    good = [(struct.pack(">i", 1), pyseq.document_bytes({1: 2}, None))]
    p = str(tmp_path / "x")
    pyseq.write_file(p, INT, DOC, good, version=7)  # a version this reader does not know
    with pytest.raises(seqfile.SeqFileError):
        seqfile.read_corpus(p, 1, 10)
    pyseq.write_file(p, INT, DOC, good, version=5)  # version 5: same header without the metadata block
    assert seqfile.read_corpus(p, 1, 10)["count"].tolist() == [2]
    pyseq.write_file(p, INT, DBL, good)  # wrong value class
    with pytest.raises(seqfile.SeqFileError):
        seqfile.read_corpus(p, 1, 10)
    pyseq.write_file(p, INT, DOC, good)
    open(p, "ab").write(b"\x00\x00\x00\x20\x00\x00")  # truncated record
    with pytest.raises(seqfile.SeqFileError):
        seqfile.read_corpus(p, 1, 10)


def test_term_index_and_eta_files_cross_check(tmp_path):
    terms = ["alpha", "bêta", "γάμμα", "x" * 300]  # a Text longer than 127 bytes needs a multi-byte vint
    seqfile.write_term_index(str(tmp_path / "term"), terms)
    kc, vc, _, records, _ = pyseq.read_file(str(tmp_path / "term"))
    assert (kc, vc) == (INT, "org.apache.hadoop.io.Text")
    for i, (k, v) in enumerate(records):
        assert struct.unpack(">i", k)[0] == i + 1  # ids 1..V in index order (ParseCorpus.java:484-487)
        n, pos = pyseq.read_vint(v, 0)
        assert v[pos:pos + n].decode("utf-8") == terms[i] and pos + n == len(v)

    seqfile.write_eta(str(tmp_path / "eta"), 3, [1, 1, 3], [5, 9, 2])
    kc, vc, _, records, _ = pyseq.read_file(str(tmp_path / "eta"))
    assert (kc, vc) == (INT, "edu.umd.cloud9.io.array.ArrayListOfIntsWritable")
    got = {}
    for k, v in records:
        (n,) = struct.unpack_from(">i", v, 0)
        assert len(v) == 4 + 4 * n
        got[struct.unpack(">i", k)[0]] = list(struct.unpack_from(">%di" % n, v, 4))
    assert got == {1: [5, 9], 2: [], 3: [2]}  # one record per topic (InformedPrior.java:139-170)


def _raw_document(entries, gamma=None):
    """Document bytes from an entry LIST (so that a term id can repeat, which a dict cannot express)."""
    out = struct.pack(">i", len(entries))
    for t, c in entries:
        out += struct.pack(">ii", t, c)
    g = list(gamma) if gamma is not None else []
    out += struct.pack(">i", len(g)) + b"".join(struct.pack(">d", x) for x in g)
    return out


def test_repeated_term_ids_close_the_gap_between_part_files(tmp_path, monkeypatch):
    """HMapII.put overwrites: a term id repeated inside one record keeps its first position and takes
    the last count.  The two-pass reader sizes the arrays from the scanned entry counts, so such
    records leave their part file short and the following files have to move down — with one thread
    and with several."""
    rng = np.random.default_rng(5)
    K = 3
    files, want_docs = [], []
    for f in range(4):
        recs = []
        for i in range(25):
            n = int(rng.integers(0, 12))
            ids = (rng.choice(60, n, replace=False) + 1).tolist()
            entries = [(t, int(rng.integers(1, 9))) for t in ids]
            if n >= 2 and (f in (0, 2)) and i % 3 == 0:  # repeat two ids with new counts
                entries += [(ids[0], 77), (ids[1], 88), (ids[0], 99)]
            merged = {}
            for t, c in entries:
                merged[t] = c  # dict keeps first position, last value: HMapII.put
            g = rng.random(K).tolist()
            doc_id = 100 * f + i
            recs.append((struct.pack(">i", doc_id), _raw_document(entries, g)))
            want_docs.append((doc_id, list(merged.items()), g))
        files.append(recs)
    d = tmp_path / "in"
    d.mkdir()
    for f, recs in enumerate(files):
        pyseq.write_file(str(d / f"part-{f:05d}"), INT, DOC, recs, sync_every=5)
    for threads in ("1", "4"):
        monkeypatch.setenv("MRLDA_IO_THREADS", threads)
        got = seqfile.read_corpus(str(d), K, 60)
        assert got["doc_id"].tolist() == [x[0] for x in want_docs]
        for i, (_, items, g) in enumerate(want_docs):
            lo, hi = got["row_ptr"][i], got["row_ptr"][i + 1]
            assert list(zip(got["term_id"][lo:hi].tolist(), got["count"][lo:hi].tolist())) == items
            assert got["gamma"][i].tolist() == g
        assert got["row_ptr"][-1] == len(got["term_id"]) == sum(len(x[1]) for x in want_docs)
        # the splits of a sharded read see the same documents (bounds come from the scanned counts)
        ids = []
        for rank in range(3):
            part = seqfile.read_corpus(str(d), K, 60, rank=rank, world=3)
            lo = part["first_doc"]
            for j in range(len(part["doc_id"])):
                a, b = part["row_ptr"][j], part["row_ptr"][j + 1]
                assert list(zip(part["term_id"][a:b].tolist(), part["count"][a:b].tolist())) == want_docs[lo + j][1]
            assert part["gamma"] is not None and part["gamma"].shape == (len(part["doc_id"]), K)
            ids += part["doc_id"].tolist()
        assert ids == [x[0] for x in want_docs]


def test_gamma_is_kept_only_when_every_kept_document_has_k_values(tmp_path):
    """DocumentMapper.java:184: a stored gamma of another length is ignored; the CSR carries gamma only
    when every document of the requested range has one of n_topics values."""
    K = 4
    a = [(struct.pack(">i", i), pyseq.document_bytes({1 + i: 2}, [0.5 + i] * K)) for i in range(6)]
    b = [(struct.pack(">i", 10 + i), pyseq.document_bytes({3: 1, 4 + i: 1}, [1.0] * (K if i else K + 1))) for i in range(6)]
    d = tmp_path / "in"
    d.mkdir()
    pyseq.write_file(str(d / "part-00000"), INT, DOC, a)
    pyseq.write_file(str(d / "part-00001"), INT, DOC, b)
    assert seqfile.read_corpus(str(d / "part-00000"), K, 20)["gamma"].shape == (6, K)
    assert seqfile.read_corpus(str(d), K, 20)["gamma"] is None  # one record of part-00001 has K + 1 values
    assert seqfile.read_corpus(str(d), K + 1, 20)["gamma"] is None
    first = seqfile.read_corpus(str(d), K, 20, rank=0, world=3)  # a third of the entries: part-00000 alone
    assert first["doc_id"].tolist() == list(range(6)) and first["gamma"].shape == (6, K)
    with pytest.raises(seqfile.SeqFileError, match="term id"):
        seqfile.read_corpus(str(d), K, 8)  # term ids above n_terms are refused in the second pass too


@pytest.mark.parametrize("bytes_per_part", [1 << 30, 4096, 1])
def test_gamma_part_files_round_trip(tmp_path, bytes_per_part):
    """write_gamma_parts: the rank's documents as concurrently written gamma_gamma-m-NNNNN files,
    rank r owning the indices from 32 r; read back in name order they are the documents (those with
    null content dropped, DocumentMapper.java:197-201) with the gamma rows embedded."""
    from mrlda_b200 import corpus
    K = 5
    c = corpus.generate(400, 300, K, 30, seed=9)
    rp = c.row_ptr.copy()
    # make documents 3 and 200 empty (null content): they must not be emitted
    keep = np.ones(c.nnz, bool)
    for dd in (3, 200):
        keep[rp[dd]:rp[dd + 1]] = False
    n = np.diff(rp)
    n[[3, 200]] = 0
    rp2 = np.concatenate([[0], np.cumsum(n)])
    term, cnt = c.term_id[keep], c.count[keep]
    gamma = np.random.default_rng(2).random((400, K))
    ids = np.arange(400, dtype=np.int32) * 3 + 7
    out = tmp_path / "gamma-1"
    out.mkdir()
    half = 180
    sl = lambda a, b: (rp2[a:b + 1] - rp2[a], term[rp2[a]:rp2[b]], cnt[rp2[a]:rp2[b]])  # noqa: E731
    n0 = seqfile.write_gamma_parts(str(out), 0, *sl(0, half), gamma[:half], doc_id=ids[:half], bytes_per_part=bytes_per_part)
    n1 = seqfile.write_gamma_parts(str(out), 1, *sl(half, 400), gamma[half:], doc_id=ids[half:], bytes_per_part=bytes_per_part)
    names = sorted(p.name for p in out.iterdir())
    assert names == [f"gamma_gamma-m-{i:05d}" for i in range(n0)] + [f"gamma_gamma-m-{32 + i:05d}" for i in range(n1)]
    assert (n0 == n1 == 1) if bytes_per_part == 1 << 30 else (2 <= n0 <= 32 and 2 <= n1 <= 32)
    got = seqfile.read_corpus(str(out), K, 300)
    kept = np.array([i for i in range(400) if i not in (3, 200)])
    assert got["doc_id"].tolist() == ids[kept].tolist()
    assert np.array_equal(got["gamma"], gamma[kept])
    assert np.array_equal(got["term_id"], term) and np.array_equal(got["count"], cnt)
    assert np.array_equal(np.diff(got["row_ptr"]), n[kept])
    # and the independent Python reader agrees, file by file
    seen = []
    for name in names:
        kc, vc, _, recs, _ = pyseq.read_file(str(out / name))
        assert (kc, vc) == (INT, DOC)
        for kb, vb in recs:
            (doc_id,) = struct.unpack(">i", kb)
            content, g = pyseq.parse_document(vb)[:2]
            i = (doc_id - 7) // 3
            assert content == dict(zip(term[rp2[i]:rp2[i + 1]].tolist(), cnt[rp2[i]:rp2[i + 1]].tolist()))
            assert g == gamma[i].tolist()
            seen.append(doc_id)
    assert seen == ids[kept].tolist()


def test_a_failed_flush_is_an_error_not_a_short_file():
    """The writers buffer 1 MB; what the final flush cannot write (disk full) must surface."""
    import os
    if not os.path.exists("/dev/full"):
        pytest.skip("no /dev/full")
    with pytest.raises(seqfile.SeqFileError, match="short write"):
        seqfile.write_corpus("/dev/full", [0, 1, 2], [1, 2], [3, 1])
    with pytest.raises(seqfile.SeqFileError, match="short write"):
        seqfile.write_alpha("/dev/full", [0.1, 0.2])


@pytest.mark.parametrize("K,V,frac", [(3, 10, 0.5), (8, 1, 1.0), (17, 40, 1.0), (64, 170, 0.9), (203, 900, 0.7)])
def test_concurrent_beta_writer_is_byte_identical_to_appending(tmp_path, monkeypatch, K, V, frac):
    """write_beta plans every record's position (sync escapes included) and lets the tiles store their
    own bytes; the file must be exactly what appending the records one by one produces, for records
    smaller and larger than the 2000-byte sync interval, and must read back through both readers."""
    rng = np.random.default_rng(K)
    dl = rng.standard_normal((V, K))
    nm = rng.random(K).astype(np.float32)
    pr = (rng.random(V) < frac).astype(np.uint8)
    pr[0] = 1
    files = {}
    for threads in ("1", "5"):
        monkeypatch.setenv("MRLDA_IO_THREADS", threads)
        # the same path both times: the sync marker is derived from it
        path = tmp_path / "beta-1"
        seqfile.write_beta(str(path), dl, nm, pr)
        files[threads] = path.read_bytes()
    assert files["1"] == files["5"]
    got = seqfile.read_beta(str(tmp_path / "beta-1"), K, V)
    m = pr.astype(bool)
    assert np.array_equal(got[0][m], dl[m]) and np.array_equal(got[1], nm) and np.array_equal(got[2], pr)
    kc, vc, _, recs, escapes = pyseq.read_file(str(tmp_path / "beta-1"))
    assert (kc, vc) == (PIF, HMAP) and len(recs) == K
    for k, (kb, vb) in enumerate(recs):
        topic, norm = struct.unpack(">if", kb)
        assert topic == k + 1 and np.float32(norm) == nm[k]
        (n,) = struct.unpack_from(">i", vb, 0)
        assert n == int(m.sum()) and len(vb) == 4 + 12 * n


def _hmap(entries):
    return struct.pack(">i", len(entries)) + b"".join(struct.pack(">id", t, x) for t, x in entries)


@pytest.mark.parametrize("threads", ["1", "4"])
def test_beta_reader_takes_any_record_and_term_order(tmp_path, monkeypatch, threads):
    """importBeta accepts the topics in any order, each with its own term order and its own subset of
    terms (HMapIDW iteration order is arbitrary); the tiled reader must too, with one thread and with
    several, and report a cell that is initialised twice (DocumentMapper.java:514-518)."""
    monkeypatch.setenv("MRLDA_IO_THREADS", threads)
    K, V = 21, 57
    rng = np.random.default_rng(8)
    want = np.zeros((V, K))
    nm = rng.random(K).astype(np.float32)
    recs = []
    for k in rng.permutation(K):
        terms = rng.permutation(V)[: int(rng.integers(1, V + 1))] + 1
        vals = rng.standard_normal(len(terms))
        want[terms - 1, k] = vals
        recs.append((struct.pack(">if", int(k) + 1, float(nm[k])), _hmap(list(zip(terms.tolist(), vals.tolist())))))
    recs.insert(5, (struct.pack(">if", 3, 0.0), _hmap([])))  # an empty map for a topic is legal
    nm_want = nm.copy()
    path = tmp_path / "beta-7"
    pyseq.write_file(str(path), PIF, HMAP, recs, sync_every=4)
    dl, got_nm, present = seqfile.read_beta(str(path), K, V)
    assert np.array_equal(dl, want)
    topics_in_order = [struct.unpack(">if", kb) for kb, _ in recs]
    for t, x in topics_in_order:
        nm_want[t - 1] = np.float32(x)  # last record of a topic wins
    assert np.array_equal(got_nm, nm_want)
    assert np.array_equal(present, (want != 0).any(1).astype(np.uint8))
    # the same cell in two records of one topic
    bad = recs + [(struct.pack(">if", 2, 1.0), _hmap([(int(np.flatnonzero(want[:, 1])[0]) + 1, 9.0)]))]
    pyseq.write_file(str(tmp_path / "beta-8"), PIF, HMAP, bad)
    with pytest.raises(seqfile.SeqFileError, match="Dual initialization"):
        seqfile.read_beta(str(tmp_path / "beta-8"), K, V)
    # ... and inside one record
    pyseq.write_file(str(tmp_path / "beta-9"), PIF, HMAP, [(struct.pack(">if", 1, 1.0), _hmap([(4, 1.0), (5, 2.0), (4, 3.0)]))])
    with pytest.raises(seqfile.SeqFileError, match="Dual initialization"):
        seqfile.read_beta(str(tmp_path / "beta-9"), K, V)
    pyseq.write_file(str(tmp_path / "beta-10"), PIF, HMAP, [(struct.pack(">if", 1, 1.0), _hmap([(V + 1, 1.0)]))])
    with pytest.raises(seqfile.SeqFileError, match="out of range"):
        seqfile.read_beta(str(tmp_path / "beta-10"), K, V)
    pyseq.write_file(str(tmp_path / "beta-11"), PIF, HMAP, [(struct.pack(">if", K + 1, 1.0), _hmap([(1, 1.0)]))])
    with pytest.raises(seqfile.SeqFileError, match="Invalid beta vector"):
        seqfile.read_beta(str(tmp_path / "beta-11"), K, V)
