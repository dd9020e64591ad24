"""Wire formats either side of the hot path: cc.mrlda.Document (golden vectors from the
reference's DocumentTest.java), Hadoop SequenceFile container, alpha / beta / gamma model files.
Runs without a GPU."""
import json
import os
import struct

import numpy as np
import pytest

from mrlda_b200 import corpus, seqfile

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "document_vectors.json")))


@pytest.mark.parametrize("name", sorted(GOLD))
def test_document_golden_vectors(name):
    g = GOLD[name]
    content = {t: c for t, c in g["content"]} if g["content"] else None
    data = seqfile.document_bytes(content, g["gamma"])
    assert data.hex() == g["hex"]
    c2, g2, tokens = seqfile.parse_document(bytes.fromhex(g["hex"]))
    assert c2 == content
    assert tokens == g["tokens"] and (len(c2) if c2 else 0) == g["types"]  # DocumentTest: 37 / 3
    if g["gamma"] is None:
        assert g2 is None
    else:
        assert np.max(np.abs(g2 - np.array(g["gamma"]))) < 1e-12  # DocumentTest.PRECISION


def test_document_entry_order_agnostic_and_nonpositive_counts():
    a = seqfile.parse_document(struct.pack(">iiiiiiii", 3, 3, 10, 1, 22, 2, 5, 0))
    assert a[0] == {1: 22, 2: 5, 3: 10} and a[1] is None and a[2] == 37
    # numEntries <= 0 => content null; numTopics <= 0 => gamma null (Document.java:151,164)
    assert seqfile.parse_document(struct.pack(">ii", -5, -1)) == (None, None, 0)
    with pytest.raises(seqfile.SeqFileError):
        seqfile.parse_document(struct.pack(">ii", 2, 1))  # truncated


def test_corpus_sequencefile_round_trip(tmp_path):
    c = corpus.generate(300, 500, 7, 60, seed=4)  # > 2000 bytes: exercises the sync escapes
    rng = np.random.default_rng(0)
    gamma = rng.random((c.n_docs, 7)) + 0.1
    ids = np.arange(c.n_docs, dtype=np.int32) * 3 + 5
    p = str(tmp_path / "part-00000")
    seqfile.write_corpus(p, c.row_ptr, c.term_id, c.count, doc_id=ids, gamma=gamma, n_topics=7)
    raw = open(p, "rb").read()
    assert raw[:4] == b"SEQ\x06"
    assert b"org.apache.hadoop.io.IntWritable" in raw[:80] and b"cc.mrlda.Document" in raw[:80]
    assert raw.count(struct.pack(">i", -1) + raw[raw.index(b"cc.mrlda.Document") + 23:][:16]) >= 1
    r = seqfile.read_corpus(p, 7, 500)
    assert np.array_equal(r["row_ptr"], c.row_ptr)
    assert np.array_equal(r["term_id"], c.term_id) and np.array_equal(r["count"], c.count)
    assert np.array_equal(r["doc_id"], ids) and np.array_equal(r["gamma"], gamma)
    # a stored gamma of another length is ignored (DocumentMapper.java:184)
    assert seqfile.read_corpus(p, 9, 500)["gamma"] is None
    with pytest.raises(seqfile.SeqFileError):
        seqfile.read_corpus(p, 7, 10)  # term ids beyond -term


def test_directory_input_skips_hidden_files_and_keeps_name_order(tmp_path):
    a = corpus.generate(5, 50, 3, 10, seed=1)
    b = corpus.generate(4, 50, 3, 10, seed=2)
    d = tmp_path / "in"
    d.mkdir()
    seqfile.write_corpus(str(d / "part-00001"), b.row_ptr, b.term_id, b.count)
    seqfile.write_corpus(str(d / "part-00000"), a.row_ptr, a.term_id, a.count)
    (d / "_SUCCESS").write_bytes(b"")
    (d / ".part-00000.crc").write_bytes(b"junk")
    r = seqfile.read_corpus(str(d), 3, 50)
    assert len(r["row_ptr"]) - 1 == 9
    assert np.array_equal(r["term_id"], np.concatenate([a.term_id, b.term_id]))
    empty = tmp_path / "empty"
    empty.mkdir()
    assert len(seqfile.read_corpus(str(empty), 3, 50)["row_ptr"]) == 1  # no documents


def test_alpha_file(tmp_path):
    p = str(tmp_path / "alpha-0")
    a = np.array([1e-3, 0.5, 7.25])
    seqfile.write_alpha(p, a)
    raw = open(p, "rb").read()
    assert b"org.apache.hadoop.io.DoubleWritable" in raw
    # records: recLen=12, keyLen=4, int32 k (1-based), float64 (VariationalInference.java:549-558)
    assert struct.pack(">iiid", 12, 4, 2, 0.5) in raw
    assert np.array_equal(seqfile.read_alpha(p, 3), a)
    with pytest.raises(seqfile.SeqFileError):
        seqfile.read_alpha(p, 4)  # importAlpha: exactly K records
    with pytest.raises(seqfile.SeqFileError):
        seqfile.read_alpha(p, 2)  # index out of (0, K]


def test_beta_file(tmp_path):
    K, V = 4, 30
    rng = np.random.default_rng(3)
    dl = -rng.random((V, K)) * 20
    nm = rng.random(K).astype(np.float32)
    pr = (rng.random(V) < 0.7).astype(np.uint8)
    p = str(tmp_path / "beta-1")
    seqfile.write_beta(p, dl, nm, pr)
    raw = open(p, "rb").read()
    assert b"edu.umd.cloud9.io.pair.PairOfIntFloat" in raw and b"edu.umd.cloud9.io.map.HMapIDW" in raw
    dl2, nm2, pr2 = seqfile.read_beta(p, K, V)
    assert np.array_equal(pr2, pr) and np.array_equal(nm2, nm)
    assert np.array_equal(dl2[pr.astype(bool)], dl[pr.astype(bool)])
    assert np.all(dl2[~pr.astype(bool)] == 0)
    with pytest.raises(seqfile.SeqFileError):
        seqfile.read_beta(p, K - 1, V)  # importBeta: topic must be in [1, K]


def test_java_double_to_string():
    for x, want in [(-147054303.548207, "-1.47054303548207E8"), (-3224.313308, "-3224.313308"),
                    (0.001, "0.001"), (1e-4, "1.0E-4"), (1.0, "1.0"), (12345678.0, "1.2345678E7"),
                    (100.0, "100.0"), (0.5, "0.5"), (-0.0, "-0.0"), (2.5e-7, "2.5E-7")]:
        assert seqfile.java_double(x) == want


def test_display_tools_read_model_files(tmp_path):
    """DisplayTopic / DisplayDocument counterparts read what the codecs write
    (DisplayTopic.java:97-142, DisplayDocument.java:87-100)."""
    import subprocess
    tool = os.path.join(os.path.dirname(os.path.dirname(__file__)), "mrlda_b200", "host", "mrlda-display")
    K, V = 2, 5
    dl = np.array([[-1.0, -9.0], [-2.0, -8.0], [-3.0, -7.5], [-4.0, -6.0], [-5.0, -5.5]])
    beta, index = str(tmp_path / "beta-1"), str(tmp_path / "term")
    seqfile.write_beta(beta, dl, np.zeros(K, dtype=np.float32), np.ones(V, dtype=np.uint8))
    seqfile.write_term_index(index, ["alpha", "beta", "gamma", "delta", "eps"])
    out = subprocess.run([tool, "topic", "-input", beta, "-index", index, "-topdisplay", "2"],
                         capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[1] == "Top ranked 2 terms for Topic 1" and out[3:5] == ["alpha\t\t-1.0", "beta\t\t-2.0"]
    assert out[6] == "Top ranked 2 terms for Topic 2" and out[8:10] == ["eps\t\t-5.5", "delta\t\t-6.0"]
    gdir = tmp_path / "gamma-1"
    gdir.mkdir()
    seqfile.write_corpus(str(gdir / "gamma_gamma-m-00000"), [0, 1, 2], [1, 2], [3, 1], doc_id=[7, 9],
                         gamma=np.array([[0.5, 1.25], [100.0, 1e-3]]), n_topics=2)
    out = subprocess.run([tool, "document", "-input", str(gdir)], capture_output=True, text=True,
                         check=True).stdout.splitlines()
    assert out == ["7 0.5 1.25 ", "9 100.0 0.001 "]
