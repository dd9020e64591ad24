"""A second, independent implementation of the wire formats either side of the path, written in plain
Python from the published layouts — test infrastructure only, used to cross-check the C++ codecs of
mrlda_b200/host/ from the outside:

  * Hadoop SequenceFile version 6, uncompressed, record-oriented: "SEQ\\x06", key and value class
    names (Text.writeString: vint length + UTF-8), two compression flags, metadata count, 16-byte sync
    marker; records `int32 recordLength, int32 keyLength, key, value`; a sync escape `int32 -1` +
    marker wherever the writer chose to put one.
  * WritableUtils vint, IntWritable, DoubleWritable, cc.mrlda.Document (Document.java:241-263),
    PairOfIntFloat, HMapIDW — all big-endian.
"""
import gzip
import struct
import zlib

SYNC_ESCAPE = -1
DEFAULT_CODEC = "org.apache.hadoop.io.compress.DefaultCodec"
GZIP_CODEC = "org.apache.hadoop.io.compress.GzipCodec"


def write_vint(i):
    if -112 <= i <= 127:
        return struct.pack(">b", i)
    length = -112
    if i < 0:
        i ^= -1
        length = -120
    tmp, n = i, 0
    while tmp:
        tmp >>= 8
        n += 1
    out = struct.pack(">b", length - n)
    for idx in range(n, 0, -1):
        out += bytes([(i >> ((idx - 1) * 8)) & 0xFF])
    return out


def read_vint(buf, pos):
    first = struct.unpack_from(">b", buf, pos)[0]
    pos += 1
    if first >= -112:
        return first, pos
    negative = first < -120
    n = (-119 - first) if negative else (-111 - first)
    v = 0
    for _ in range(n - 1):
        v = (v << 8) | buf[pos]
        pos += 1
    return (v ^ -1) if negative else v, pos


def text(s):
    b = s.encode("utf-8")
    return write_vint(len(b)) + b


def read_file(path):
    """-> (key_class, value_class, sync, [(key_bytes, value_bytes)], n_sync_escapes)"""
    buf = open(path, "rb").read()
    assert buf[:3] == b"SEQ" and buf[3] == 6, buf[:4]
    pos = 4
    n, pos = read_vint(buf, pos)
    key_class = buf[pos:pos + n].decode()
    pos += n
    n, pos = read_vint(buf, pos)
    value_class = buf[pos:pos + n].decode()
    pos += n
    compressed, block = buf[pos], buf[pos + 1]
    assert compressed == 0 and block == 0
    pos += 2
    (meta,) = struct.unpack_from(">i", buf, pos)
    assert meta == 0
    pos += 4
    sync = buf[pos:pos + 16]
    pos += 16
    records, escapes = [], 0
    while pos < len(buf):
        (rec_len,) = struct.unpack_from(">i", buf, pos)
        pos += 4
        if rec_len == SYNC_ESCAPE:
            assert buf[pos:pos + 16] == sync
            pos += 16
            escapes += 1
            continue
        (key_len,) = struct.unpack_from(">i", buf, pos)
        pos += 4
        assert 0 <= key_len <= rec_len
        records.append((buf[pos:pos + key_len], buf[pos + key_len:pos + rec_len]))
        pos += rec_len
    assert pos == len(buf)
    return key_class, value_class, sync, records, escapes


def write_file(path, key_class, value_class, records, sync=bytes(range(16)), sync_every=None, version=6):
    """sync_every: put a sync escape before every record whose index is a multiple of it (None: never
    after the header) — a reader must cope with any placement.  version 5 files have no metadata
    block."""
    out = b"SEQ" + bytes([version]) + text(key_class) + text(value_class) + b"\x00\x00"
    if version >= 6:
        out += struct.pack(">i", 0)
    out += sync
    for i, (k, v) in enumerate(records):
        if sync_every and i and i % sync_every == 0:
            out += struct.pack(">i", SYNC_ESCAPE) + sync
        out += struct.pack(">ii", len(k) + len(v), len(k)) + k + v
    open(path, "wb").write(out)


def _deflate(data, codec):
    return gzip.compress(data) if codec == GZIP_CODEC else zlib.compress(data)


def write_compressed_file(path, key_class, value_class, records, block, codec=DEFAULT_CODEC,
                          sync=bytes(range(16)), per_block=3, sync_every=None):
    """SequenceFile.RecordCompressWriter (block=False: every value is one codec stream, keys stay
    plain) or SequenceFile.BlockCompressWriter (block=True: sync escape, vint record count, then key
    lengths / keys / value lengths / values, each a vint byte count + one codec stream)."""
    out = (b"SEQ\x06" + text(key_class) + text(value_class) + b"\x01" + (b"\x01" if block else b"\x00") +
           text(codec) + struct.pack(">i", 0) + sync)
    if not block:
        for i, (k, v) in enumerate(records):
            if sync_every and i and i % sync_every == 0:
                out += struct.pack(">i", SYNC_ESCAPE) + sync
            cv = _deflate(v, codec)
            out += struct.pack(">ii", len(k) + len(cv), len(k)) + k + cv
    else:
        for lo in range(0, len(records), per_block):
            chunk = records[lo:lo + per_block]
            out += struct.pack(">i", SYNC_ESCAPE) + sync + write_vint(len(chunk))
            for buf in (b"".join(write_vint(len(k)) for k, _ in chunk), b"".join(k for k, _ in chunk),
                        b"".join(write_vint(len(v)) for _, v in chunk), b"".join(v for _, v in chunk)):
                c = _deflate(buf, codec)
                out += write_vint(len(c)) + c
    open(path, "wb").write(out)


def document_bytes(content, gamma):
    """Document.write: content None or {} -> 0 entries; gamma None -> 0 topics."""
    items = list(content.items()) if content else []
    out = struct.pack(">i", len(items))
    for t, c in items:
        out += struct.pack(">ii", t, c)
    g = list(gamma) if gamma is not None else []
    out += struct.pack(">i", len(g))
    for x in g:
        out += struct.pack(">d", x)
    return out


def parse_document(b):
    (n,) = struct.unpack_from(">i", b, 0)
    pos = 4
    content = None
    if n > 0:
        content = {}
        for _ in range(n):
            t, c = struct.unpack_from(">ii", b, pos)
            pos += 8
            content[t] = c  # HMapII.put: a repeated id keeps the last count (Document.java:155-160)
    (k,) = struct.unpack_from(">i", b, pos)
    pos += 4
    gamma = None
    if k > 0:
        gamma = list(struct.unpack_from(">%dd" % k, b, pos))
        pos += 8 * k
    assert pos == len(b)
    return content, gamma
