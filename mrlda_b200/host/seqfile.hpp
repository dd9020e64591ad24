// seqfile.hpp — the wire formats either side of the hot path:
// Hadoop SequenceFile (versions 5 and 6; written uncompressed, read uncompressed, record- or
// block-compressed with the zlib / gzip codecs) and the Writables Mr.LDA stores in it.
//
//   cc.mrlda.Document                      Document.java:147-172 (readFields), :241-263 (write)
//   alpha files  <IntWritable, DoubleWritable>   VariationalInference.java:521-558
//   beta files   <PairOfIntFloat, HMapIDW>       VariationalInference.java:221-223, TermReducer.java:173-195
//   gamma files  <IntWritable, Document>         VariationalInference.java:232-235, DocumentMapper.java:342-346
//
// The SequenceFile container and the cloud9 Writables (PairOfIntFloat, HMapIDW) are third-party
// code that is not under /root/reference; their layouts are restated from the published formats:
// header "SEQ" 0x06, key/value class names as vint-prefixed UTF-8, two compression booleans,
// metadata count, 16-byte sync marker; records int32 recordLength, int32 keyLength, key, value; a
// sync escape (int32 -1 + marker) before a record once 2000 bytes have passed since the last one.
// All integers and IEEE values are big-endian (java.io.DataOutput).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include <unistd.h>
#include <zlib.h>

namespace mrlda_host {

static const char *const kIntWritable = "org.apache.hadoop.io.IntWritable";
static const char *const kDoubleWritable = "org.apache.hadoop.io.DoubleWritable";
static const char *const kDocument = "cc.mrlda.Document";
static const char *const kPairOfIntFloat = "edu.umd.cloud9.io.pair.PairOfIntFloat";
static const char *const kHMapIDW = "edu.umd.cloud9.io.map.HMapIDW";

struct ByteWriter {
  std::vector<uint8_t> buf;
  void u8(uint8_t v) { buf.push_back(v); }
  // room for n more bytes; returns where they go (callers fill all n)
  uint8_t *grow(size_t n) {
    const size_t at = buf.size();
    buf.resize(at + n);
    return buf.data() + at;
  }
  static void put32(uint8_t *p, int32_t v) {
    const uint32_t u = __builtin_bswap32((uint32_t)v);  // big-endian, as java.io.DataOutput writes
    memcpy(p, &u, 4);
  }
  static void put64(uint8_t *p, int64_t v) {
    const uint64_t u = __builtin_bswap64((uint64_t)v);
    memcpy(p, &u, 8);
  }
  void i32(int32_t v) { put32(grow(4), v); }
  void i64(int64_t v) { put64(grow(8), v); }
  void f32(float v) {
    int32_t i;
    memcpy(&i, &v, 4);
    i32(i);
  }
  void f64(double v) {
    int64_t i;
    memcpy(&i, &v, 8);
    i64(i);
  }
  void bytes(const void *p, size_t n) {
    const uint8_t *b = (const uint8_t *)p;
    buf.insert(buf.end(), b, b + n);
  }
  // org.apache.hadoop.io.WritableUtils.writeVLong
  void vlong(int64_t i) {
    if (i >= -112 && i <= 127) {
      u8((uint8_t)(int8_t)i);
      return;
    }
    int len = -112;
    if (i < 0) {
      i ^= -1LL;
      len = -120;
    }
    for (int64_t tmp = i; tmp != 0; tmp >>= 8) len--;
    u8((uint8_t)(int8_t)len);
    len = (len < -120) ? -(len + 120) : -(len + 112);
    for (int idx = len; idx != 0; idx--) u8((uint8_t)((i >> ((idx - 1) * 8)) & 0xff));
  }
  void text(const std::string &s) {  // org.apache.hadoop.io.Text.writeString
    vlong((int64_t)s.size());
    bytes(s.data(), s.size());
  }
};

struct ByteReader {
  const uint8_t *p, *end;
  ByteReader(const uint8_t *b, size_t n) : p(b), end(b + n) {}
  void need(size_t n) const {
    if ((size_t)(end - p) < n) throw std::runtime_error("unexpected end of data");
  }
  uint8_t u8() {
    need(1);
    return *p++;
  }
  static int32_t get32(const uint8_t *q) {
    uint32_t u;
    memcpy(&u, q, 4);
    return (int32_t)__builtin_bswap32(u);
  }
  static int64_t get64(const uint8_t *q) {
    uint64_t u;
    memcpy(&u, q, 8);
    return (int64_t)__builtin_bswap64(u);
  }
  int32_t i32() {
    need(4);
    const int32_t v = get32(p);
    p += 4;
    return v;
  }
  int64_t i64() {
    need(8);
    const int64_t v = get64(p);
    p += 8;
    return v;
  }
  float f32() {
    int32_t i = i32();
    float v;
    memcpy(&v, &i, 4);
    return v;
  }
  double f64() {
    int64_t i = i64();
    double v;
    memcpy(&v, &i, 8);
    return v;
  }
  int64_t vlong() {  // WritableUtils.readVLong
    int8_t first = (int8_t)u8();
    int len = first >= -112 ? 1 : (first < -120 ? -119 - first : -111 - first);
    if (len == 1) return first;
    int64_t i = 0;
    for (int idx = 0; idx < len - 1; idx++) i = (i << 8) | u8();
    const bool neg = first < -120 || (first >= -112 && first < 0);
    return neg ? (i ^ -1LL) : i;
  }
  std::string text() {
    int64_t n = vlong();
    if (n < 0) throw std::runtime_error("negative string length");
    need((size_t)n);
    std::string s((const char *)p, (size_t)n);
    p += n;
    return s;
  }
  bool eof() const { return p >= end; }
};

// HMapII.put overwrites an existing key, so a repeated term id inside one Document record keeps its
// first position and takes the last count.  Repeats are found with a small open-addressing table
// (term id -> position) that is reused from record to record; `stamp` marks this record's slots.
struct TermSlots {
  void begin(size_t n) {
    size_t slots = 16;
    while (slots < n * 2) slots <<= 1;
    if (key_.size() < slots) {
      key_.assign(slots, 0);
      pos_.assign(slots, 0);
      stamp_of_.assign(slots, 0);
      stamp_ = 0;
    }
    mask_ = key_.size() - 1;
    if (++stamp_ == 0) {  // the 32-bit stamp wrapped: start over with clean slots
      std::fill(stamp_of_.begin(), stamp_of_.end(), 0u);
      stamp_ = 1;
    }
  }
  // position of `id` among this record's entries, or -1 after remembering it at `pos`
  int64_t find_or_add(int32_t id, int64_t pos) {
    size_t h = ((uint32_t)id * 2654435761u) & mask_;
    for (;;) {
      if (stamp_of_[h] != stamp_) {
        stamp_of_[h] = stamp_;
        key_[h] = id;
        pos_[h] = pos;
        return -1;
      }
      if (key_[h] == id) return pos_[h];
      h = (h + 1) & mask_;
    }
  }

 private:
  std::vector<int32_t> key_;
  std::vector<int64_t> pos_;
  std::vector<uint32_t> stamp_of_;
  uint32_t stamp_ = 0;
  size_t mask_ = 0;
};

// cc.mrlda.Document (Document.java:241-263 / :147-172)
struct Document {
  std::vector<std::pair<int32_t, int32_t>> content;  // (termID, count); empty = null content
  std::vector<double> gamma;                         // empty = null gamma
  int64_t tokens() const {
    int64_t t = 0;
    for (auto &e : content) t += e.second;
    return t;
  }
  void write(ByteWriter &w) const {
    // one allocation, then straight stores: 4 + 8n + 4 + 8k bytes
    uint8_t *q = w.grow(8 + 8 * content.size() + 8 * gamma.size());
    ByteWriter::put32(q, (int32_t)content.size());  // 0 when content == null
    q += 4;
    for (auto &e : content) {
      ByteWriter::put32(q, e.first);
      ByteWriter::put32(q + 4, e.second);
      q += 8;
    }
    ByteWriter::put32(q, (int32_t)gamma.size());  // 0 when gamma == null
    q += 4;
    for (double g : gamma) {
      int64_t bits;
      memcpy(&bits, &g, 8);
      ByteWriter::put64(q, bits);
      q += 8;
    }
  }
  void read(ByteReader &r) {
    content.clear();
    gamma.clear();
    const int32_t n = r.i32();
    if (n > 0) {  // numEntries <= 0  =>  content = null
      r.need((size_t)n * 8);
      content.reserve((size_t)n);
      slots_.begin((size_t)n);
      for (int32_t i = 0; i < n; i++) {
        const int32_t id = ByteReader::get32(r.p), c = ByteReader::get32(r.p + 4);
        r.p += 8;
        const int64_t at = slots_.find_or_add(id, (int64_t)content.size());
        if (at < 0)
          content.emplace_back(id, c);
        else
          content[(size_t)at].second = c;
      }
    }
    const int32_t k = r.i32();
    if (k > 0) {  // numTopics <= 0  =>  gamma = null
      r.need((size_t)k * 8);
      gamma.resize((size_t)k);
      for (int32_t i = 0; i < k; i++) {
        const int64_t bits = ByteReader::get64(r.p);
        r.p += 8;
        memcpy(&gamma[(size_t)i], &bits, 8);
      }
    }
  }

 private:
  TermSlots slots_;
};

class SeqWriter {
 public:
  SeqWriter(const std::string &path, const std::string &key_class, const std::string &value_class)
      : f_(fopen(path.c_str(), "wb")) {
    if (!f_) throw std::runtime_error("cannot create " + path);
    setvbuf(f_, nullptr, _IOFBF, 1 << 20);  // records are a few KB: write them out 1 MB at a time
    ByteWriter h;
    h.bytes("SEQ", 3);
    h.u8(6);
    h.text(key_class);
    h.text(value_class);
    h.u8(0);   // value compression
    h.u8(0);   // block compression
    h.i32(0);  // metadata entries
    // the marker is an MD5 of a fresh UID in Hadoop; any 16 bytes serve, derive them from the path
    uint64_t a = 0xcbf29ce484222325ull, b = 0x9e3779b97f4a7c15ull;
    for (char c : path + key_class + value_class) {
      a = (a ^ (uint8_t)c) * 0x100000001b3ull;
      b = (b + a) * 0xff51afd7ed558ccdull;
    }
    memcpy(sync_, &a, 8);
    memcpy(sync_ + 8, &b, 8);
    h.bytes(sync_, 16);
    put(h.buf);
  }
  ~SeqWriter() {  // a destructor cannot report: writers call close() themselves to see flush errors
    if (f_) fclose(f_);
  }
  void append(const ByteWriter &key, const ByteWriter &value) {
    if (pos_ >= last_sync_ + 2000) {  // SequenceFile.Writer.checkAndWriteSync
      ByteWriter s;
      s.i32(-1);
      s.bytes(sync_, 16);
      put(s.buf);
      last_sync_ = pos_;
    }
    uint8_t head[8];
    ByteWriter::put32(head, (int32_t)(key.buf.size() + value.buf.size()));
    ByteWriter::put32(head + 4, (int32_t)key.buf.size());
    put(head, 8);
    put(key.buf);
    put(value.buf);
  }
  // Positions of n further records of rec_bytes each (8-byte record header + key + value), exactly
  // where append() would put them, for writers that build records concurrently and store them with
  // write_at(): `at` is where the record header goes; with sync_before the 20 bytes in front of it
  // are a sync escape (sync_escape()).  The writer's position moves behind the last record; the
  // caller must fill every planned byte before close().
  struct Slot {
    int64_t at;
    bool sync_before;
  };
  std::vector<Slot> plan(size_t n, size_t rec_bytes) {
    if (fflush(f_) != 0) throw std::runtime_error("short write");
    std::vector<Slot> slots(n);
    for (size_t i = 0; i < n; i++) {
      const bool sync = pos_ >= last_sync_ + 2000;
      if (sync) {
        pos_ += 20;
        last_sync_ = pos_;
      }
      slots[i] = Slot{pos_, sync};
      pos_ += (int64_t)rec_bytes;
    }
    if (fseeko(f_, (off_t)pos_, SEEK_SET) != 0) throw std::runtime_error("cannot seek");
    return slots;
  }
  void sync_escape(uint8_t *dst) const {  // 20 bytes
    ByteWriter::put32(dst, -1);
    memcpy(dst + 4, sync_, 16);
  }
  void write_at(int64_t at, const uint8_t *b, size_t n) const {  // thread-safe (pwrite)
    const int fd = fileno(f_);
    while (n > 0) {
      const ssize_t w = pwrite(fd, b, n, (off_t)at);
      if (w <= 0) throw std::runtime_error("short write");
      b += w;
      at += w;
      n -= (size_t)w;
    }
  }
  // flushes the buffered tail; a failed flush (disk full) is an error like a short write
  void close() {
    FILE *f = f_;
    f_ = nullptr;
    if (f && fclose(f) != 0) throw std::runtime_error("short write (close)");
  }

 private:
  void put(const std::vector<uint8_t> &b) { put(b.data(), b.size()); }
  void put(const uint8_t *b, size_t n) {
    if (n && fwrite(b, 1, n, f_) != n) throw std::runtime_error("short write");
    pos_ += (int64_t)n;
  }
  FILE *f_;
  uint8_t sync_[16];
  int64_t pos_ = 0, last_sync_ = 0;
};

// Streaming reader: the file is read through a bounded buffer, one record (or one compressed block)
// at a time, so memory does not grow with the file.  Handles what SequenceFile.Reader handles for
// the codecs that need no native Hadoop library: uncompressed records, record-compressed values
// and block-compressed files with org.apache.hadoop.io.compress.DefaultCodec (zlib streams) or
// GzipCodec (the reference opens its inputs with a plain SequenceFile.Reader,
// VariationalInference.java:282,288 and DocumentMapper.java:475-536, so whatever codec the cluster
// wrote is accepted there).  Other codecs (Snappy, LZ4, BZip2) are refused by name.
class SeqReader {
 public:
  explicit SeqReader(const std::string &path) : path_(path), f_(fopen(path.c_str(), "rb")) {
    if (!f_) throw std::runtime_error("cannot open " + path);
    setvbuf(f_, nullptr, _IOFBF, 1 << 20);
    try {
      uint8_t magic[4];
      read_exact(magic, 4, "header");
      if (memcmp(magic, "SEQ", 3) != 0) throw std::runtime_error(path + ": not a SequenceFile");
      const int ver = magic[3];
      if (ver < 5 || ver > 6) throw std::runtime_error(path + ": unsupported SequenceFile version");
      key_class = read_text();
      value_class = read_text();
      uint8_t flags[2];
      read_exact(flags, 2, "header");
      compressed_ = flags[0] != 0;
      block_ = flags[1] != 0;
      if (compressed_) {
        codec = read_text();
        if (codec != "org.apache.hadoop.io.compress.DefaultCodec" &&
            codec != "org.apache.hadoop.io.compress.GzipCodec")
          throw std::runtime_error(path + ": SequenceFile codec " + codec +
                                   " is not supported (DefaultCodec and GzipCodec are)");
      }
      if (ver >= 6) {  // the metadata block came with version 6 (VERSION_WITH_METADATA)
        const int32_t meta = read_i32("header");
        for (int32_t i = 0; i < meta; i++) {
          read_text();
          read_text();
        }
      }
      read_exact(sync_, 16, "header");
    } catch (...) {
      fclose(f_);
      f_ = nullptr;
      throw;
    }
  }
  ~SeqReader() {
    if (f_) fclose(f_);
  }
  SeqReader(const SeqReader &) = delete;
  SeqReader &operator=(const SeqReader &) = delete;

  // next record; key/value stay valid until the following call
  bool next(const uint8_t *&key, size_t &klen, const uint8_t *&val, size_t &vlen) {
    if (block_) return next_in_block(key, klen, val, vlen);
    for (;;) {
      uint8_t h[4];
      const size_t got = fread(h, 1, 4, f_);
      if (got == 0) return false;
      if (got != 4) throw std::runtime_error("unexpected end of data");
      const int32_t rec = be32(h);
      if (rec == -1) {  // sync escape
        check_sync();
        continue;
      }
      const int32_t kl = read_i32("record header");
      if (rec < 0 || kl < 0 || kl > rec) throw std::runtime_error("corrupt record header");
      rec_.resize((size_t)rec);
      read_exact(rec_.data(), (size_t)rec, "record");
      key = rec_.data();
      klen = (size_t)kl;
      if (compressed_) {  // the value alone is a codec stream
        inflate_into(rec_.data() + kl, (size_t)(rec - kl), val_);
        val = val_.data();
        vlen = val_.size();
      } else {
        val = rec_.data() + kl;
        vlen = (size_t)(rec - kl);
      }
      return true;
    }
  }
  std::string key_class, value_class, codec;
  bool compressed() const { return compressed_; }
  bool block_compressed() const { return block_; }

 private:
  static int32_t be32(const uint8_t *h) {
    return (int32_t)(((uint32_t)h[0] << 24) | ((uint32_t)h[1] << 16) | ((uint32_t)h[2] << 8) | h[3]);
  }
  void read_exact(uint8_t *dst, size_t n, const char *what) {
    if (n && fread(dst, 1, n, f_) != n)
      throw std::runtime_error(std::string("unexpected end of data (") + what + ")");
  }
  int32_t read_i32(const char *what) {
    uint8_t h[4];
    read_exact(h, 4, what);
    return be32(h);
  }
  int64_t read_vlong() {  // WritableUtils.readVLong from the stream
    uint8_t b[9];
    read_exact(b, 1, "vint");
    const int8_t first = (int8_t)b[0];
    const int len = first >= -112 ? 1 : (first < -120 ? -119 - first : -111 - first);
    read_exact(b + 1, (size_t)(len - 1), "vint");
    ByteReader r(b, (size_t)len);
    return r.vlong();
  }
  std::string read_text() {
    const int64_t n = read_vlong();
    if (n < 0 || n > (1 << 20)) throw std::runtime_error("corrupt string length");
    std::string s((size_t)n, '\0');
    read_exact(reinterpret_cast<uint8_t *>(&s[0]), (size_t)n, "string");
    return s;
  }
  void check_sync() {
    uint8_t m[16];
    read_exact(m, 16, "sync marker");
    if (memcmp(m, sync_, 16) != 0) throw std::runtime_error("corrupt sync marker");
  }
  // zlib or gzip stream -> out (the 32 added to the window bits selects the format from the header)
  void inflate_into(const uint8_t *src, size_t n, std::vector<uint8_t> &out) {
    z_stream z;
    memset(&z, 0, sizeof z);
    if (inflateInit2(&z, 15 + 32) != Z_OK) throw std::runtime_error("inflateInit2 failed");
    out.resize(std::max<size_t>(out.capacity(), std::max<size_t>(4 * n, 256)));
    z.next_in = const_cast<Bytef *>(src);
    z.avail_in = (uInt)n;
    size_t produced = 0;
    for (;;) {
      z.next_out = out.data() + produced;
      z.avail_out = (uInt)(out.size() - produced);
      const int rc = inflate(&z, Z_NO_FLUSH);
      produced = out.size() - z.avail_out;
      if (rc == Z_STREAM_END) break;
      if (rc != Z_OK && rc != Z_BUF_ERROR) {
        inflateEnd(&z);
        throw std::runtime_error(path_ + ": corrupt compressed data");
      }
      if (z.avail_out == 0) {
        out.resize(out.size() * 2);
      } else if (z.avail_in == 0) {
        inflateEnd(&z);
        throw std::runtime_error(path_ + ": truncated compressed data");
      }
    }
    inflateEnd(&z);
    out.resize(produced);
  }
  void read_block_buffer(std::vector<uint8_t> &out) {
    const int64_t n = read_vlong();
    if (n < 0 || n > ((int64_t)1 << 31)) throw std::runtime_error("corrupt block buffer length");
    rec_.resize((size_t)n);
    read_exact(rec_.data(), (size_t)n, "block buffer");
    inflate_into(rec_.data(), (size_t)n, out);
  }
  // Block-compressed files (SequenceFile.BlockCompressWriter): sync escape, record count, then the
  // key lengths, keys, value lengths and values of the block as four codec streams.
  bool next_in_block(const uint8_t *&key, size_t &klen, const uint8_t *&val, size_t &vlen) {
    while (left_ == 0) {
      uint8_t h[4];
      const size_t got = fread(h, 1, 4, f_);
      if (got == 0) return false;
      if (got != 4) throw std::runtime_error("unexpected end of data");
      if (be32(h) != -1) throw std::runtime_error("corrupt block header (no sync escape)");
      check_sync();
      const int64_t n = read_vlong();
      if (n < 0) throw std::runtime_error("corrupt block record count");
      read_block_buffer(klens_);
      read_block_buffer(keys_);
      read_block_buffer(vlens_);
      read_block_buffer(vals_);
      left_ = n;
      kl_off_ = k_off_ = vl_off_ = v_off_ = 0;
    }
    ByteReader kr(klens_.data() + kl_off_, klens_.size() - kl_off_);
    ByteReader vr(vlens_.data() + vl_off_, vlens_.size() - vl_off_);
    const int64_t kl = kr.vlong(), vl = vr.vlong();
    kl_off_ = (size_t)(kr.p - klens_.data());
    vl_off_ = (size_t)(vr.p - vlens_.data());
    if (kl < 0 || vl < 0 || k_off_ + (size_t)kl > keys_.size() || v_off_ + (size_t)vl > vals_.size())
      throw std::runtime_error("corrupt block (lengths exceed the data)");
    key = keys_.data() + k_off_;
    klen = (size_t)kl;
    val = vals_.data() + v_off_;
    vlen = (size_t)vl;
    k_off_ += (size_t)kl;
    v_off_ += (size_t)vl;
    left_--;
    return true;
  }

  std::string path_;
  FILE *f_;
  bool compressed_ = false, block_ = false;
  uint8_t sync_[16];
  std::vector<uint8_t> rec_, val_;
  std::vector<uint8_t> klens_, keys_, vlens_, vals_;
  int64_t left_ = 0;
  size_t kl_off_ = 0, k_off_ = 0, vl_off_ = 0, v_off_ = 0;
};

}  // namespace mrlda_host
