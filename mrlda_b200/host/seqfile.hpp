// seqfile.hpp — the wire formats either side of the hot path:
// Hadoop SequenceFile (version 6, uncompressed records) and the Writables Mr.LDA stores in it.
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
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace mrlda_host {

static const char *const kIntWritable = "org.apache.hadoop.io.IntWritable";
static const char *const kDoubleWritable = "org.apache.hadoop.io.DoubleWritable";
static const char *const kDocument = "cc.mrlda.Document";
static const char *const kPairOfIntFloat = "edu.umd.cloud9.io.pair.PairOfIntFloat";
static const char *const kHMapIDW = "edu.umd.cloud9.io.map.HMapIDW";

struct ByteWriter {
  std::vector<uint8_t> buf;
  void u8(uint8_t v) { buf.push_back(v); }
  void i32(int32_t v) {
    uint32_t u = (uint32_t)v;
    for (int s = 24; s >= 0; s -= 8) buf.push_back((uint8_t)(u >> s));
  }
  void i64(int64_t v) {
    uint64_t u = (uint64_t)v;
    for (int s = 56; s >= 0; s -= 8) buf.push_back((uint8_t)(u >> s));
  }
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
  int32_t i32() {
    need(4);
    uint32_t u = 0;
    for (int i = 0; i < 4; i++) u = (u << 8) | *p++;
    return (int32_t)u;
  }
  int64_t i64() {
    need(8);
    uint64_t u = 0;
    for (int i = 0; i < 8; i++) u = (u << 8) | *p++;
    return (int64_t)u;
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
    w.i32((int32_t)content.size());  // 0 when content == null
    for (auto &e : content) {
      w.i32(e.first);
      w.i32(e.second);
    }
    w.i32((int32_t)gamma.size());  // 0 when gamma == null
    for (double g : gamma) w.f64(g);
  }
  void read(ByteReader &r) {
    content.clear();
    gamma.clear();
    int32_t n = r.i32();
    if (n > 0) {  // numEntries <= 0  =>  content = null
      content.reserve((size_t)n);
      for (int32_t i = 0; i < n; i++) {
        int32_t id = r.i32(), c = r.i32();
        bool dup = false;  // HMapII.put overwrites an existing key
        for (auto &e : content)
          if (e.first == id) {
            e.second = c;
            dup = true;
            break;
          }
        if (!dup) content.emplace_back(id, c);
      }
    }
    int32_t k = r.i32();
    if (k > 0) {  // numTopics <= 0  =>  gamma = null
      gamma.resize((size_t)k);
      for (int32_t i = 0; i < k; i++) gamma[i] = r.f64();
    }
  }
};

class SeqWriter {
 public:
  SeqWriter(const std::string &path, const std::string &key_class, const std::string &value_class)
      : f_(fopen(path.c_str(), "wb")) {
    if (!f_) throw std::runtime_error("cannot create " + path);
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
  ~SeqWriter() { close(); }
  void append(const ByteWriter &key, const ByteWriter &value) {
    if (pos_ >= last_sync_ + 2000) {  // SequenceFile.Writer.checkAndWriteSync
      ByteWriter s;
      s.i32(-1);
      s.bytes(sync_, 16);
      put(s.buf);
      last_sync_ = pos_;
    }
    ByteWriter r;
    r.i32((int32_t)(key.buf.size() + value.buf.size()));
    r.i32((int32_t)key.buf.size());
    put(r.buf);
    put(key.buf);
    put(value.buf);
  }
  void close() {
    if (f_) fclose(f_);
    f_ = nullptr;
  }

 private:
  void put(const std::vector<uint8_t> &b) {
    if (!b.empty() && fwrite(b.data(), 1, b.size(), f_) != b.size())
      throw std::runtime_error("short write");
    pos_ += (int64_t)b.size();
  }
  FILE *f_;
  uint8_t sync_[16];
  int64_t pos_ = 0, last_sync_ = 0;
};

class SeqReader {
 public:
  explicit SeqReader(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    data_.resize((size_t)n);
    if (n && fread(data_.data(), 1, (size_t)n, f) != (size_t)n) {
      fclose(f);
      throw std::runtime_error("short read " + path);
    }
    fclose(f);
    ByteReader r(data_.data(), data_.size());
    r.need(4);
    if (memcmp(r.p, "SEQ", 3) != 0) throw std::runtime_error(path + ": not a SequenceFile");
    r.p += 3;
    int ver = r.u8();
    if (ver < 5 || ver > 6) throw std::runtime_error(path + ": unsupported SequenceFile version");
    key_class = r.text();
    value_class = r.text();
    bool comp = r.u8() != 0, block = r.u8() != 0;
    if (comp || block)
      throw std::runtime_error(path + ": compressed SequenceFiles are not supported (ParseCorpus "
                                      "writes uncompressed records, ParseCorpus.java:680)");
    if (ver >= 6) {  // the metadata block came with version 6 (VERSION_WITH_METADATA)
      int32_t meta = r.i32();
      for (int32_t i = 0; i < meta; i++) {
        r.text();
        r.text();
      }
    }
    r.need(16);
    memcpy(sync_, r.p, 16);
    r.p += 16;
    off_ = (size_t)(r.p - data_.data());
  }
  // next record; key/value point into the file buffer
  bool next(const uint8_t *&key, size_t &klen, const uint8_t *&val, size_t &vlen) {
    for (;;) {
      if (off_ >= data_.size()) return false;
      ByteReader r(data_.data() + off_, data_.size() - off_);
      int32_t rec = r.i32();
      if (rec == -1) {  // sync escape
        r.need(16);
        if (memcmp(r.p, sync_, 16) != 0) throw std::runtime_error("corrupt sync marker");
        off_ += 20;
        continue;
      }
      int32_t kl = r.i32();
      if (rec < 0 || kl < 0 || kl > rec) throw std::runtime_error("corrupt record header");
      r.need((size_t)rec);
      key = r.p;
      klen = (size_t)kl;
      val = r.p + kl;
      vlen = (size_t)(rec - kl);
      off_ += 8 + (size_t)rec;
      return true;
    }
  }
  std::string key_class, value_class;

 private:
  std::vector<uint8_t> data_;
  uint8_t sync_[16];
  size_t off_ = 0;
};

}  // namespace mrlda_host
