// model_io.hpp — reading the ParseCorpus document files into CSR and reading / writing the
// alpha-N, beta-N and gamma-N model files of cc.mrlda.VariationalInference, with the reference's
// validity checks (Preconditions.checkArgument -> std::invalid_argument).
#pragma once
#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "seqfile.hpp"

namespace mrlda_host {

struct Corpus {
  std::vector<int64_t> row_ptr{0};
  std::vector<int32_t> term_id, count, doc_id;
  std::vector<double> gamma;  // D*K when every document carried a K-vector, else empty
  int64_t n_docs() const { return (int64_t)doc_id.size(); }
};

inline bool is_dir(const std::string &p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
inline bool exists(const std::string &p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0;
}

// Hadoop's FileInputFormat skips names starting with '_' or '.'; part files are read in name order.
inline std::vector<std::string> input_files(const std::string &path) {
  std::vector<std::string> out;
  if (!is_dir(path)) {
    if (!exists(path)) throw std::invalid_argument("input path does not exist: " + path);
    out.push_back(path);
    return out;
  }
  DIR *d = opendir(path.c_str());
  if (!d) throw std::invalid_argument("cannot list " + path);
  while (dirent *e = readdir(d)) {
    std::string n = e->d_name;
    if (n.empty() || n[0] == '_' || n[0] == '.') continue;
    std::string full = path + (path.back() == '/' ? "" : "/") + n;
    if (!is_dir(full)) out.push_back(full);
  }
  closedir(d);
  std::sort(out.begin(), out.end());
  return out;
}

// SequenceFile<IntWritable, cc.mrlda.Document> -> CSR.  Entry order within a document is the
// writer's HMapII iteration order (arbitrary); it is kept as found.  Only the documents with
// global index in [first, last) are kept (the others are skipped without being parsed).
inline Corpus read_corpus(const std::string &path, int n_topics, int n_terms, int64_t first = 0,
                          int64_t last = INT64_MAX) {
  Corpus c;
  bool all_gamma = true;
  int64_t index = 0;
  for (const std::string &file : input_files(path)) {
    if (index >= last) break;
    SeqReader r(file);
    if (r.key_class != kIntWritable || r.value_class != kDocument)
      throw std::invalid_argument(file + ": expected <" + kIntWritable + ", " + kDocument + ">, found <" +
                                  r.key_class + ", " + r.value_class + ">");
    const uint8_t *k, *v;
    size_t kl, vl;
    Document doc;
    while (index < last && r.next(k, kl, v, vl)) {
      if (index++ < first) continue;
      ByteReader kr(k, kl), vr(v, vl);
      c.doc_id.push_back(kr.i32());
      doc.read(vr);
      for (auto &e : doc.content) {
        if (e.first < 1 || e.first > n_terms)
          throw std::invalid_argument("term id " + std::to_string(e.first) + " outside [1," +
                                      std::to_string(n_terms) + "] in document " +
                                      std::to_string(c.doc_id.back()));
        c.term_id.push_back(e.first);
        c.count.push_back(e.second);
      }
      c.row_ptr.push_back((int64_t)c.term_id.size());
      if ((int)doc.gamma.size() == n_topics && all_gamma)
        c.gamma.insert(c.gamma.end(), doc.gamma.begin(), doc.gamma.end());
      else
        all_gamma = false;  // DocumentMapper.java:184: a stored gamma of another length is ignored
    }
  }
  if (!all_gamma) c.gamma.clear();
  return c;
}

// The input split of one rank (the map phase's input splits): a first streaming pass reads only
// the entry count at the head of every Document value and cuts the document range into `world`
// contiguous pieces balanced by entries, as corpus.shard_bounds does; the second pass parses this
// rank's piece alone, so no rank ever holds more than its own shard.  first_doc receives the
// global index of the shard's first document.
inline Corpus read_corpus_shard(const std::string &path, int n_topics, int n_terms, int rank, int world,
                                int64_t *first_doc = nullptr, int64_t *total_docs = nullptr) {
  if (world <= 1) {
    Corpus c = read_corpus(path, n_topics, n_terms);
    if (first_doc) *first_doc = 0;
    if (total_docs) *total_docs = c.n_docs();
    return c;
  }
  std::vector<int64_t> rp{0};
  for (const std::string &file : input_files(path)) {
    SeqReader r(file);
    if (r.key_class != kIntWritable || r.value_class != kDocument)
      throw std::invalid_argument(file + ": expected <" + kIntWritable + ", " + kDocument + ">, found <" +
                                  r.key_class + ", " + r.value_class + ">");
    const uint8_t *k, *v;
    size_t kl, vl;
    while (r.next(k, kl, v, vl)) {
      ByteReader vr(v, vl);
      const int32_t n = vr.i32();
      rp.push_back(rp.back() + (n > 0 ? n : 0));
    }
  }
  const int64_t D = (int64_t)rp.size() - 1, Z = rp.back();
  auto bound = [&](int r) -> int64_t {
    if (r <= 0) return 0;
    if (r >= world) return D;
    return std::lower_bound(rp.begin(), rp.end(), Z * r / world) - rp.begin();
  };
  const int64_t lo = std::min(bound(rank), D);
  const int64_t hi = std::min(std::max(bound(rank + 1), lo), D);
  if (first_doc) *first_doc = lo;
  if (total_docs) *total_docs = D;
  return read_corpus(path, n_topics, n_terms, lo, hi);
}

// exportAlpha (VariationalInference.java:549-558)
inline void write_alpha(const std::string &path, const std::vector<double> &alpha) {
  SeqWriter w(path, kIntWritable, kDoubleWritable);
  for (size_t i = 0; i < alpha.size(); i++) {
    ByteWriter k, v;
    k.i32((int32_t)i + 1);
    v.f64(alpha[i]);
    w.append(k, v);
  }
}

// importAlpha (VariationalInference.java:521-540)
inline std::vector<double> read_alpha(const std::string &path, int n_topics) {
  SeqReader r(path);
  std::vector<double> alpha((size_t)n_topics, 0.0);
  int counts = 0;
  const uint8_t *k, *v;
  size_t kl, vl;
  while (r.next(k, kl, v, vl)) {
    ByteReader kr(k, kl), vr(v, vl);
    int32_t idx = kr.i32();
    if (!(idx > 0 && idx <= n_topics))
      throw std::invalid_argument("Invalid alpha index: must be an integer in (0, " +
                                  std::to_string(n_topics) + "]...");
    alpha[(size_t)idx - 1] = vr.f64();
    counts++;
  }
  if (counts != n_topics) throw std::invalid_argument("Invalid alpha vector...");
  return alpha;
}

// The TermReducer output (TermReducer.java:173-195,232-235): one record per topic,
// key PairOfIntFloat(topic 1..K, (float) normaliser), value HMapIDW{term -> digamma(lambda)}.
inline void write_beta(const std::string &path, int K, int V, const double *dl, const float *nrm,
                       const uint8_t *present) {
  SeqWriter w(path, kPairOfIntFloat, kHMapIDW);
  int32_t seen = 0;
  for (int v = 0; v < V; v++) seen += present[v] ? 1 : 0;
  for (int k = 0; k < K; k++) {
    ByteWriter key, val;
    key.i32(k + 1);
    key.f32(nrm[k]);
    val.buf.reserve((size_t)seen * 12 + 4);
    val.i32(seen);
    for (int v = 0; v < V; v++)
      if (present[v]) {
        val.i32(v + 1);
        val.f64(dl[(size_t)v * K + k]);
      }
    w.append(key, val);
  }
}

// importBeta's parse (DocumentMapper.java:475-536) into the arrays mrlda_set_beta takes.
inline void read_beta(const std::string &path, int K, int V, std::vector<double> &dl,
                      std::vector<float> &nrm, std::vector<uint8_t> &present) {
  SeqReader r(path);
  if (r.key_class != kPairOfIntFloat || r.value_class != kHMapIDW)
    throw std::invalid_argument(path + ": not a beta file");
  dl.assign((size_t)V * K, 0.0);
  nrm.assign((size_t)K, 0.0f);
  present.assign((size_t)V, 0);
  std::vector<uint8_t> cell((size_t)V * K, 0);
  const uint8_t *k, *v;
  size_t kl, vl;
  while (r.next(k, kl, v, vl)) {
    ByteReader kr(k, kl), vr(v, vl);
    int32_t topic = kr.i32();
    float norm = kr.f32();
    if (!(topic > 0 && topic <= K))
      throw std::invalid_argument("Invalid beta vector for term " + std::to_string(topic) + "...");
    nrm[(size_t)topic - 1] = norm;
    int32_t n = vr.i32();
    for (int32_t i = 0; i < n; i++) {
      int32_t term = vr.i32();
      double val = vr.f64();
      if (term < 1 || term > V)
        throw std::invalid_argument("beta file: term id " + std::to_string(term) + " out of range");
      size_t at = (size_t)(term - 1) * K + (size_t)(topic - 1);
      if (cell[at])
        throw std::invalid_argument("Dual initialization for term " + std::to_string(term) +
                                    " in topic " + std::to_string(topic - 1) + "...");
      cell[at] = 1;
      dl[at] = val;
      present[(size_t)term - 1] = 1;
    }
  }
}

// The "gamma" named output: the input documents with the updated gamma embedded
// (DocumentMapper.java:342-346).  Documents with null content are not emitted (:197-201).
inline void write_gamma(const std::string &file, const Corpus &c, int64_t lo, int64_t hi,
                        const double *gamma /* (hi-lo) x K */, int K) {
  SeqWriter w(file, kIntWritable, kDocument);
  Document doc;
  for (int64_t d = lo; d < hi; d++) {
    int64_t b = c.row_ptr[d], e = c.row_ptr[d + 1];
    if (e == b) continue;
    doc.content.clear();
    for (int64_t z = b; z < e; z++) doc.content.emplace_back(c.term_id[z], c.count[z]);
    doc.gamma.assign(gamma + (size_t)(d - lo) * K, gamma + (size_t)(d - lo + 1) * K);
    ByteWriter key, val;
    key.i32(c.doc_id[d]);
    doc.write(val);
    w.append(key, val);
  }
}

// The eta file of -informedprior: SequenceFile<IntWritable topic (1-based), ArrayListOfIntsWritable
// seed term ids> as written by InformedPrior.exportTerms (InformedPrior.java:139-170) and read by
// importEta (:179-202).  ArrayListOfIntsWritable = int32 size, then size int32 values.
static const char *const kArrayListOfInts = "edu.umd.cloud9.io.array.ArrayListOfIntsWritable";
inline void read_eta(const std::string &path, std::vector<int32_t> &topic, std::vector<int32_t> &term) {
  SeqReader r(path);
  const uint8_t *k, *v;
  size_t kl, vl;
  while (r.next(k, kl, v, vl)) {
    ByteReader kr(k, kl), vr(v, vl);
    int32_t t = kr.i32();
    if (t <= 0) throw std::invalid_argument("Invalid eta prior for term " + std::to_string(t) + "...");
    int32_t n = vr.i32();
    for (int32_t i = 0; i < n; i++) {
      topic.push_back(t);
      term.push_back(vr.i32());
    }
  }
}
inline void write_eta(const std::string &path, const std::vector<std::vector<int32_t>> &seeds) {
  SeqWriter w(path, kIntWritable, kArrayListOfInts);
  for (size_t t = 0; t < seeds.size(); t++) {
    ByteWriter k, v;
    k.i32((int32_t)t + 1);
    v.i32((int32_t)seeds[t].size());
    for (int32_t x : seeds[t]) v.i32(x);
    w.append(k, v);
  }
}

// java.lang.Double.toString: shortest digits that round-trip; decimal notation for
// 1e-3 <= |d| < 1e7, otherwise computerised scientific notation.
inline std::string java_double(double d) {
  if (d != d) return "NaN";
  if (d == 1.0 / 0.0) return "Infinity";
  if (d == -1.0 / 0.0) return "-Infinity";
  if (d == 0) return (1.0 / d < 0) ? "-0.0" : "0.0";
  char buf[64];
  int prec = 1;
  for (; prec <= 17; prec++) {
    snprintf(buf, sizeof buf, "%.*e", prec - 1, d);
    if (strtod(buf, nullptr) == d) break;
  }
  std::string s(buf);
  size_t epos = s.find('e');
  std::string mant = s.substr(0, epos);
  int exp10 = atoi(s.c_str() + epos + 1);
  bool neg = mant[0] == '-';
  if (neg) mant = mant.substr(1);
  std::string digits;
  for (char ch : mant)
    if (ch != '.') digits.push_back(ch);
  std::string out;
  double ad = d < 0 ? -d : d;
  if (ad >= 1e-3 && ad < 1e7) {
    if (exp10 >= 0) {
      std::string ip = digits.substr(0, std::min(digits.size(), (size_t)exp10 + 1));
      while ((int)ip.size() < exp10 + 1) ip.push_back('0');
      std::string fp = digits.size() > (size_t)exp10 + 1 ? digits.substr((size_t)exp10 + 1) : "0";
      out = ip + "." + fp;
    } else {
      out = "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
    }
  } else {
    out = digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : "0") + "E" +
          std::to_string(exp10);
  }
  return (neg ? "-" : "") + out;
}

}  // namespace mrlda_host
