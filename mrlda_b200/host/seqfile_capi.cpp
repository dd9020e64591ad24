// seqfile_capi.cpp — C entry points over seqfile.hpp / model_io.hpp so that the tests (ctypes) can
// exercise the same codecs the native driver uses.  Errors: negative return, text via
// mrlda_io_last_error().
#include <cstring>
#include <string>

#include "model_io.hpp"

using namespace mrlda_host;

static thread_local std::string g_err;
#define GUARD(...)                    \
  try {                               \
    __VA_ARGS__                       \
  } catch (const std::exception &e) { \
    g_err = e.what();                 \
    return -1;                        \
  }

extern "C" {

const char *mrlda_io_last_error(void) { return g_err.c_str(); }

// Document.write (Document.java:241-263).  Returns the number of bytes, or -1 if cap is too small.
int64_t mrlda_doc_serialize(int32_t n, const int32_t *term, const int32_t *count, int32_t k,
                            const double *gamma, uint8_t *out, int64_t cap) {
  GUARD(Document d; for (int i = 0; i < n; i++) d.content.emplace_back(term[i], count[i]);
        if (k > 0) d.gamma.assign(gamma, gamma + k); ByteWriter w; d.write(w);
        if ((int64_t)w.buf.size() > cap) { g_err = "buffer too small"; return -1; }
        memcpy(out, w.buf.data(), w.buf.size()); return (int64_t)w.buf.size();)
}

// Document.readFields (Document.java:147-172).  n_out / k_out receive the entry and topic counts.
int mrlda_doc_deserialize(const uint8_t *buf, int64_t len, int32_t *term, int32_t *count,
                          int32_t cap_n, int32_t *n_out, double *gamma, int32_t cap_k,
                          int32_t *k_out, int64_t *tokens_out) {
  GUARD(ByteReader r(buf, (size_t)len); Document d; d.read(r);
        if ((int)d.content.size() > cap_n || (int)d.gamma.size() > cap_k) { g_err = "buffer too small"; return -1; }
        for (size_t i = 0; i < d.content.size(); i++) { term[i] = d.content[i].first; count[i] = d.content[i].second; }
        for (size_t i = 0; i < d.gamma.size(); i++) gamma[i] = d.gamma[i];
        *n_out = (int32_t)d.content.size(); *k_out = (int32_t)d.gamma.size(); *tokens_out = d.tokens(); return 0;)
}

// writes a SequenceFile<IntWritable, cc.mrlda.Document> (what ParseCorpus / the gamma output hold)
int mrlda_io_write_corpus(const char *path, int64_t n_docs, const int64_t *row_ptr,
                          const int32_t *term, const int32_t *count, const int32_t *doc_id,
                          const double *gamma, int32_t k) {
  GUARD(SeqWriter w(path, kIntWritable, kDocument); Document d;
        for (int64_t i = 0; i < n_docs; i++) {
          d.content.clear(); d.gamma.clear();
          for (int64_t z = row_ptr[i]; z < row_ptr[i + 1]; z++) d.content.emplace_back(term[z], count[z]);
          if (gamma && k > 0) d.gamma.assign(gamma + (size_t)i * k, gamma + (size_t)(i + 1) * k);
          ByteWriter key, val; key.i32(doc_id ? doc_id[i] : (int32_t)i + 1); d.write(val); w.append(key, val);
        } w.close(); return 0;)
}

// the gamma output of one rank as concurrently written part files (write_gamma_parts); returns the
// number of part files
int mrlda_io_write_gamma_parts(const char *dir, int32_t rank, int64_t n_docs, const int64_t *row_ptr,
                               const int32_t *term, const int32_t *count, const int32_t *doc_id,
                               const double *gamma, int32_t k, int64_t bytes_per_part) {
  GUARD(Corpus c; c.row_ptr.assign(row_ptr, row_ptr + n_docs + 1);
        c.term_id.assign(term, term + row_ptr[n_docs]); c.count.assign(count, count + row_ptr[n_docs]);
        c.doc_id.resize((size_t)n_docs);
        for (int64_t i = 0; i < n_docs; i++) c.doc_id[(size_t)i] = doc_id ? doc_id[i] : (int32_t)i + 1;
        return write_gamma_parts(dir, rank, c, 0, n_docs, gamma, k, bytes_per_part);)
}

struct mrlda_io_corpus { Corpus c; };

mrlda_io_corpus *mrlda_io_read_corpus(const char *path, int32_t k, int32_t v) {
  try {
    auto *h = new mrlda_io_corpus();
    h->c = read_corpus(path, k, v);
    return h;
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}
// the input split of one rank (read_corpus_shard): first_doc receives the shard's first global index
mrlda_io_corpus *mrlda_io_read_corpus_shard(const char *path, int32_t k, int32_t v, int32_t rank,
                                            int32_t world, int64_t *first_doc, int64_t *total_docs) {
  try {
    auto *h = new mrlda_io_corpus();
    h->c = read_corpus_shard(path, k, v, rank, world, first_doc, total_docs);
    return h;
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}
int64_t mrlda_io_corpus_docs(const mrlda_io_corpus *h) { return h->c.n_docs(); }
int64_t mrlda_io_corpus_nnz(const mrlda_io_corpus *h) { return (int64_t)h->c.term_id.size(); }
int mrlda_io_corpus_has_gamma(const mrlda_io_corpus *h) { return h->c.gamma.empty() ? 0 : 1; }
void mrlda_io_corpus_copy(const mrlda_io_corpus *h, int64_t *row_ptr, int32_t *term, int32_t *count,
                          int32_t *doc_id, double *gamma) {
  const Corpus &c = h->c;
  memcpy(row_ptr, c.row_ptr.data(), c.row_ptr.size() * 8);
  if (!c.term_id.empty()) {
    memcpy(term, c.term_id.data(), c.term_id.size() * 4);
    memcpy(count, c.count.data(), c.count.size() * 4);
  }
  if (!c.doc_id.empty()) memcpy(doc_id, c.doc_id.data(), c.doc_id.size() * 4);
  if (gamma && !c.gamma.empty()) memcpy(gamma, c.gamma.data(), c.gamma.size() * 8);
}
void mrlda_io_corpus_free(mrlda_io_corpus *h) { delete h; }

int mrlda_io_write_alpha(const char *path, int32_t k, const double *alpha) {
  GUARD(write_alpha(path, std::vector<double>(alpha, alpha + k)); return 0;)
}
int mrlda_io_read_alpha(const char *path, int32_t k, double *alpha) {
  GUARD(auto a = read_alpha(path, k); memcpy(alpha, a.data(), (size_t)k * 8); return 0;)
}
int mrlda_io_write_beta(const char *path, int32_t k, int32_t v, const double *dl, const float *nrm,
                        const uint8_t *present) {
  GUARD(write_beta(path, k, v, dl, nrm, present); return 0;)
}
int mrlda_io_read_beta(const char *path, int32_t k, int32_t v, double *dl, float *nrm,
                       uint8_t *present) {
  GUARD(std::vector<double> d; std::vector<float> n; std::vector<uint8_t> p; read_beta(path, k, v, d, n, p);
        memcpy(dl, d.data(), d.size() * 8); memcpy(nrm, n.data(), n.size() * 4); memcpy(present, p.data(), p.size());
        return 0;)
}
// term index file of ParseCorpus: SequenceFile<IntWritable, Text> (ids 1..n, '\n'-separated terms)
int mrlda_io_write_term_index(const char *path, int32_t n, const char *terms_nl) {
  GUARD(SeqWriter w(path, kIntWritable, "org.apache.hadoop.io.Text"); const char *p = terms_nl;
        for (int32_t i = 1; i <= n; i++) {
          const char *e = strchr(p, '\n'); std::string t = e ? std::string(p, e) : std::string(p);
          ByteWriter k, v; k.i32(i); v.text(t); w.append(k, v); p = e ? e + 1 : p + t.size();
        } w.close(); return 0;)
}
// eta file of -informedprior: n (topic, term) pairs, topics 1..k
int mrlda_io_write_eta(const char *path, int32_t k, int64_t n, const int32_t *topic, const int32_t *term) {
  GUARD(std::vector<std::vector<int32_t>> seeds((size_t)k);
        for (int64_t i = 0; i < n; i++) seeds.at((size_t)topic[i] - 1).push_back(term[i]);
        write_eta(path, seeds); return 0;)
}
// java.lang.Double.toString, for the likelihood log lines
int mrlda_io_java_double(double d, char *out, int32_t cap) {
  std::string s = java_double(d);
  if ((int)s.size() + 1 > cap) return -1;
  memcpy(out, s.c_str(), s.size() + 1);
  return (int)s.size();
}

}  // extern "C"
