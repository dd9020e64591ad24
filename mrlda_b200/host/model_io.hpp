// model_io.hpp — reading the ParseCorpus document files into CSR and reading / writing the
// alpha-N, beta-N and gamma-N model files of cc.mrlda.VariationalInference, with the reference's
// validity checks (Preconditions.checkArgument -> std::invalid_argument).
#pragma once
#include <dirent.h>
#include <sys/mman.h>
#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <exception>
#include <string>
#include <thread>
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

// How many host threads the file readers / writers may use (part files are independent streams).
inline int io_threads(size_t jobs) {
  unsigned hw = std::thread::hardware_concurrency();
  if (const char *e = getenv("MRLDA_IO_THREADS")) hw = (unsigned)std::max(1, atoi(e));
  return (int)std::max<size_t>(1, std::min<size_t>({jobs, (size_t)std::max(1u, hw), (size_t)16}));
}

// Runs job(0..n-1) on up to io_threads(n) threads; the first failure (in job order) is rethrown.
template <class F>
inline void parallel_jobs(size_t n, F job) {
  const int T = io_threads(n);
  if (T <= 1) {
    for (size_t i = 0; i < n; i++) job(i);
    return;
  }
  std::vector<std::exception_ptr> err(n);
  std::atomic<size_t> next{0};
  std::vector<std::thread> pool;
  for (int t = 0; t < T; t++)
    pool.emplace_back([&] {
      for (size_t i; (i = next.fetch_add(1)) < n;) {
        try {
          job(i);
        } catch (...) {
          err[i] = std::current_exception();
        }
      }
    });
  for (auto &th : pool) th.join();
  for (auto &e : err)
    if (e) std::rethrow_exception(e);
}

// Big arrays are sized once and backed by transparent huge pages where the kernel offers them on
// request: first-touch page faults of 4 KB pages are serialised per process and would otherwise
// bound the reader (0.12 s per 256 MB here, against 0.04 s with 2 MB pages).
template <class T>
inline void alloc_huge(std::vector<T> &v, size_t n) {
  v.reserve(n);
#ifdef MADV_HUGEPAGE
  const size_t bytes = v.capacity() * sizeof(T), two_mb = (size_t)2 << 20;
  if (bytes >= 4 * two_mb) {
    const uintptr_t lo = ((uintptr_t)v.data() + two_mb - 1) & ~(uintptr_t)(two_mb - 1);
    const uintptr_t hi = ((uintptr_t)v.data() + bytes) & ~(uintptr_t)(two_mb - 1);
    if (hi > lo) madvise((void *)lo, hi - lo, MADV_HUGEPAGE);  // a hint: failure is harmless
  }
#endif
  v.resize(n);
}

// First pass over one part file: per record the entry count at the head of the Document value and
// whether the record carries a gamma of exactly n_topics values.
struct FileScan {
  std::vector<int32_t> entries;
  std::vector<uint8_t> k_gamma;
};
inline void check_document_file(const SeqReader &r, const std::string &file) {
  if (r.key_class != kIntWritable || r.value_class != kDocument)
    throw std::invalid_argument(file + ": expected <" + kIntWritable + ", " + kDocument + ">, found <" +
                                r.key_class + ", " + r.value_class + ">");
}
inline FileScan scan_corpus_file(const std::string &file, int n_topics) {
  SeqReader r(file);
  check_document_file(r, file);
  FileScan s;
  const uint8_t *k, *v;
  size_t kl, vl;
  while (r.next(k, kl, v, vl)) {
    ByteReader vr(v, vl);
    const int32_t n = vr.i32();
    if (n > 0) {
      vr.need((size_t)n * 8);
      vr.p += (size_t)n * 8;
    }
    const int32_t kk = vr.i32();
    s.entries.push_back(n > 0 ? n : 0);
    s.k_gamma.push_back(kk == n_topics && kk > 0 ? 1 : 0);
  }
  return s;
}

// Second pass over one part file: the records with file-local index in [skip, skip + keep) are
// parsed straight into the final arrays — document i of them at row doc_base + i, its entries from
// entry_base on.  Entry order within a document is the writer's HMapII iteration order (arbitrary)
// and is kept as found; a repeated term id keeps its first position and the last count
// (HMapII.put), which leaves the file's entries short of the scanned count: the number actually
// written is returned and the caller closes the gap.  gamma (may be null) receives D x K values.
inline int64_t parse_corpus_file(const std::string &file, const FileScan &scan, int n_topics, int n_terms,
                                 int64_t skip, int64_t keep, int64_t doc_base, int64_t entry_base,
                                 Corpus &c, double *gamma) {
  SeqReader r(file);
  check_document_file(r, file);
  const uint8_t *k, *v;
  size_t kl, vl;
  TermSlots slots;
  int64_t at = entry_base;
  for (int64_t rec = 0; rec < skip + keep; rec++) {
    if (!r.next(k, kl, v, vl)) throw std::runtime_error(file + ": changed while it was being read");
    if (rec < skip) continue;
    const int64_t d = doc_base + (rec - skip);
    ByteReader kr(k, kl), vr(v, vl);
    const int32_t doc = kr.i32();
    c.doc_id[(size_t)d] = doc;
    const int32_t n = vr.i32();
    if ((n > 0 ? n : 0) != scan.entries[(size_t)rec])
      throw std::runtime_error(file + ": changed while it was being read");
    if (n > 0) {  // numEntries <= 0  =>  content = null
      vr.need((size_t)n * 8);
      slots.begin((size_t)n);
      for (int32_t i = 0; i < n; i++) {
        const int32_t id = ByteReader::get32(vr.p), cnt = ByteReader::get32(vr.p + 4);
        vr.p += 8;
        if (id < 1 || id > n_terms)
          throw std::invalid_argument("term id " + std::to_string(id) + " outside [1," +
                                      std::to_string(n_terms) + "] in document " + std::to_string(doc));
        const int64_t seen = slots.find_or_add(id, at);
        if (seen < 0) {
          c.term_id[(size_t)at] = id;
          c.count[(size_t)at] = cnt;
          at++;
        } else {
          c.count[(size_t)seen] = cnt;
        }
      }
    }
    c.row_ptr[(size_t)d + 1] = at;
    if (gamma) {
      const int32_t kk = vr.i32();
      if (kk != n_topics) throw std::runtime_error(file + ": changed while it was being read");
      vr.need((size_t)kk * 8);
      double *g = gamma + (size_t)d * (size_t)n_topics;
      for (int32_t i = 0; i < kk; i++) {
        const int64_t bits = ByteReader::get64(vr.p + (size_t)i * 8);
        memcpy(&g[i], &bits, 8);
      }
    }
  }
  return at - entry_base;
}

// SequenceFile<IntWritable, cc.mrlda.Document> directory -> CSR: the documents with global index in
// [first, last), in part-file name order.  Two passes, each taking the part files concurrently: the
// first reads the entry count of every record (`scans`, when the caller has done that already),
// which fixes every array size and every part file's place in them; the second parses each file's
// records straight into place.  Gamma is kept only when every kept document carries a vector of
// n_topics values (DocumentMapper.java:184: a stored gamma of another length is ignored).
inline Corpus read_corpus(const std::string &path, int n_topics, int n_terms, int64_t first = 0,
                          int64_t last = INT64_MAX, const std::vector<FileScan> *scans = nullptr) {
  const std::vector<std::string> files = input_files(path);
  std::vector<FileScan> own;
  if (!scans || scans->size() != files.size()) {
    own.resize(files.size());
    parallel_jobs(files.size(), [&](size_t f) { own[f] = scan_corpus_file(files[f], n_topics); });
    scans = &own;
  }
  const size_t F = files.size();
  std::vector<int64_t> skip(F, 0), keep(F, 0), doc_base(F, 0), entry_base(F, 0), counted(F, 0);
  int64_t index = 0, D = 0, Z = 0;
  bool all_gamma = true;
  if (first < 0) first = 0;
  for (size_t f = 0; f < F; f++) {
    const FileScan &s = (*scans)[f];
    const int64_t n = (int64_t)s.entries.size();
    const int64_t a = std::min(std::max(first - index, (int64_t)0), n);
    const int64_t b = std::min(std::max(last - index, (int64_t)0), n);
    skip[f] = a;
    keep[f] = std::max(b - a, (int64_t)0);
    doc_base[f] = D;
    entry_base[f] = Z;
    for (int64_t i = a; i < a + keep[f]; i++) {
      counted[f] += s.entries[(size_t)i];
      if (!s.k_gamma[(size_t)i]) all_gamma = false;
    }
    D += keep[f];
    Z += counted[f];
    index += n;
  }
  Corpus c;
  alloc_huge(c.row_ptr, (size_t)D + 1);
  alloc_huge(c.doc_id, (size_t)D);
  alloc_huge(c.term_id, (size_t)Z);
  alloc_huge(c.count, (size_t)Z);
  if (all_gamma && D > 0) alloc_huge(c.gamma, (size_t)D * (size_t)n_topics);
  c.row_ptr[0] = 0;
  std::vector<int64_t> used(F, 0);
  parallel_jobs(F, [&](size_t f) {
    if (keep[f] == 0) return;
    used[f] = parse_corpus_file(files[f], (*scans)[f], n_topics, n_terms, skip[f], keep[f], doc_base[f],
                                entry_base[f], c, c.gamma.empty() ? nullptr : c.gamma.data());
  });
  // repeated term ids inside a record left a file short of its scanned entry count: close the gaps
  int64_t gap = 0;
  for (size_t f = 0; f < F; f++) {
    if (gap > 0 && keep[f] > 0) {
      memmove(&c.term_id[(size_t)(entry_base[f] - gap)], &c.term_id[(size_t)entry_base[f]], (size_t)used[f] * 4);
      memmove(&c.count[(size_t)(entry_base[f] - gap)], &c.count[(size_t)entry_base[f]], (size_t)used[f] * 4);
      for (int64_t d = doc_base[f]; d < doc_base[f] + keep[f]; d++) c.row_ptr[(size_t)d + 1] -= gap;
    }
    gap += counted[f] - used[f];
  }
  c.term_id.resize((size_t)(Z - gap));
  c.count.resize((size_t)(Z - gap));
  return c;
}

// The input split of one rank (the map phase's input splits): the first pass (entry counts of every
// record, part files taken concurrently) cuts the document range into `world` contiguous pieces
// balanced by entries, as corpus.shard_bounds does; the second pass parses this rank's piece alone,
// so no rank ever holds more than its own shard.  first_doc receives the global index of the
// shard's first document.
inline Corpus read_corpus_shard(const std::string &path, int n_topics, int n_terms, int rank, int world,
                                int64_t *first_doc = nullptr, int64_t *total_docs = nullptr) {
  if (world <= 1) {
    Corpus c = read_corpus(path, n_topics, n_terms);
    if (first_doc) *first_doc = 0;
    if (total_docs) *total_docs = c.n_docs();
    return c;
  }
  const std::vector<std::string> files = input_files(path);
  std::vector<FileScan> scans(files.size());
  parallel_jobs(files.size(), [&](size_t f) { scans[f] = scan_corpus_file(files[f], n_topics); });
  std::vector<int64_t> rp{0};
  for (auto &s : scans)
    for (int32_t n : s.entries) rp.push_back(rp.back() + n);
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
  return read_corpus(path, n_topics, n_terms, lo, hi, &scans);
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
  w.close();
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
// dl is term-major (K contiguous), a record is one topic's column: eight topics (one cache line of
// a row) are serialised together so that every line of dl is read once per eight records.  All
// records have one size, so their file positions (sync escapes included) are known in advance
// (SeqWriter::plan): the tiles are built concurrently and each stores its own bytes with pwrite —
// the file is byte for byte what appending the records one after another gives (and that is what
// happens with a single thread).
inline void write_beta(const std::string &path, int K, int V, const double *dl, const float *nrm,
                       const uint8_t *present) {
  SeqWriter w(path, kPairOfIntFloat, kHMapIDW);
  std::vector<int32_t> terms;  // the terms that have a row (ascending)
  for (int v = 0; v < V; v++)
    if (present[v]) terms.push_back(v);
  const size_t seen = terms.size(), val_bytes = 4 + 12 * seen, key_bytes = 8, rec_bytes = 8 + key_bytes + val_bytes;
  const int kTile = 8, tiles = (K + kTile - 1) / kTile;
  // fills dst[i] (val_bytes each) with the HMapIDW values of topics k0 .. k0 + kn - 1
  auto build = [&](int k0, int kn, uint8_t *const *dst) {
    uint8_t *q[kTile];
    for (int i = 0; i < kn; i++) {
      q[i] = dst[i];
      ByteWriter::put32(q[i], (int32_t)seen);
      q[i] += 4;
    }
    for (size_t s = 0; s < seen; s++) {
      const int32_t v = terms[s];
      const double *row = dl + (size_t)v * K + k0;
      for (int i = 0; i < kn; i++) {
        int64_t bits;
        memcpy(&bits, &row[i], 8);
        ByteWriter::put32(q[i], v + 1);
        ByteWriter::put64(q[i] + 4, bits);
        q[i] += 12;
      }
    }
  };
  if (io_threads((size_t)tiles) <= 1) {
    std::vector<ByteWriter> val((size_t)kTile);
    for (int t = 0; t < tiles; t++) {
      const int k0 = t * kTile, kn = std::min(kTile, K - k0);
      uint8_t *dst[kTile];
      for (int i = 0; i < kn; i++) {
        val[(size_t)i].buf.clear();
        dst[i] = val[(size_t)i].grow(val_bytes);
      }
      build(k0, kn, dst);
      for (int i = 0; i < kn; i++) {
        ByteWriter key;
        key.i32(k0 + i + 1);
        key.f32(nrm[k0 + i]);
        w.append(key, val[(size_t)i]);
      }
    }
  } else {
    const std::vector<SeqWriter::Slot> slot = w.plan((size_t)K, rec_bytes);
    parallel_jobs((size_t)tiles, [&](size_t t) {
      const int k0 = (int)t * kTile, kn = std::min(kTile, K - k0);
      // one buffer from the tile's first byte (its sync escape, if any) to the end of its last record
      const int64_t from = slot[(size_t)k0].at - (slot[(size_t)k0].sync_before ? 20 : 0);
      const int64_t to = slot[(size_t)(k0 + kn - 1)].at + (int64_t)rec_bytes;
      std::vector<uint8_t> buf((size_t)(to - from));
      uint8_t *dst[kTile];
      for (int i = 0; i < kn; i++) {
        const SeqWriter::Slot &sl = slot[(size_t)(k0 + i)];
        uint8_t *rec = buf.data() + (sl.at - from);
        if (sl.sync_before) w.sync_escape(rec - 20);
        ByteWriter::put32(rec, (int32_t)(key_bytes + val_bytes));
        ByteWriter::put32(rec + 4, (int32_t)key_bytes);
        ByteWriter::put32(rec + 8, k0 + i + 1);
        int32_t fbits;
        memcpy(&fbits, &nrm[k0 + i], 4);
        ByteWriter::put32(rec + 12, fbits);
        dst[i] = rec + 16;
      }
      build(k0, kn, dst);
      w.write_at(from, buf.data(), buf.size());
    });
  }
  w.close();
}

// importBeta's parse (DocumentMapper.java:475-536) into the arrays mrlda_set_beta takes.  A record
// is one topic's column of the term-major matrix, so storing it entry by entry touches one cache
// line per value: records are taken eight at a time and walked in lockstep (HMapIDW files written
// from one key set list the terms in one order), which puts up to eight values into each line, and
// the tiles of a wave are scattered concurrently (distinct topics: distinct cells).  A topic that
// appears twice is handled on its own so that "Dual initialization" is reported as the reference
// reports it.
inline void read_beta(const std::string &path, int K, int V, std::vector<double> &dl,
                      std::vector<float> &nrm, std::vector<uint8_t> &present) {
  SeqReader r(path);
  if (r.key_class != kPairOfIntFloat || r.value_class != kHMapIDW)
    throw std::invalid_argument(path + ": not a beta file");
  dl.clear();
  alloc_huge(dl, (size_t)V * K);
  nrm.assign((size_t)K, 0.0f);
  present.assign((size_t)V, 0);
  std::vector<uint8_t> cell;
  alloc_huge(cell, (size_t)V * K);
  struct Rec {
    int32_t topic = 0, n = 0;
    const uint8_t *data = nullptr;  // n x (int32 term, float64 value)
    std::vector<uint8_t> own;       // the copy `data` points to when the record waits for its wave
  };
  const int kTile = 8;
  const bool direct = io_threads(64) <= 1;  // one thread: store every record as it is read, no copy
  const size_t wave = direct ? 1 : (size_t)io_threads(64) * kTile;
  std::vector<Rec> recs(wave);
  std::vector<uint8_t> topic_seen((size_t)K, 0);
  size_t have = 0;
  auto store = [&](const Rec *tile, int cnt) {  // the records of one tile, in lockstep where they agree
    int32_t n_min = tile[0].n;
    for (int i = 1; i < cnt; i++) n_min = std::min(n_min, tile[i].n);
    for (int32_t s = 0; s < n_min; s++) {
      for (int i = 0; i < cnt; i++) {
        const uint8_t *e = tile[i].data + (size_t)s * 12;
        const int32_t term = ByteReader::get32(e);
        if (term < 1 || term > V)
          throw std::invalid_argument("beta file: term id " + std::to_string(term) + " out of range");
        const size_t at = (size_t)(term - 1) * K + (size_t)(tile[i].topic - 1);
        if (cell[at])
          throw std::invalid_argument("Dual initialization for term " + std::to_string(term) +
                                      " in topic " + std::to_string(tile[i].topic - 1) + "...");
        cell[at] = 1;
        const int64_t bits = ByteReader::get64(e + 4);
        memcpy(&dl[at], &bits, 8);
      }
    }
    for (int i = 0; i < cnt; i++)  // what the longer records hold beyond the shortest
      for (int32_t s = n_min; s < tile[i].n; s++) {
        const uint8_t *e = tile[i].data + (size_t)s * 12;
        const int32_t term = ByteReader::get32(e);
        if (term < 1 || term > V)
          throw std::invalid_argument("beta file: term id " + std::to_string(term) + " out of range");
        const size_t at = (size_t)(term - 1) * K + (size_t)(tile[i].topic - 1);
        if (cell[at])
          throw std::invalid_argument("Dual initialization for term " + std::to_string(term) +
                                      " in topic " + std::to_string(tile[i].topic - 1) + "...");
        cell[at] = 1;
        const int64_t bits = ByteReader::get64(e + 4);
        memcpy(&dl[at], &bits, 8);
      }
  };
  auto flush = [&] {
    const size_t tiles = (have + kTile - 1) / kTile;
    parallel_jobs(tiles, [&](size_t t) {
      store(&recs[t * kTile], (int)std::min<size_t>(kTile, have - t * kTile));
    });
    have = 0;
  };
  const uint8_t *k, *v;
  size_t kl, vl;
  while (r.next(k, kl, v, vl)) {
    ByteReader kr(k, kl), vr(v, vl);
    const int32_t topic = kr.i32();
    const float norm = kr.f32();
    if (!(topic > 0 && topic <= K))
      throw std::invalid_argument("Invalid beta vector for term " + std::to_string(topic) + "...");
    nrm[(size_t)topic - 1] = norm;
    const int32_t n = vr.i32();
    if (n < 0) continue;  // an empty map
    vr.need((size_t)n * 12);
    const bool repeat = topic_seen[(size_t)topic - 1] != 0;
    topic_seen[(size_t)topic - 1] = 1;
    if (repeat) flush();  // its cells may collide with a record of the wave: keep the order
    Rec &rec = recs[have++];
    rec.topic = topic;
    rec.n = n;
    if (direct) {
      rec.data = vr.p;  // valid until the next r.next(): stored right away
    } else {
      rec.own.assign(vr.p, vr.p + (size_t)n * 12);
      rec.data = rec.own.data();
    }
    if (repeat || have == wave) flush();
  }
  flush();
  for (int v = 0; v < V; v++)  // a term is present when any topic's record listed it
    present[(size_t)v] = memchr(&cell[(size_t)v * K], 1, (size_t)K) ? 1 : 0;
}

// The "gamma" named output: the input documents with the updated gamma embedded
// (DocumentMapper.java:342-346).  Documents with null content are not emitted (:197-201).
inline void write_gamma(const std::string &file, const Corpus &c, int64_t lo, int64_t hi,
                        const double *gamma /* (hi-lo) x K */, int K) {
  SeqWriter w(file, kIntWritable, kDocument);
  ByteWriter key, val;
  for (int64_t d = lo; d < hi; d++) {
    const int64_t b = c.row_ptr[d], e = c.row_ptr[d + 1];
    if (e == b) continue;
    key.buf.clear();
    val.buf.clear();
    key.i32(c.doc_id[d]);
    // Document.write (Document.java:241-263), straight from the CSR row and the gamma row
    uint8_t *q = val.grow(8 + 8 * (size_t)(e - b) + 8 * (size_t)K);
    ByteWriter::put32(q, (int32_t)(e - b));
    q += 4;
    for (int64_t z = b; z < e; z++) {
      ByteWriter::put32(q, c.term_id[z]);
      ByteWriter::put32(q + 4, c.count[z]);
      q += 8;
    }
    ByteWriter::put32(q, K);
    q += 4;
    const double *g = gamma + (size_t)(d - lo) * K;
    for (int k = 0; k < K; k++) {
      int64_t bits;
      memcpy(&bits, &g[k], 8);
      ByteWriter::put64(q, bits);
      q += 8;
    }
    w.append(key, val);
  }
  w.close();
}

// Part files one rank may write per gamma directory: rank r owns the indices [r * 32, r * 32 + 32),
// so the name order of a directory is the document order whatever the rank count was.
static const int kGammaPartsPerRank = 32;

// The gamma output of one rank as several part files written concurrently, the way the map phase
// writes one "gamma_gamma-m-NNNNN" per map task (VariationalInference.java:232-235,366-373): the
// rank's documents are cut into contiguous pieces of about `bytes_per_part` bytes (at most
// kGammaPartsPerRank, at most one per host thread), piece t going to index rank * 32 + t.  Returns
// the number of part files.
inline int write_gamma_parts(const std::string &dir, int rank, const Corpus &c, int64_t lo, int64_t hi,
                             const double *gamma /* (hi-lo) x K */, int K,
                             int64_t bytes_per_part = (int64_t)64 << 20) {
  auto bytes_upto = [&](int64_t d) {  // serialized size of documents [lo, d), null-content ones included
    return (c.row_ptr[d] - c.row_ptr[lo]) * 8 + (d - lo) * ((int64_t)K * 8 + 28);
  };
  const int64_t total = bytes_upto(hi);
  const int parts = (int)std::max<int64_t>(
      1, std::min<int64_t>({total / std::max<int64_t>(1, bytes_per_part) + 1, (int64_t)kGammaPartsPerRank,
                            (int64_t)io_threads(kGammaPartsPerRank), std::max<int64_t>(1, hi - lo)}));
  std::vector<int64_t> cut((size_t)parts + 1, hi);
  cut[0] = lo;
  for (int t = 1; t < parts; t++) {  // first document at which t/parts of the bytes are behind us
    int64_t a = cut[(size_t)t - 1], b = hi;
    const int64_t want = total * t / parts;
    while (a < b) {
      const int64_t m = (a + b) / 2;
      if (bytes_upto(m) < want) a = m + 1; else b = m;
    }
    cut[(size_t)t] = a;
  }
  parallel_jobs((size_t)parts, [&](size_t t) {
    char name[64];
    snprintf(name, sizeof name, "/gamma_gamma-m-%05d", rank * kGammaPartsPerRank + (int)t);
    write_gamma(dir + name, c, cut[t], cut[t + 1], gamma + (size_t)(cut[t] - lo) * K, K);
  });
  return parts;
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
  w.close();
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
