// mrlda_display.cpp — native counterparts of cc.mrlda.DisplayTopic and cc.mrlda.DisplayDocument
// (SURVEY.md section 8f-3): they read the beta-N / gamma-N files the way the reference's inspection
// tools do, which cross-checks that what mrlda-vi writes is readable by the reference.
//
//   mrlda-display topic -input <beta file> -index <term index file> [-topdisplay N]
//       DisplayTopic.java:97-142: per topic the N terms with the largest values, "term\t\tvalue"
//   mrlda-display document -input <gamma directory>
//       DisplayDocument.java:87-100: "docID gamma_1 ... gamma_K " per line
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>

#include "model_io.hpp"

using namespace mrlda_host;

static int display_topic(const std::string &beta, const std::string &index, int top) {
  std::map<int32_t, std::string> term_index;
  {
    SeqReader r(index);  // SequenceFile<IntWritable, Text> written by ParseCorpus (ParseCorpus.java:492-560)
    const uint8_t *k, *v;
    size_t kl, vl;
    while (r.next(k, kl, v, vl)) {
      ByteReader kr(k, kl), vr(v, vl);
      int32_t id = kr.i32();
      term_index[id] = vr.text();
    }
  }
  SeqReader r(beta);
  if (r.key_class != kPairOfIntFloat || r.value_class != kHMapIDW)
    throw std::invalid_argument("Invalid beta path...");
  const uint8_t *k, *v;
  size_t kl, vl;
  while (r.next(k, kl, v, vl)) {
    ByteReader kr(k, kl), vr(v, vl);
    int32_t topic = kr.i32();
    std::map<double, int32_t> tree;  // TreeMap<Double, Integer> keyed by -value
    printf("==============================\n");
    printf("Top ranked %d terms for Topic %d\n", top, topic);
    printf("==============================\n");
    int32_t n = vr.i32();
    for (int32_t i = 0; i < n; i++) {
      int32_t term = vr.i32();
      double val = vr.f64();
      tree[-val] = term;
      if ((int)tree.size() > top) tree.erase(std::prev(tree.end()));
    }
    for (auto &e : tree) {
      auto it = term_index.find(e.second);
      if (it != term_index.end())
        printf("%s\t\t%s\n", it->second.c_str(), java_double(-e.first).c_str());
      else
        printf("How embarrassing! Term index not found...\n");
    }
  }
  return 0;
}

static int display_document(const std::string &gamma_dir) {
  if (!is_dir(gamma_dir)) throw std::invalid_argument("Invalid gamma path...");
  for (const std::string &file : input_files(gamma_dir)) {
    SeqReader r(file);
    const uint8_t *k, *v;
    size_t kl, vl;
    Document doc;
    while (r.next(k, kl, v, vl)) {
      ByteReader kr(k, kl), vr(v, vl);
      int32_t id = kr.i32();
      doc.read(vr);
      if (doc.gamma.empty())
        throw std::invalid_argument("topic distribution not specified for document " + std::to_string(id));
      printf("%d ", id);
      for (double g : doc.gamma) printf("%s ", java_double(g).c_str());
      printf("\n");
    }
  }
  return 0;
}

int main(int argc, char **argv) {
  std::string mode = argc > 1 ? argv[1] : "", input, index;
  int top = 10;  // DisplayTopic.TOP_DISPLAY
  for (int i = 2; i < argc; i++) {
    std::string a = argv[i];
    while (!a.empty() && a[0] == '-') a = a.substr(1);
    if ((a == "input" || a == "index" || a == "topdisplay") && i + 1 < argc) {
      std::string val = argv[++i];
      if (a == "input") input = val;
      else if (a == "index") index = val;
      else top = atoi(val.c_str());
    }
  }
  try {
    if (mode == "topic" && !input.empty() && !index.empty()) return display_topic(input, index, top);
    if (mode == "document" && !input.empty()) return display_document(input);
  } catch (const std::exception &e) {
    fprintf(stderr, "ERROR %s\n", e.what());
    return 1;
  }
  fprintf(stderr, "usage: mrlda-display topic -input <beta> -index <term index> [-topdisplay N]\n"
                  "       mrlda-display document -input <gamma dir>\n");
  return 0;
}
