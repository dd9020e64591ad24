// mrlda_vi.cpp — native stand-in for `hadoop jar Mr.LDA.jar cc.mrlda.VariationalInference ...`.
//
// Mirrors VariationalInference.run (VariationalInference.java:70-397) around the hot path: the same
// options (VariationalInferenceOptions.java:55-273), the same model files (alpha-N, beta-N,
// gamma-N/), the same gamma-directory rotation (:359-379), log lines (:257-275,381-384) and
// convergence test (:383-386).  JobClient.runJob + the alpha update are one mrlda_iterate call.
// Extensions (not in the reference): -device, -seed, -rank/-world/-ncclid for one process per GPU,
// -deterministic (fixed-point statistics: bit-identical models for any rank count; slower).
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mrlda_b200.h"
#include "model_io.hpp"

using namespace mrlda_host;

namespace {

struct Options {
  std::string input, output, model_path, informed_prior;
  int topics = 0, terms = 0;
  int iterations = 30;  // Settings.DEFAULT_GLOBAL_MAXIMUM_ITERATION
  int mappers = 100, reducers = 50;
  bool training = true, resume = false, random_start = false, symmetric = false, direct_emit = false;
  int snapshot = 0;
  int device = 0, rank = 0, world = 1;
  unsigned long long seed = 0x4d724c4441ull;
  bool deterministic = false;
  std::string nccl_id_file;
};

void info(const std::string &s) { printf("INFO  VariationalInference: %s\n", s.c_str()); fflush(stdout); }

[[noreturn]] void usage(const std::string &why) {
  if (!why.empty()) fprintf(stderr, "ERROR %s\n", why.c_str());
  fprintf(stderr,
          "usage: cc.mrlda.VariationalInference\n"
          " -directemit            disable in-mapper-combiner (accepted; statistics always sum over all documents)\n"
          " -help                  print the help message\n"
          " -informedprior <path>  seed informed prior (eta file written by cc.mrlda.InformedPrior)\n"
          " -input <path>          input file or directory\n"
          " -iteration <int>       number of iterations (default - 30)\n"
          " -mapper <int>          number of mappers (accepted, ignored)\n"
          " -modelindex <int>      the iteration/index of current model parameters\n"
          " -output <path>         output directory\n"
          " -randomstart           start gamma from random point every iteration\n"
          " -reducer <int>         number of reducers (accepted, ignored)\n"
          " -symmetricalpha        symmetric topic Dirichlet prior\n"
          " -truncatebeta <int>    parsed and unused, as in the reference (VariationalInferenceOptions.java:116-120)\n"
          " -term <int>            number of terms\n"
          " -test <path>           run program in inference mode, i.e. test held-out likelihood\n"
          " -topic <int>           number of topics\n"
          " -device/-seed/-rank/-world/-ncclid  (extensions: GPU ordinal, beta seed, one process per GPU)\n"
          " -deterministic         (extension) fixed-point statistics: the model is bit-identical from run to\n"
          "                        run and for any -world; every document takes the general kernel (slower)\n");
  exit(0);  // the reference prints the help and exits 0 on a parse error (VIO.java:251-266)
}

bool check(bool cond, const std::string &msg) {
  if (!cond) usage(msg);
  return cond;
}

Options parse(int argc, char **argv) {
  Options o;
  bool has_iter = false, has_test = false, has_index = false, has_sym = false;
  std::string iter_arg;
  auto need = [&](int &i) -> std::string {
    if (i + 1 >= argc) usage(std::string("Missing argument for option: ") + argv[i]);
    return argv[++i];
  };
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    while (!a.empty() && a[0] == '-') a = a.substr(1);  // GnuParser accepts -opt and --opt
    if (a == "help") usage("");
    else if (a == "input") o.input = need(i);
    else if (a == "output") o.output = need(i);
    else if (a == "term") o.terms = atoi(need(i).c_str());
    else if (a == "topic") o.topics = atoi(need(i).c_str());
    else if (a == "iteration") { has_iter = true; iter_arg = need(i); }
    else if (a == "mapper") o.mappers = atoi(need(i).c_str());
    else if (a == "reducer") o.reducers = atoi(need(i).c_str());
    else if (a == "test") { has_test = true; o.model_path = need(i); }
    else if (a == "modelindex") { has_index = true; o.snapshot = atoi(need(i).c_str()); }
    else if (a == "randomstart") o.random_start = true;
    else if (a == "symmetricalpha") has_sym = true;
    else if (a == "directemit") o.direct_emit = true;
    else if (a == "truncatebeta") need(i);  // parsed, unused (VIO.java:116-120)
    else if (a == "informedprior") o.informed_prior = need(i);
    else if (a == "device") o.device = atoi(need(i).c_str());
    else if (a == "seed") o.seed = strtoull(need(i).c_str(), nullptr, 10);
    else if (a == "deterministic") o.deterministic = true;
    else if (a == "rank") o.rank = atoi(need(i).c_str());
    else if (a == "world") o.world = atoi(need(i).c_str());
    else if (a == "ncclid") o.nccl_id_file = need(i);
    else usage("Unrecognized option: " + std::string(argv[i]));
  }
  check(!o.input.empty(), "Missing required option: input");
  check(!o.output.empty(), "Missing required option: output");
  if (o.output.back() != '/') o.output += '/';
  // the reference evaluates -iteration while training is still true (VIO.java:141-150)
  if (has_iter) {
    o.iterations = atoi(iter_arg.c_str());
    check(o.iterations > 0, "Illegal settings for iteration option: must be strictly positive...");
  }
  if (has_index && !has_test) {
    o.resume = true;
    check(o.snapshot < o.iterations,
          "Option iteration and option modelindex do not agree with each other: option iteration "
          "must be strictly larger than option modelindex...");
  }
  if (has_test) {
    check(has_index, "Model index missing: modelindex was not initialized...");
    if (o.model_path.back() != '/') o.model_path += '/';
    o.training = false;
    o.resume = false;
  }
  if (has_sym) {
    check(o.training, "Option symmetricalpha ignored due to testing mode...");
    check(!o.resume, "Option symmetricalpha ignored due to resume...");
    o.symmetric = true;
  }
  check(o.topics > 0, "Illegal settings for topic option: must be strictly positive...");
  check(o.terms > 0, "Illegal settings for term option: must be strictly positive...");
  if (!o.training) o.random_start = false;  // "ignored in testing mode" (VIO.java:216-223)
  check(o.mappers > 0, "Illegal settings for mapper option: must be strictly positive...");
  if (!o.training) o.informed_prior.clear();  // "ignored in test mode" (VIO.java:226-233)
  check(o.world >= 1 && o.rank >= 0 && o.rank < o.world, "bad -rank/-world");
  return o;
}

// Recursive delete that never follows symbolic links: a link is unlinked, its target is left alone.
void rm_rf(const std::string &p) {
  struct stat st;
  if (lstat(p.c_str(), &st) != 0) return;
  if (S_ISDIR(st.st_mode)) {
    DIR *d = opendir(p.c_str());
    if (d) {
      while (dirent *e = readdir(d)) {
        std::string n = e->d_name;
        if (n == "." || n == "..") continue;
        rm_rf(p + (p.back() == '/' ? "" : "/") + n);
      }
      closedir(d);
    }
    rmdir(p.c_str());
  } else {
    unlink(p.c_str());
  }
}

// The gamma rotation (VariationalInference.java:359-379) deletes the previous iteration's gamma
// directory.  The reference calls fs.delete on whatever the previous input directory was, which in
// "-test <model> -modelindex N" with N > 0 is the user's -input path; here only a gamma-N directory
// under the output path is ever removed (deliberate deviation).
bool is_own_gamma_dir(const std::string &dir, const std::string &output) {
  const std::string prefix = output + "gamma-";
  if (dir.compare(0, prefix.size(), prefix) != 0 || dir.size() == prefix.size()) return false;
  for (size_t i = prefix.size(); i < dir.size(); i++)
    if (dir[i] < '0' || dir[i] > '9') return false;
  return true;
}

void mkdirs(const std::string &p) {
  std::string cur;
  for (size_t i = 0; i < p.size(); i++) {
    cur.push_back(p[i]);
    if (p[i] == '/' || i + 1 == p.size()) mkdir(cur.c_str(), 0777);
  }
}

void die(mrlda_ctx *ctx, const char *what) {
  fprintf(stderr, "ERROR %s: %s\n", what, mrlda_last_error(ctx));
  exit(1);
}

}  // namespace

int main(int argc, char **argv) {
  Options o = parse(argc, argv);
  const int K = o.topics, V = o.terms;
  info("Tool: VariationalInference");
  info(" - input path: " + o.input);
  info(" - output path: " + o.output);
  info(" - number of topics: " + std::to_string(K));
  info(" - number of terms: " + std::to_string(V));
  info(" - number of iterations: " + std::to_string(o.iterations));
  info(" - number of mappers: " + std::to_string(o.mappers));
  info(" - number of reducers: " + std::to_string(o.reducers));
  info(std::string(" - training mode: ") + (o.training ? "true" : "false"));
  info(std::string(" - random start gamma: ") + (o.random_start ? "true" : "false"));
  info(std::string(" - resume training: ") + (o.resume ? "true" : "false"));
  info(std::string(" - direct emit from mapper: ") + (o.direct_emit ? "true" : "false"));
  info(std::string(" - symmetric alpha: ") + (o.symmetric ? "true" : "false"));

  try {
    // VariationalInference.java:111-116
    if (o.rank == 0) {
      if (!o.resume && exists(o.output)) rm_rf(o.output);
      mkdirs(o.output);
    }
    std::string input_dir = o.input;
    std::string alpha_dir, beta_dir;
    const std::string alpha_path = o.output + "alpha-", beta_path = o.output + "beta-";
    if (!o.training) {
      alpha_dir = o.model_path + "alpha-" + std::to_string(o.snapshot);
      beta_dir = o.model_path + "beta-" + std::to_string(o.snapshot);
    } else if (!o.resume) {
      alpha_dir = alpha_path + "0";
      if (o.rank == 0) write_alpha(alpha_dir, std::vector<double>((size_t)K, 1e-3));  // :155-168
    } else {
      alpha_dir = alpha_path + std::to_string(o.snapshot);
      beta_dir = beta_path + std::to_string(o.snapshot);
      input_dir = o.output + "gamma-" + std::to_string(o.snapshot);
    }

    mrlda_config cfg;
    mrlda_default_config(&cfg);
    cfg.n_topics = K;
    cfg.n_terms = V;
    cfg.learning = o.training;
    cfg.random_start = o.random_start;
    cfg.symmetric_alpha = o.symmetric;
    cfg.device = o.device;
    cfg.rank = o.rank;
    cfg.world_size = o.world;
    cfg.seed = o.seed;
    cfg.deterministic = o.deterministic ? 1 : 0;
    mrlda_ctx *ctx = nullptr;
    if (mrlda_create(&cfg, &ctx) != MRLDA_OK) die(nullptr, "mrlda_create");

    if (!o.informed_prior.empty()) {  // VariationalInference.java:118-124,198-201
      check(exists(o.informed_prior) && !is_dir(o.informed_prior),
            "Illegal informed prior file: must be an existing file...");
      std::vector<int32_t> topic, term;
      read_eta(o.informed_prior, topic, term);
      if (o.rank == 0) {  // the reference copies the file to <output>eta
        FILE *src = fopen(o.informed_prior.c_str(), "rb"), *dst = fopen((o.output + "eta").c_str(), "wb");
        if (src && dst) {
          char buf[1 << 16];
          size_t got;
          while ((got = fread(buf, 1, sizeof buf, src)) > 0) fwrite(buf, 1, got, dst);
        }
        if (src) fclose(src);
        if (dst) fclose(dst);
      }
      if (mrlda_set_informed_prior(ctx, (int64_t)topic.size(), topic.data(), term.data()) != MRLDA_OK)
        die(ctx, "mrlda_set_informed_prior");
      info(" - informed prior: " + o.informed_prior + " (" + std::to_string(topic.size()) + " seed terms)");
    }

    if (o.world > 1) {  // one process per GPU: rank 0 publishes the NCCL id through a file
      check(!o.nccl_id_file.empty(), "-ncclid <file> is required with -world > 1");
      uint8_t id[MRLDA_NCCL_ID_BYTES];
      if (o.rank == 0) {
        if (mrlda_nccl_unique_id(id) != MRLDA_OK) die(nullptr, "mrlda_nccl_unique_id");
        FILE *f = fopen((o.nccl_id_file + ".tmp").c_str(), "wb");
        fwrite(id, 1, sizeof id, f);
        fclose(f);
        rename((o.nccl_id_file + ".tmp").c_str(), o.nccl_id_file.c_str());
      } else {
        while (!exists(o.nccl_id_file)) std::this_thread::sleep_for(std::chrono::milliseconds(50));
        FILE *f = fopen(o.nccl_id_file.c_str(), "rb");
        if (!f || fread(id, 1, sizeof id, f) != sizeof id) usage("cannot read " + o.nccl_id_file);
        fclose(f);
      }
      if (mrlda_comm_init(ctx, id) != MRLDA_OK) die(ctx, "mrlda_comm_init");
    }

    // the input split of this rank: contiguous documents balanced by nnz; a rank reads (and keeps)
    // only its own shard
    int64_t first_doc = 0, total_docs = 0;
    Corpus all = read_corpus_shard(input_dir, K, V, o.rank, o.world, &first_doc, &total_docs);
    const int64_t lo = 0, hi = all.n_docs();
    const bool stored_gamma = !all.gamma.empty();
    if (mrlda_set_corpus_csr(ctx, hi, all.row_ptr.data(), all.term_id.data(), all.count.data(),
                             all.doc_id.data(), stored_gamma ? all.gamma.data() : nullptr) != MRLDA_OK)
      die(ctx, "mrlda_set_corpus_csr");
    if (o.world > 1)
      info(" - input split of rank " + std::to_string(o.rank) + ": documents [" + std::to_string(first_doc) +
           ", " + std::to_string(first_doc + hi) + ") of " + std::to_string(total_docs));

    double last_ll = 0;
    int iteration = o.snapshot;
    bool model_loaded = false;
    std::vector<double> gamma((size_t)(hi - lo) * K), dl((size_t)V * K), alpha((size_t)K);
    std::vector<float> nrm((size_t)K);
    std::vector<uint8_t> present((size_t)V);

    do {
      if (!model_loaded) {
        if (iteration != 0) {  // VariationalInference.java:191-194
          check(exists(beta_dir), "Missing model parameter beta...");
          std::vector<double> bdl;
          std::vector<float> bn;
          std::vector<uint8_t> bp;
          read_beta(beta_dir, K, V, bdl, bn, bp);
          if (mrlda_set_beta(ctx, bdl.data(), bn.data(), bp.data()) != MRLDA_OK) die(ctx, "mrlda_set_beta");
        } else {
          // no beta file at iteration 0: every row is initialised lazily by retrieveBeta
          if (mrlda_init_beta_random(ctx) != MRLDA_OK) die(ctx, "mrlda_init_beta_random");
        }
        check(exists(alpha_dir), "Missing model parameter alpha...");  // :195
        std::vector<double> a0 = read_alpha(alpha_dir, K);
        if (mrlda_set_alpha(ctx, a0.data()) != MRLDA_OK) die(ctx, "mrlda_set_alpha");
        model_loaded = true;
      }
      auto t0 = std::chrono::steady_clock::now();
      mrlda_iter_result res;
      if (mrlda_iterate(ctx, &res) != MRLDA_OK) die(ctx, "mrlda_iterate");
      double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      info("Iteration " + std::to_string(iteration + 1) + " finished in " + java_double(secs) + " seconds");
      const double ll = res.log_likelihood;  // -counter / 1e6 (VariationalInference.java:261-262)
      info("Log likelihood of the model is: " + java_double(ll));
      info("Total number of documents is: " + std::to_string((int)res.n_docs));
      info("Total number of terms is: " + std::to_string((int)res.n_terms_seen));
      // CONFIG_TIME / TRAINING_TIME counters per document (VariationalInference.java:270-275): the
      // model stays resident on the device (no per-mapper configure); the E-step is the training time
      info("Average time elapsed for mapper configuration (ms): " + java_double(0.0));
      info("Average time elapsed for processing a document (ms): " +
           java_double(res.n_docs > 0 ? res.e_step_ms / (double)res.n_docs : 0.0));
      info("E-step " + java_double(res.e_step_ms) + " ms, all-reduce " + java_double(res.allreduce_ms) +
           " ms, M-step " + java_double(res.m_step_ms) + " ms, alpha " + java_double(res.alpha_ms) +
           " ms on the device");

      if (o.training) {
        if (mrlda_get_alpha(ctx, alpha.data()) != MRLDA_OK) die(ctx, "mrlda_get_alpha");
        // :284,291,312 — the old alpha and the alpha sufficient statistics never leave the device
        info("Successfully import old alpha vector from file " + alpha_dir);
        info("Successfully import alpha sufficient statistics tokens from file " + o.output + "temp/part-00000");
        info("Successfully update new alpha vector.");
        alpha_dir = alpha_path + std::to_string(iteration + 1);
        if (mrlda_get_beta(ctx, dl.data(), nrm.data(), present.data()) != MRLDA_OK) die(ctx, "mrlda_get_beta");
        beta_dir = beta_path + std::to_string(iteration + 1);
        if (o.rank == 0) {
          write_alpha(alpha_dir, alpha);  // :315-318
          info("Successfully export new alpha vector to file " + alpha_dir);
          write_beta(beta_dir, K, V, dl.data(), nrm.data(), present.data());  // :346-348
        }
      }
      if (!o.random_start || !o.training) {  // :359-379
        std::string gamma_dir = input_dir;
        input_dir = o.output + "gamma-" + std::to_string(iteration + 1);
        mkdirs(input_dir);
        if (mrlda_get_gamma(ctx, gamma.data()) != MRLDA_OK) die(ctx, "mrlda_get_gamma");
        // one part file per ~64 MB, written concurrently (the map phase's gamma_gamma-m-NNNNN files)
        write_gamma_parts(input_dir, o.rank, all, lo, hi, gamma.data(), K);
        if (iteration != 0 && o.rank == 0 && is_own_gamma_dir(gamma_dir, o.output)) rm_rf(gamma_dir);
      }
      info("Log likelihood after iteration " + std::to_string(iteration + 1) + " is " + java_double(ll));
      if (std::fabs((last_ll - ll) / last_ll) <= 0.000001) {  // Settings.java:56
        info("Model converged after " + std::to_string(iteration + 1) + " iterations...");
        break;
      }
      last_ll = ll;
      iteration++;
    } while (iteration < o.iterations);
    mrlda_destroy(ctx);
  } catch (const std::exception &e) {
    fprintf(stderr, "ERROR %s\n", e.what());
    return 1;
  }
  return 0;
}
