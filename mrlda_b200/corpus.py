"""Synthetic corpora in the shapes BASELINE.json names, CSR packing and document sharding.

The reference's input is ParseCorpus's SequenceFile<IntWritable, cc.mrlda.Document>: per document a
bag {termID -> count} with 1-based term ids ranked by document frequency, most frequent first
(ParseCorpus.java:465,484-487; README.md:284).  generate() draws documents from the LDA generative
model and re-ranks ids the same way; pack/shard helpers mirror what the map phase's input splits do
(documents are independent, VariationalInference.java:218,237).
"""
from dataclasses import dataclass

import numpy as np
import torch

# name -> (D, V, K, mean tokens/doc, heavy_tail)   BASELINE.json configs[0..4]
SHAPES = {
    "ap": (2246, 10473, 20, 194, False),
    "20news": (11000, 8000, 100, 120, False),
    "wikipedia": (1_000_000, 100_000, 200, 200, False),
    "largek": (2_000_000, 200_000, 1000, 200, True),
}


@dataclass
class Corpus:
    row_ptr: np.ndarray  # int64 [D+1]
    term_id: np.ndarray  # int32 [Z], 1-based
    count: np.ndarray    # int32 [Z]
    n_terms: int
    n_topics: int

    @property
    def n_docs(self):
        return len(self.row_ptr) - 1

    @property
    def nnz(self):
        return int(self.row_ptr[-1])

    @property
    def n_tokens(self):
        return int(self.count.sum(dtype=np.int64))

    def slice_docs(self, lo, hi):
        b, e = int(self.row_ptr[lo]), int(self.row_ptr[hi])
        return Corpus((self.row_ptr[lo:hi + 1] - b).astype(np.int64), self.term_id[b:e],
                      self.count[b:e], self.n_terms, self.n_topics)


def generate(D, V, K, mean_len, seed=0, heavy_tail=False, device="cpu", chunk_tokens=40_000_000,
             topic_conc=0.1, doc_conc=0.1):
    """LDA generative sampler.  Returns a Corpus with DF-ranked 1-based term ids."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed) * 7919 + 17)
    f64 = torch.float64

    def gamma_rand(shape, conc):
        # torch's _standard_gamma honours the generator on both the CPU and the CUDA backend
        a = torch.full(shape, conc, dtype=f64, device=dev)
        return torch._standard_gamma(a, generator=g).clamp_min_(1e-300)

    # topics: sparse Gamma draws on a Zipf-Mandelbrot base measure (exponent 1.2, offset 3: gives
    # ~130 distinct terms per 200-token document at V = 100k)
    zipf = 1.0 / (torch.arange(V, dtype=f64, device=dev) + 3.0) ** 1.2
    beta = gamma_rand((K, V), topic_conc) * zipf
    beta /= beta.sum(1, keepdim=True)
    cdf_b = torch.cumsum(beta, 1)
    cdf_b[:, -1] = 1.0
    flat_b = (cdf_b + torch.arange(K, dtype=f64, device=dev)[:, None]).reshape(-1)
    del beta, cdf_b

    # document lengths: Poisson-lognormal
    sigma = 0.9 if heavy_tail else 0.5
    ln = torch.randn(D, dtype=f64, device=dev, generator=g) * sigma + (np.log(mean_len) - 0.5 * sigma**2)
    lengths = torch.poisson(torch.exp(ln), generator=g).clamp_min_(1).to(torch.int64)
    if heavy_tail:  # >= 1 % of documents above 2000 tokens (cfg 5)
        n_long = max(1, D // 100)
        idx = torch.randperm(D, device=dev, generator=g)[:n_long]
        lengths[idx] = torch.randint(2001, 6000, (n_long,), device=dev, generator=g)

    keys = []
    start = 0
    csum = torch.cumsum(lengths, 0)
    while start < D:
        base = int(csum[start - 1]) if start else 0
        end = int(torch.searchsorted(csum, torch.tensor(base + chunk_tokens, device=dev)).item()) + 1
        end = min(max(end, start + 1), D)
        nd = end - start
        lens = lengths[start:end]
        theta = gamma_rand((nd, K), doc_conc)
        theta /= theta.sum(1, keepdim=True)
        cdf_t = torch.cumsum(theta, 1)
        cdf_t[:, -1] = 1.0
        flat_t = (cdf_t + torch.arange(nd, dtype=f64, device=dev)[:, None]).reshape(-1)
        del theta, cdf_t
        doc = torch.repeat_interleave(torch.arange(nd, device=dev), lens)
        nt = doc.numel()
        u = torch.rand(nt, dtype=f64, device=dev, generator=g).clamp_(0.0, 1.0 - 1e-12)
        z = torch.searchsorted(flat_t, doc.to(f64) + u) - doc * K
        z.clamp_(0, K - 1)
        u = torch.rand(nt, dtype=f64, device=dev, generator=g).clamp_(0.0, 1.0 - 1e-12)
        w = torch.searchsorted(flat_b, z.to(f64) + u) - z * V
        w.clamp_(0, V - 1)
        keys.append((doc + start) * V + w)
        del flat_t, doc, u, z, w
        start = end
    key = torch.cat(keys)
    del keys
    key, _ = torch.sort(key)
    uniq, cnt = torch.unique_consecutive(key, return_counts=True)
    del key
    doc = torch.div(uniq, V, rounding_mode="floor")
    term = uniq - doc * V
    # document-frequency ranking, most frequent term gets id 1 (ParseCorpus.java:465,484-487)
    df = torch.bincount(term, minlength=V)
    order = torch.argsort(df, descending=True, stable=True)
    rank = torch.empty(V, dtype=torch.int64, device=dev)
    rank[order] = torch.arange(1, V + 1, device=dev)
    term_id = rank[term]
    nnz_per_doc = torch.bincount(doc, minlength=D)
    row_ptr = torch.zeros(D + 1, dtype=torch.int64, device=dev)
    row_ptr[1:] = torch.cumsum(nnz_per_doc, 0)
    return Corpus(row_ptr.cpu().numpy(), term_id.to(torch.int32).cpu().numpy(),
                  cnt.to(torch.int32).cpu().numpy(), int(V), int(K))


def generate_shape(name, seed=None, device="cpu", scale=1.0):
    """One of BASELINE.json's shapes; scale < 1 shrinks the document count only."""
    D, V, K, mean_len, heavy = SHAPES[name]
    D = max(1, int(round(D * scale)))
    if seed is None:
        seed = list(SHAPES).index(name) + 1
    return generate(D, V, K, mean_len, seed=seed, heavy_tail=heavy, device=device)


def init_logbeta(V, K, seed=0):
    """Injected initial log-beta in the reference's own initialisation form log(2*U1/V + U2)
    (DocumentMapper.java:456), from a fixed seed so that oracle and GPU start identically."""
    rng = np.random.default_rng(seed)
    return np.log(2.0 * rng.random((V, K)) / V + rng.random((V, K)))


def from_docs(docs, n_terms, n_topics):
    """docs: list of {termID: count} (the HMapII content of cc.mrlda.Document; None/{} = null)."""
    row_ptr = [0]
    tid, cnt = [], []
    for d in docs:
        if d:
            for t, c in d.items():
                tid.append(t)
                cnt.append(c)
        row_ptr.append(len(tid))
    return Corpus(np.asarray(row_ptr, dtype=np.int64), np.asarray(tid, dtype=np.int32),
                  np.asarray(cnt, dtype=np.int32), n_terms, n_topics)


def shard_bounds(row_ptr, world_size):
    """Contiguous document ranges balanced by nnz (not by document count): rank r owns documents
    [bounds[r], bounds[r+1]).  Replaces the input splits of the map phase (-mapper)."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    D = len(row_ptr) - 1
    Z = int(row_ptr[-1])
    bounds = [0]
    for r in range(1, world_size):
        target = Z * r // world_size
        b = int(np.searchsorted(row_ptr, target, side="left"))
        bounds.append(min(max(b, bounds[-1]), D))
    bounds.append(D)
    return bounds


def shard(corpus, rank, world_size):
    b = shard_bounds(corpus.row_ptr, world_size)
    return corpus.slice_docs(b[rank], b[rank + 1]), (b[rank], b[rank + 1])
