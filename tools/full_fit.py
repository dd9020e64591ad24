#!/usr/bin/env python
"""Full LDA fit (E-step + beta finalize + alpha update per iteration) on one GPU for one of
BASELINE.json's shapes; prints the reference's likelihood log lines and the device-time breakdown.
Writes gpurun_out/r1_full_fit_<workload>.json."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrlda_b200 import capi, corpus, seqfile  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="ap")
    ap.add_argument("--iterations", type=int, default=50)
    ap.add_argument("--symmetric", action="store_true")
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    c = corpus.generate_shape(args.workload, scale=args.scale, device="cuda")
    K, V = c.n_topics, c.n_terms
    rows = []
    t0 = time.time()
    with capi.Context(K, V, symmetric_alpha=int(args.symmetric)) as ctx:
        ctx.set_corpus_csr(c.row_ptr, c.term_id, c.count)
        ctx.set_alpha(np.full(K, 1e-3))
        ctx.set_logbeta(corpus.init_logbeta(V, K, seed=7))
        last = 0.0
        for it in range(1, args.iterations + 1):
            r = ctx.iterate()
            ll = r["log_likelihood"]
            print(f"Log likelihood after iteration {it} is {seqfile.java_double(ll)}   "
                  f"[E {r['e_step_ms']:.2f} ms, M {r['m_step_ms']:.3f} ms, alpha {r['alpha_ms']:.3f} ms]",
                  flush=True)
            rows.append(dict(iteration=it, log_likelihood=ll, e_step_ms=r["e_step_ms"],
                             m_step_ms=r["m_step_ms"], alpha_ms=r["alpha_ms"],
                             slow_path_rows=r["slow_path_rows"], n_terms_seen=r["n_terms_seen"]))
            if last != 0.0 and abs((last - ll) / last) <= 1e-6:
                print(f"Model converged after {it} iterations...")
                break
            last = ll
    wall = time.time() - t0
    out = dict(workload=args.workload, docs=c.n_docs, nnz=c.nnz, tokens=c.n_tokens, K=K, V=V,
               symmetric=args.symmetric, iterations=len(rows), wall_seconds=wall,
               device_seconds=sum(x["e_step_ms"] + x["m_step_ms"] + x["alpha_ms"] for x in rows) / 1e3,
               mean_e_step_ms=float(np.mean([x["e_step_ms"] for x in rows])),
               mean_m_step_ms=float(np.mean([x["m_step_ms"] for x in rows])),
               mean_alpha_ms=float(np.mean([x["alpha_ms"] for x in rows])),
               tokens_per_second_e_step=c.n_tokens / (np.mean([x["e_step_ms"] for x in rows]) * 1e-3),
               rows=rows)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out",
                        f"r1_full_fit_{args.workload}{'_symmetric' if args.symmetric else ''}.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(out, open(path, "w"), indent=1)
    print("SUMMARY", json.dumps({k: v for k, v in out.items() if k != "rows"}))


if __name__ == "__main__":
    main()
