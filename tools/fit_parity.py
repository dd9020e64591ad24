#!/usr/bin/env python
"""Full EM fit on the GPU next to the CPU oracle, both starting from the same injected beta.

Two things are reported per iteration:
  step parity  — the GPU iteration and the oracle iteration are run on the SAME inputs (the GPU's
                 current alpha / beta / gamma), i.e. the per-iteration tolerance of the north star;
  trajectory   — the oracle is also run free (feeding its own outputs forward); the drift between
                 the two trajectories is printed for information (rounding differences compound
                 through the EM map, so this is not expected to stay at 1e-9).
Writes a JSON summary (default profiles/r1_fit_parity_<workload>.json).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrlda_b200 import capi, corpus  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker only)


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="ap")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--iterations", type=int, default=50)
    ap.add_argument("--symmetric", action="store_true")
    ap.add_argument("--free-oracle", action="store_true", help="also run the oracle trajectory")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    c = corpus.generate_shape(args.workload, scale=args.scale, device="cuda")
    K, V = c.n_topics, c.n_terms
    alpha = np.full(K, 1e-3)
    lb = corpus.init_logbeta(V, K, seed=7)
    rows = []
    with capi.Context(K, V, symmetric_alpha=int(args.symmetric)) as ctx:
        ctx.set_corpus_csr(c.row_ptr, c.term_id, c.count)
        ctx.set_alpha(alpha)
        ctx.set_logbeta(lb)
        g_alpha, g_lb, g_gamma = alpha, lb, None
        f_alpha, f_lb, f_gamma = alpha, lb, None
        gpu_ms = 0.0
        for it in range(1, args.iterations + 1):
            r = ctx.iterate()
            gpu_ms += r["e_step_ms"] + r["m_step_ms"] + r["alpha_ms"]
            t0 = time.time()
            ref = O.iterate(K, V, c.row_ptr, c.term_id, c.count, g_lb, g_alpha, gamma=g_gamma,
                            symmetric=args.symmetric)
            cpu_s = time.time() - t0
            seen = ref["present"].astype(bool)
            new_lb, new_alpha, new_gamma = ctx.get_logbeta(), ctx.get_alpha(), ctx.get_gamma()
            row = dict(iteration=it, log_likelihood=r["log_likelihood"],
                       oracle_log_likelihood=ref["log_likelihood"],
                       elbo_rel=abs(r["log_likelihood"] - ref["log_likelihood"]) / abs(ref["log_likelihood"]),
                       logbeta_rel=rel(new_lb[seen], ref["logbeta"][seen]),
                       alpha_rel=rel(new_alpha, ref["alpha"]), gamma_rel=rel(new_gamma, ref["gamma"]),
                       normaliser_equal=bool(np.array_equal(ctx.get_beta()[1], ref["normaliser"])),
                       slow_path_rows=r["slow_path_rows"], oracle_seconds=round(cpu_s, 2))
            if args.free_oracle:
                fr = O.iterate(K, V, c.row_ptr, c.term_id, c.count, f_lb, f_alpha, gamma=f_gamma,
                               symmetric=args.symmetric)
                f_lb = np.where(fr["present"].astype(bool)[:, None], fr["logbeta"], f_lb)
                f_alpha, f_gamma = fr["alpha"], fr["gamma"]
                row["trajectory_elbo_rel"] = abs(r["log_likelihood"] - fr["log_likelihood"]) / abs(fr["log_likelihood"])
            rows.append(row)
            print(json.dumps(row), flush=True)
            g_lb, g_alpha, g_gamma = new_lb, new_alpha, new_gamma
    summary = dict(workload=args.workload, docs=c.n_docs, nnz=c.nnz, tokens=c.n_tokens, K=K, V=V,
                   iterations=args.iterations, symmetric=args.symmetric,
                   max_step_elbo_rel=max(x["elbo_rel"] for x in rows),
                   max_step_logbeta_rel=max(x["logbeta_rel"] for x in rows),
                   max_step_alpha_rel=max(x["alpha_rel"] for x in rows),
                   max_step_gamma_rel=max(x["gamma_rel"] for x in rows),
                   normaliser_mismatches=sum(not x["normaliser_equal"] for x in rows),
                   gpu_device_seconds=gpu_ms / 1e3, rows=rows)
    out = args.out or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                   "gpurun_out", f"r1_fit_parity_{args.workload}.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        json.dump(summary, f, indent=1)
    print("SUMMARY", json.dumps({k: v for k, v in summary.items() if k != "rows"}))


if __name__ == "__main__":
    main()
