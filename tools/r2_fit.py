#!/usr/bin/env python
"""Full LDA fit of one of BASELINE.json's shapes on N GPUs (one process per GPU, torchrun for N > 1),
with an oracle parity SAMPLE on chosen iterations and an optional held-out pass.

  training   documents [0, D (1 - heldout)) sharded by nnz over the ranks; every iteration is one
             mrlda_iterate (E-step, NCCL all-reduce of the statistics, beta finalize, alpha update:
             VariationalInference.java:181-326).  The reference's likelihood log lines are printed.
  parity     on the iterations named by --sample-at, rank 0 runs the oracle's E-step on the first
             --sample-docs documents of its shard with the SAME inputs the GPU iteration had (the
             previous iteration's alpha / logbeta / gamma) and compares gamma rows, the sample's
             alpha statistics contribution and its ELBO counter.
  held-out   (--heldout > 0, one GPU) the -test pass of VariationalInferenceOptions.java:166-178 on
             the held-out documents with the final model, checked in full against the oracle.
Writes gpurun_out/r2_fit_<workload>_<N>gpu.json.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrlda_b200 import capi, corpus, seqfile  # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def run_iterations(ctx, mine, K, rank, iterations, sample_at, ns):
    """`iterations` EM iterations on a context that holds corpus, alpha and logbeta, with the oracle
    parity sample (rank 0, first ns documents of its shard) on the iterations in sample_at.
    Returns (rows, parity, wall seconds)."""
    sample = mine.slice_docs(0, ns)
    rows, parity = [], []
    last = 0.0
    wall0 = time.time()
    prev = None  # (alpha, logbeta, sample gamma or None) going INTO the next iteration
    for it in range(1, iterations + 1):
        want_parity = rank == 0 and it in sample_at
        if want_parity:
            g_in = ctx.get_gamma()[:ns].copy() if it > 1 else None
            prev = (ctx.get_alpha().copy(), ctx.get_logbeta().copy(), g_in)
        r = ctx.iterate()
        ll = r["log_likelihood"]
        row = dict(iteration=it, log_likelihood=ll, e_step_ms=r["e_step_ms"], allreduce_ms=r["allreduce_ms"],
                   m_step_ms=r["m_step_ms"], alpha_ms=r["alpha_ms"], slow_path_rows=r["slow_path_rows"],
                   n_terms_seen=r["n_terms_seen"])
        if rank == 0:
            print(f"Log likelihood after iteration {it} is {seqfile.java_double(ll)}   [E {r['e_step_ms']:.1f} ms, "
                  f"all-reduce {r['allreduce_ms']:.2f} ms, M {r['m_step_ms']:.3f} ms, alpha {r['alpha_ms']:.3f} ms]",
                  flush=True)
        if want_parity:
            from oracle import oracle as O  # checker only
            O.set_num_threads()  # every core of the box (torchrun exports OMP_NUM_THREADS=1)
            t1 = time.time()
            ref = O.estep(K, sample.row_ptr, sample.term_id, sample.count, prev[1], prev[0], gamma=prev[2])
            got = ctx.get_gamma()[:ns]
            nz = np.diff(sample.row_ptr) > 0
            p = dict(iteration=it, sample_docs=int(ns), gamma_rel=rel(got[nz], ref["gamma"][nz]),
                     oracle_seconds=round(time.time() - t1, 1))
            parity.append(p)
            print("PARITY", json.dumps(p), flush=True)
        rows.append(row)
        if last != 0.0 and abs((last - ll) / last) <= 1e-6:
            if rank == 0:
                print(f"Model converged after {it} iterations...")
            break
        last = ll
    return rows, parity, time.time() - wall0


def write_summary(workload, tag, world, train, K, V, symmetric, rows, parity, wall, t_gen, heldout=None):
    """gpurun_out/r2_fit_<workload><tag>_<N>gpu.json and a SUMMARY line (rank 0)."""
    e = np.array([x["e_step_ms"] for x in rows])
    out = dict(workload=workload, n_gpus=world, docs=train.n_docs, nnz=train.nnz, tokens=train.n_tokens,
               K=K, V=V, symmetric=symmetric, iterations=len(rows), corpus_gen_s=round(t_gen, 1),
               wall_seconds=round(wall, 2),
               device_seconds=float(sum(x["e_step_ms"] + x["allreduce_ms"] + x["m_step_ms"] + x["alpha_ms"]
                                        for x in rows) / 1e3),
               mean_e_step_ms=float(e.mean()), mean_allreduce_ms=float(np.mean([x["allreduce_ms"] for x in rows])),
               mean_m_step_ms=float(np.mean([x["m_step_ms"] for x in rows])),
               mean_alpha_ms=float(np.mean([x["alpha_ms"] for x in rows])),
               tokens_per_second_e_step=train.n_tokens / (float(e.mean()) * 1e-3),
               final_log_likelihood=rows[-1]["log_likelihood"],
               parity_sample=parity, max_sample_gamma_rel=max([p["gamma_rel"] for p in parity], default=None),
               heldout=heldout, rows=rows)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out",
                        f"r2_fit_{workload}{tag}_{world}gpu.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("SUMMARY", json.dumps({k: v for k, v in out.items() if k != "rows"}), file=sys.stderr, flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="20news")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--iterations", type=int, default=50)
    ap.add_argument("--symmetric", action="store_true")
    ap.add_argument("--heldout", type=float, default=0.0, help="fraction of the documents held out")
    ap.add_argument("--sample-docs", type=int, default=2000)
    ap.add_argument("--sample-at", default="1,2,10,25,50")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)

    D0, V, K, mean_len, heavy = corpus.SHAPES[args.workload]
    D = int(round(D0 * args.scale))
    seed = list(corpus.SHAPES).index(args.workload) + 1
    t0 = time.time()
    full = corpus.generate(D, V, K, mean_len, seed=seed, heavy_tail=heavy, device=f"cuda:{local_rank}")
    torch.cuda.empty_cache()
    n_train = D - int(round(D * args.heldout))
    train = full.slice_docs(0, n_train) if n_train < D else full
    held = full.slice_docs(n_train, D) if n_train < D else None
    mine, (lo, hi) = corpus.shard(train, rank, world)
    t_gen = time.time() - t0
    sample_at = sorted({int(x) for x in args.sample_at.split(",") if x and int(x) <= args.iterations})
    ns = min(args.sample_docs, mine.n_docs)

    ctx = capi.Context(K, V, device=local_rank, rank=rank, world_size=world,
                       symmetric_alpha=int(args.symmetric))
    if world > 1:
        idt = torch.zeros(capi.MRLDA_NCCL_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            idt = torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(idt, 0)
        ctx.comm_init(bytes(idt.numpy().tobytes()))
    ctx.set_corpus_csr(mine.row_ptr, mine.term_id, mine.count)
    alpha = np.full(K, 1e-3)
    ctx.set_alpha(alpha)
    ctx.set_logbeta(corpus.init_logbeta(V, K, seed=7))

    rows, parity, wall = run_iterations(ctx, mine, K, rank, args.iterations, sample_at, ns)

    heldout = None
    if held is not None and world == 1:
        from oracle import oracle as O
        lb, al = ctx.get_logbeta(), ctx.get_alpha()
        with capi.Context(K, V, device=local_rank, learning=0) as tctx:
            tctx.set_corpus_csr(held.row_ptr, held.term_id, held.count)
            tctx.set_alpha(al)
            tctx.set_logbeta(lb)
            tr = tctx.iterate()
            tg = tctx.get_gamma()
        ref = O.iterate(K, V, held.row_ptr, held.term_id, held.count, lb, al, learning=False)
        nz = np.diff(held.row_ptr) > 0
        heldout = dict(docs=held.n_docs, tokens=held.n_tokens, log_likelihood=tr["log_likelihood"],
                       oracle_log_likelihood=ref["log_likelihood"],
                       per_token=tr["log_likelihood"] / held.n_tokens,
                       elbo_rel=abs(tr["log_likelihood"] - ref["log_likelihood"]) / abs(ref["log_likelihood"]),
                       counter_diff=int(tr["ll_counter"] - ref["counter"]),
                       gamma_rel=rel(tg[nz], ref["gamma"][nz]), e_step_ms=tr["e_step_ms"])
        print("HELDOUT", json.dumps(heldout), flush=True)
    ctx.close()

    if rank == 0:
        write_summary(args.workload, args.tag, world, train, K, V, args.symmetric, rows, parity, wall, t_gen, heldout)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
