#!/usr/bin/env python
"""bench.py — E-step tokens/sec of the Mr.LDA variational-inference hot path on B200.

Contract: `python bench.py --gpus N --steps K --warmup W` (under torchrun for N > 1) prints ONE
JSON line.  A step is one pass of the E-step (kernel K1 incl. the fused statistics scatter, plus the
NCCL exchange when N > 1) over the Wikipedia-shape synthetic corpus (BASELINE.json configs[3]:
1M docs, ~200 tokens/doc, V=100k, K=200), sharded by nnz over the N ranks (strong scaling: the
corpus is fixed).  `--impl reference` times the CPU restatement of the reference's own algorithm
(oracle/, log-space updatePhi, 99 sweeps) on the host cores: the reference itself is Java/Hadoop and
this image has no JVM.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "E-step tokens/sec (device-timed)"
UNIT = "tokens/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self._stop = index, [], threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.startswith("Active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def measured_traffic(workload, n_gpus):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/r2_traffic.json: the sum over the E-step launches of one step); None when no capture exists for this workload / GPU count."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        t = json.load(f)
    if t.get("workload") != workload or t.get("n_gpus") != n_gpus:
        return None
    return t["dram_bytes_read"] + t["dram_bytes_write"]


def algorithmic_bytes(Z, D, K, s=8):
    """B_E = Z(8 + 2Ks) + 2DKs + 8(D+1)  (SURVEY.md section 8d)."""
    return Z * (8 + 2 * K * s) + 2 * D * K * s + 8 * (D + 1)


def algorithmic_flops(Z, D, K, sweeps=99):
    """F_E = I*Z*4K + I*D*K*60 (linear-space form)."""
    return sweeps * Z * 4 * K + sweeps * D * K * 60


def cpu_reference_rate(c, logbeta, alpha, target_s=15.0):
    """Times the oracle's E-step (all host threads) on a bounded document sample."""
    from oracle import oracle as O
    cores = O.set_num_threads()  # every core this process may run on; OMP_NUM_THREADS is ignored
    D = c.n_docs
    pilot = min(D, 4 * max(cores, 8))
    sub = c.slice_docs(0, pilot)
    O.estep(c.n_topics, sub.row_ptr, sub.term_id, sub.count, logbeta, alpha)  # thread-pool warm-up
    t0 = time.perf_counter()
    O.estep(c.n_topics, sub.row_ptr, sub.term_id, sub.count, logbeta, alpha)
    dt = max(time.perf_counter() - t0, 1e-3)
    n = int(min(D, max(pilot, pilot * target_s / dt)))
    n = min(max(cores, n // cores * cores), D)
    sub = c.slice_docs(0, n)
    t0 = time.perf_counter()
    O.estep(c.n_topics, sub.row_ptr, sub.term_id, sub.count, logbeta, alpha)
    dt = time.perf_counter() - t0
    return sub.n_tokens / dt, cores, n, sub.n_tokens, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="wikipedia", choices=["ap", "20news", "wikipedia", "largek"])
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the workload's documents")
    ap.add_argument("--fit-iterations", type=int, default=0,
                    help="after the bench: a full fit of this many EM iterations on the same context "
                         "(tools/r2_fit.py; oracle parity samples on rank 0; not part of the bench line)")
    ap.add_argument("--fit-sample-at", default="1,25,50")
    ap.add_argument("--fit-sample-docs", type=int, default=500)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--fp32", action="store_true", help="optional fp32 mode (1e-4 tolerance); not the default")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    import torch
    from mrlda_b200 import corpus

    D0, V, K, mean_len, _ = corpus.SHAPES[args.workload]
    wl = {"workload": f"{args.workload}-shape synthetic corpus", "D": int(round(D0 * args.scale)),
          "V": V, "K": K, "sweeps": 99, "init": "alpha=1e-3, logbeta=log(2U1/V+U2) seed 7, cold gamma"}

    seed = list(corpus.SHAPES).index(args.workload) + 1
    heavy = corpus.SHAPES[args.workload][4]

    def config_of(c, world_):
        return dict(wl, docs=c.n_docs, nnz=c.nnz, tokens=c.n_tokens,
                    parallelism=f"documents sharded by nnz over {world_} GPU(s)",
                    l2="inputs (CSR + E + gamma + stats > 1 GB per GPU) exceed the 126 MB L2; no flush")

    if args.impl == "reference":
        if rank != 0:
            return
        # The CPU arm: the oracle's restatement of the reference's own algorithm (log-space
        # updatePhi, 99 sweeps) on every host core this process may use, each step over a bounded
        # sample = the first n documents of the SAME corpus object the GPU arm runs on (same
        # generator, seed and size; generated on cuda:0 when a GPU is present because the 1M-document
        # sampler takes minutes on the CPU, otherwise the head of the document range is generated
        # on the CPU with the same seed and the line says so).
        from oracle import oracle as O
        cores = O.set_num_threads()
        t_gen = time.perf_counter()
        if torch.cuda.is_available():
            full = corpus.generate(wl["D"], V, K, mean_len, seed=seed, heavy_tail=heavy, device="cuda:0")
            torch.cuda.empty_cache()
            origin = "the first {n} documents of the same corpus object (full.slice_docs)"
        else:
            full = corpus.generate(min(wl["D"], 4096), V, K, mean_len, seed=seed, heavy_tail=heavy)
            origin = "the first {n} of a 4096-document corpus from the same generator and seed (no GPU to generate the full one)"
        t_gen = time.perf_counter() - t_gen
        lb = corpus.init_logbeta(V, K, seed=7)
        alpha = np.full(K, 1e-3)
        rate, cores, n, ntok, dt = cpu_reference_rate(full, lb, alpha, target_s=8.0)
        sub = full.slice_docs(0, n)
        for _ in range(1 if args.warmup > 0 else 0):
            O.estep(K, sub.row_ptr, sub.term_id, sub.count, lb, alpha)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.estep(K, sub.row_ptr, sub.term_id, sub.count, lb, alpha)
        dt = (time.perf_counter() - t0) / args.steps
        val = sub.n_tokens / dt
        sample = origin.format(n=n) + f", {sub.n_tokens} tokens per step"
        cfg_full = config_of(full, world) if full.n_docs == wl["D"] else dict(wl, parallelism="host cores")
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg_full, "corpus_gen_s": round(t_gen, 2),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": sample,
                             "note": "CPU restatement of Mr.LDA's path (Hadoop LocalJobRunner unavailable: no JVM)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)

    from mrlda_b200 import capi
    t_gen = time.perf_counter()
    full = corpus.generate(wl["D"], V, K, mean_len, seed=seed, heavy_tail=heavy,
                           device=f"cuda:{local_rank}")
    torch.cuda.empty_cache()
    t_gen = time.perf_counter() - t_gen
    mine, (lo, hi) = corpus.shard(full, rank, world)
    lb = corpus.init_logbeta(V, K, seed=7)
    alpha = np.full(K, 1e-3)

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy(), t

    row_ptr, _k1 = pinned(mine.row_ptr)
    term_id, _k2 = pinned(mine.term_id)
    count, _k3 = pinned(mine.count)
    lb_p, _k4 = pinned(lb)
    gamma_out_t = torch.empty((mine.n_docs, K), dtype=torch.float64).pin_memory()
    gamma_out = gamma_out_t.numpy()

    ctx = capi.Context(K, V, device=local_rank, rank=rank, world_size=world, fp32_mode=int(args.fp32))
    if world > 1:
        idt = torch.zeros(capi.MRLDA_NCCL_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            idt = torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(idt, 0)
        ctx.comm_init(bytes(idt.numpy().tobytes()))
    ctx.set_corpus_csr(row_ptr, term_id, count)
    ctx.set_alpha(alpha)
    ctx.set_logbeta(lb_p)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg: inputs already in HBM -------------------------------------------
    for _ in range(args.warmup):
        ctx.e_step()
    barrier()
    results = []
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            results.append(ctx.e_step())
        barrier()
        wall = time.perf_counter() - t0
    dev_ms = sum(r["e_step_ms"] + r["allreduce_ms"] for r in results)
    k1_ms = sum(r["e_step_ms"] for r in results) / args.steps
    ar_ms = sum(r["allreduce_ms"] for r in results) / args.steps
    launches = sum(r["estep_launches"] for r in results)

    # ---- end-to-end leg: host buffers in, host results out, every step -------------------------
    e2e = None
    if not args.no_e2e:
        def e2e_step():
            ctx.set_corpus_csr(row_ptr, term_id, count)
            ctx.set_alpha(alpha)
            ctx.set_logbeta(lb_p)
            r = ctx.e_step()
            ctx.get_gamma(out=gamma_out)
            ctx.get_alpha_ss()
            return r
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, args.steps)
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        e2e_wall = (time.perf_counter() - t0) / n_e2e
        h2d = row_ptr.nbytes + term_id.nbytes + count.nbytes + 2 * 4 * mine.n_docs + lb_p.nbytes + 8 * K
        d2h = gamma_out.nbytes + 8 * K + 32 + 16
        e2e = dict(wall=e2e_wall, h2d=h2d, d2h=d2h)

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    wall_max = allmax(wall)
    dev_ms_max = allmax(dev_ms)
    k1_ms_max = allmax(k1_ms)
    ar_ms_max = allmax(ar_ms)
    total_tokens = full.n_tokens
    launches_total = int(allsum(launches))
    e2e_wall_max = allmax(e2e["wall"]) if e2e else None
    h2d_tot = allsum(e2e["h2d"]) if e2e else 0
    d2h_tot = allsum(e2e["d2h"]) if e2e else 0

    # ---- after the timed regions: what the exchange produced, for the N = 1/2/4/8 lines to agree on.
    # stats_sum: the K x V statistics after exchange() must add up to the corpus's token count
    # (sum_k phi = 1); then one M-step (finalize + alpha) from those statistics and a checksum of the
    # new logbeta / alpha.  Every rank holds the same replica; rank 0 reports.
    check = None
    if True:
        stats = ctx.get_stats()
        ssum = float(stats.sum(dtype=np.float64))
        del stats
        mres = ctx.m_step()
        dl, nm, pr = ctx.get_beta()
        seen = pr.astype(bool)
        new_lb = ctx.get_logbeta()[seen]
        wts = 1.0 + (np.arange(new_lb.shape[1], dtype=np.float64) % 7.0)
        check = {"stats_sum": ssum, "stats_sum_rel_err": abs(ssum - total_tokens) / total_tokens,
                 "terms_seen": int(seen.sum()),
                 "logbeta_sum": float(new_lb.sum(dtype=np.float64)),
                 "logbeta_weighted_sum": float((new_lb * wts[None, :]).sum(dtype=np.float64)),
                 "normaliser_sum": float(nm.astype(np.float64).sum()),
                 "alpha_sum": float(ctx.get_alpha().sum()),
                 "m_step_ms": mres["m_step_ms"], "alpha_ms": mres["alpha_ms"]}
        del dl, new_lb

    fp64_peak = ctx.measure_fp64_peak() if rank == 0 else 0.0
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, cores, n, ntok, dt = cpu_reference_rate(full, lb, alpha, target_s=15.0)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {n} documents ({ntok} tokens) of the same corpus, {dt:.1f} s, "
                         f"oracle log-space E-step, OpenMP over documents",
               "note": "CPU restatement of Mr.LDA's path (Hadoop LocalJobRunner unavailable: no JVM)"}
    if rank == 0:
        hbm, how = _peaks()
        Zmax = max(int(full.row_ptr[b1] - full.row_ptr[b0]) for b0, b1 in
                   zip(corpus.shard_bounds(full.row_ptr, world)[:-1],
                       corpus.shard_bounds(full.row_ptr, world)[1:]))
        Dmax = max(b1 - b0 for b0, b1 in zip(corpus.shard_bounds(full.row_ptr, world)[:-1],
                                             corpus.shard_bounds(full.row_ptr, world)[1:]))
        B = algorithmic_bytes(Zmax, Dmax, K)
        F = algorithmic_flops(Zmax, Dmax, K)
        ach = B / (k1_ms_max * 1e-3) / 1e9
        value = total_tokens * args.steps / (dev_ms_max * 1e-3)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall_max / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 tile / f64 elsewhere (optional mode)" if args.fp32 else "f64",
            "data": "synthetic",
            "config": config_of(full, world), "corpus_gen_s": round(t_gen, 2),
            "device_ms_per_step": dev_ms_max / args.steps,
            "k1_ms": k1_ms_max, "allreduce_ms": ar_ms_max,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s",
                         "frac": ach / hbm,
                         "traffic": measured_traffic(args.workload, world) if args.scale == 1.0 else None,
                         "peak_source": how,
                         "kernel": "estep2f_kernel (fp32 tile) + estep_kernel (fp64) for longer documents" if args.fp32
                         else capi.describe_kernels(K),  # the library's own plan for this topic count
                         "algorithmic_bytes_per_launch": B,
                         "note": ("optional fp32 mode: estep2f_kernel keeps an fp32 tile on chip, but E rows, gamma "
                                  "and the statistics stay fp64 in HBM, so the algorithmic bytes are the s = 8 "
                                  "figure; the fp64 block below is the DFMA fraction of the fp64 formula and does "
                                  "not describe this mode's FFMA work") if args.fp32
                         else "binding resource is the FP64 pipe, see fp64"},
            "fp64": {"achieved_tflops": F / (k1_ms_max * 1e-3) / 1e12, "peak_tflops": fp64_peak,
                     "frac": (F / (k1_ms_max * 1e-3) / 1e12) / fp64_peak if fp64_peak else None,
                     "peak_source": "DFMA probe kernel, this run", "algorithmic_flops_per_launch": F},
            "clocks": clk.summary(),
            "gpu_launches": launches_total,
            "slow_path_rows": results[-1]["slow_path_rows"],
            "log_likelihood": results[-1]["log_likelihood"],
        }
        if check:
            out["exchange_check"] = check
        if e2e:
            out["e2e"] = {"value": total_tokens / e2e_wall_max, "unit": UNIT,
                          "h2d_bytes_per_step": int(h2d_tot), "d2h_bytes_per_step": int(d2h_tot),
                          "ms_per_step": e2e_wall_max * 1e3}
        if cpu:
            out["cpu_baseline"] = cpu
        print(json.dumps(out), flush=True)
    # ---- optional: a full fit on the same context (saves a second process start on a multi-GPU box) ----
    if args.fit_iterations > 0:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
        import r2_fit
        ctx.set_corpus_csr(row_ptr, term_id, count)  # cold gamma again
        ctx.set_alpha(alpha)
        ctx.set_logbeta(lb_p)
        sample_at = sorted({int(x) for x in args.fit_sample_at.split(",") if x and int(x) <= args.fit_iterations})
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):  # stdout carries the one JSON line only
            rows, parity, fit_wall = r2_fit.run_iterations(ctx, mine, K, rank, args.fit_iterations, sample_at,
                                                           min(args.fit_sample_docs, mine.n_docs))
            if rank == 0:
                r2_fit.write_summary(args.workload, "", world, full, K, V, False, rows, parity, fit_wall, t_gen)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
