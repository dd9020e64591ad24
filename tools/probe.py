#!/usr/bin/env python
"""E-step timing probe for kernel experiments (not part of the bench contract).

Generates a reduced Wikipedia-shape corpus, optionally keeps only documents whose distinct-term
count lies in [--min-rows, --max-rows], and times mrlda_e_step.  Kernel variant / CTAs per SM are
selected by the MRLDA_ESTEP_VARIANT / MRLDA_ESTEP_CTAS_PER_SM environment variables.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrlda_b200 import capi, corpus  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="wikipedia")
    ap.add_argument("--docs", type=int, default=20000)
    ap.add_argument("--topics", type=int, default=0, help="override the workload's topic count")
    ap.add_argument("--terms", type=int, default=0, help="override the workload's vocabulary size")
    ap.add_argument("--min-rows", type=int, default=0)
    ap.add_argument("--max-rows", type=int, default=1 << 30)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--sweep-cap", type=int, default=100)
    ap.add_argument("--fp32", action="store_true", help="optional fp32 mode")
    ap.add_argument("--no-learning", action="store_true", help="test mode: no statistics scatter")
    ap.add_argument("--deterministic", action="store_true", help="deterministic statistics mode")
    args = ap.parse_args()
    D0, V, K, mean_len, heavy = corpus.SHAPES[args.workload]
    K = args.topics or K
    V = args.terms or V
    c = corpus.generate(args.docs, V, K, mean_len, seed=4, heavy_tail=heavy, device="cuda")
    n = np.diff(c.row_ptr)
    keep = np.nonzero((n >= args.min_rows) & (n <= args.max_rows))[0]
    if len(keep) != c.n_docs:
        rp = np.zeros(len(keep) + 1, dtype=np.int64)
        rp[1:] = np.cumsum(n[keep])
        idx = np.concatenate([np.arange(c.row_ptr[d], c.row_ptr[d + 1]) for d in keep])
        c = corpus.Corpus(rp, c.term_id[idx], c.count[idx], V, K)
    n = np.diff(c.row_ptr)
    with capi.Context(K, V, sweep_cap=args.sweep_cap, fp32_mode=int(args.fp32),
                      learning=0 if args.no_learning else 1, deterministic=int(args.deterministic)) as ctx:
        ctx.set_corpus_csr(c.row_ptr, c.term_id, c.count)
        ctx.set_alpha(np.full(K, 1e-3))
        ctx.set_logbeta(corpus.init_logbeta(V, K, seed=7))
        ctx.e_step()
        ms = [ctx.e_step()["e_step_ms"] for _ in range(args.steps)]
        peak = ctx.measure_fp64_peak()
        ph = ctx.debug_phase_cycles()
    if ph[7]:
        names = (['thetaLDS', 'passA', 'norms', 'passB', 'pub+B1', 'gamma', 'B2'] if os.environ.get('MRLDA_ESTEP_KERNEL', '3') == '3'
                 else ['theta', 'dots', 'B1', 'norms', 'B2', 'S', 'RS'])
        print('phase cycles per sweep (block 0, thread 0): ' + ', '.join(f'{n}={ph[i] / ph[7]:.0f}' for i, n in enumerate(names)) + f'  total={ph[:7].sum() / ph[7]:.0f}  slots/sweep={ph[8] / ph[7]:.2f}  expdigamma={ph[9] / ph[7]:.0f}')
    if ph[13]:
        print(f'per document (block 0): load={ph[10] / ph[13]:.0f} sweeps={ph[11] / ph[13]:.0f} '
              f'last+epilogue={ph[12] / ph[13]:.0f} scatter={ph[14] / ph[13]:.0f} cycles over {ph[13]} documents')
    t = min(ms) * 1e-3
    flops = (args.sweep_cap - 1) * (c.nnz * 4 * K + c.n_docs * K * 60)
    print(f"docs={c.n_docs} nnz={c.nnz} rows mean={n.mean():.1f} p90={np.percentile(n, 90):.0f} "
          f"max={n.max()} tokens={c.n_tokens}  k1={min(ms):.2f} ms  {c.n_tokens / t:.3e} tok/s  "
          f"{flops / t / 1e12:.2f} TF/s = {flops / t / 1e12 / peak:.1%} of {peak:.1f} (variant "
          f"{os.environ.get('MRLDA_ESTEP_VARIANT', 'auto')}, ctas/sm "
          f"{os.environ.get('MRLDA_ESTEP_CTAS_PER_SM', 'auto')})")


if __name__ == "__main__":
    main()
