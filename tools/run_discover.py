#!/usr/bin/env python3
"""Minimal driver for profiling: synthetic index + a few ff_discover calls (same kernels bench.py times)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import flashfry_b200.api as ff  # noqa: E402
from bench import make_guides, SEED_DB, SEED_GUIDES, SEED_PLANTED, ENZYME  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--targets", type=int, default=300_000_000)
ap.add_argument("--guides", type=int, default=100_000)
ap.add_argument("--k", type=int, default=4)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--score", action="store_true")
a = ap.parse_args()
ctx = ff.Context(0)
ctx.synth_database(ENZYME, a.targets, SEED_DB)
n_t = int(ctx.info().n_targets)
rng = np.random.default_rng(17)
pool = np.concatenate([ctx.copy_targets(int(s), 2048) for s in rng.integers(0, max(1, n_t - 2048), 16)])
guides = make_guides(a.guides, SEED_GUIDES, pool, SEED_PLANTED)
for i in range(a.steps):
    t0 = time.perf_counter()
    if a.score:
        h = ctx.discover_score(guides, a.k, 2000)[0]
    else:
        h = ctx.discover(guides, a.k, 2000)
    dt = time.perf_counter() - t0
    tm = ctx.timings()
    print("step %d: %.1f ms wall, scan %.2f ms, prep %.2f, order %.2f, cut %.2f, score %.2f, hits %d, compares %d, sub_bases %d"
          % (i, dt * 1e3, tm.scan_ms, tm.prep_ms, tm.order_ms, tm.cut_ms, tm.score_ms, len(h.targets), h.n_compares,
             ctx.info().seed_split_a), flush=True)
