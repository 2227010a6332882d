"""Exploration: time the general path (bulge / windowed) on the bench-sized synthetic index.
usage: python tools/bulge_probe.py G k flags [cells]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import flashfry_b200.api as ff
from bench import make_guides

G, k, flags = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
if len(sys.argv) > 4:
    os.environ["FF_WINDOW_CELLS"] = sys.argv[4]
n_t = int(os.environ.get("N_TARGETS", "300000000"))
ctx = ff.Context(0)
ctx.synth_database(3, n_t, 3001)
guides = make_guides(G, 3002)
d = torch.from_numpy(guides.view(np.int64)).cuda()
for it in range(3):
    t0 = time.perf_counter()
    r = ctx.discover_bulge_device(d.data_ptr(), G, k, 2000, flags) if flags >= 0 else ctx.discover_device(d.data_ptr(), G, k, 2000, 0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tm = ctx.timings()
    print("G=%d k=%d flags=%d: %.1f ms wall, scan %.2f ms, order %.2f, cut %.2f, total %.2f; hits %d cand %d compares/guide %.0f launches %d scan_launches %d  -> %.0f guides/s" % (
        G, k, flags, dt * 1e3, tm.scan_ms, tm.order_ms, tm.cut_ms, tm.total_ms, r.n_hits, r.n_candidate_hits, r.n_compares / G, tm.kernel_launches, tm.scan_launches, G / dt), flush=True)
