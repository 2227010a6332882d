"""Development probe: run the same discover call through every scan kernel, check they agree, print timings.

    python tools/scan_probe.py [n_targets] [n_guides] [k] [modes e.g. 0,1,2] [reps]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
import flashfry_b200.api as ff  # noqa: E402

n_t = int(float(sys.argv[1])) if len(sys.argv) > 1 else 300_000_000
G = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 4
modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["1", "2"]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5

ctx = ff.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
t0 = time.time()
ctx.synth_database(3, n_t, bench.SEED_DB)
print("db: %d targets, %.1f s, %.2f GB" % (ctx.info().n_targets, time.time() - t0, ctx.info().device_bytes / 1e9), flush=True)
n = int(ctx.info().n_targets)
sample = ctx.copy_targets(0, min(n, 4_000_000))
guides = bench.make_guides(G, bench.SEED_GUIDES, sample, bench.SEED_PLANTED)
d_g = torch.from_numpy(guides.view(np.int64)).cuda()

ref = None
for m in modes:  # "2" or "2:pair_kernel=1:pair_segs=8" (scan kernel, then options)
    parts = m.split(":")
    ctx.set_option("scan_kernel", int(parts[0]))
    for kv in parts[1:]:
        key, val = kv.split("=")
        ctx.set_option(key, int(val))
    best = None
    for r in range(reps):
        res = ctx.discover_device(d_g.data_ptr(), len(guides), k, 2000)
        tm = ctx.timings()
        if best is None or tm.total_ms < best[0]:
            best = (tm.total_ms, tm.prep_ms, tm.scan_ms, tm.order_ms, tm.cut_ms, tm.scan_part1_ms, tm.scan_part2_ms)
    H = int(res.n_hits)
    rp = torch.empty(len(guides) + 1, dtype=torch.int64, device="cuda")
    tg = torch.empty(max(H, 1), dtype=torch.int64, device="cuda")
    mm = torch.empty(max(H, 1), dtype=torch.uint8, device="cuda")
    import ctypes as C
    cudart = C.CDLL("libcudart.so")
    cudart.cudaMemcpy(C.c_void_p(rp.data_ptr()), C.c_void_p(res.d_row_ptr), C.c_size_t(rp.numel() * 8), 3)
    if H:
        cudart.cudaMemcpy(C.c_void_p(tg.data_ptr()), C.c_void_p(res.d_targets), C.c_size_t(H * 8), 3)
        cudart.cudaMemcpy(C.c_void_p(mm.data_ptr()), C.c_void_p(res.d_mismatches), C.c_size_t(H), 3)
    torch.cuda.synchronize()
    cur = (rp.cpu().numpy(), tg.cpu().numpy()[:H], mm.cpu().numpy()[:H])
    print("mode %s: total %.3f ms (prep %.3f scan %.3f order %.3f cut %.3f; part one %.3f part two %.3f) hits %d cand %d compares %.3e" %
          ((m,) + best + (H, int(res.n_candidate_hits), float(res.n_compares))), flush=True)
    if ref is None:
        ref = cur
    else:
        same = all(np.array_equal(a, b) for a, b in zip(ref, cur))
        print("   agrees with mode %s: %s" % (modes[0], same), flush=True)
        if not same:
            for g in range(len(guides)):
                if ref[0][g + 1] - ref[0][g] != cur[0][g + 1] - cur[0][g]:
                    print("   first differing guide", g, "rows", ref[0][g + 1] - ref[0][g], cur[0][g + 1] - cur[0][g])
                    break
