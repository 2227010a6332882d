"""Development probe: end-to-end ff_discover (host buffers) against the device-resident step, for several sub-batch cuts,
with full hit lists and with compact ones (+ ff_hits_resolve)."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flashfry_b200.api as ff
from flashfry_b200 import _native as N
from bench import make_guides
ctx = ff.Context(0); ctx.synth_database(3, 300_000_000, 3001)
G = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
g = make_guides(G, 3002)
pinned = torch.from_numpy(g.view(np.int64)).pin_memory(); gh = pinned.numpy().view(np.uint64)
gp = gh.ctypes.data_as(C.POINTER(C.c_uint64)); hp = C.POINTER(N.FFHits)()
d_g = torch.from_numpy(g.view(np.int64)).cuda()
for _ in range(3):
    ctx.discover_device(d_g.data_ptr(), G, 4, 2000)
print("device step %.3f ms" % ctx.timings().total_ms, flush=True)
t0 = time.perf_counter(); ctx_ptr = N.lib().ff_db_host_targets(ctx._h); print("host mirror %.2f s" % (time.perf_counter() - t0), flush=True)


def run(label, resolve=False, reps=20):
    for _ in range(3):
        N.check(N.lib().ff_discover(ctx._h, gp, G, 4, 2000, 0, C.byref(hp)))
        if resolve: N.check(N.lib().ff_hits_resolve(ctx._h, hp))
        N.lib().ff_hits_free(hp)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        N.check(N.lib().ff_discover(ctx._h, gp, G, 4, 2000, 0, C.byref(hp)))
        if resolve: N.check(N.lib().ff_hits_resolve(ctx._h, hp))
        N.lib().ff_hits_free(hp)
    print(label, "%.2f ms" % ((time.perf_counter() - t0) / reps * 1e3), flush=True)


for two in (55, 60, 65, 70, 75):
    ctx.set_option("subbatch_two", two); ctx.set_option("subbatch_min", 45000)
    ctx.set_option("compact_hits", 0); run("full    two sub-batches, first %d %%:" % two)
    ctx.set_option("compact_hits", 1); run("compact two sub-batches, first %d %%:" % two)
ctx.set_option("compact_hits", 0)
for cuts, mb in (("55,85", 30000), ("60,88", 30000), ("65,90", 30000)):
    c1, c2 = cuts.split(","); ctx.set_option("subbatch_c1", int(c1)); ctx.set_option("subbatch_c2", int(c2)); ctx.set_option("subbatch_min", mb)
    run("full    three sub-batches %s:" % cuts)
