import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flashfry_b200.api as ff
from flashfry_b200 import _native as N
from bench import make_guides
ctx = ff.Context(0); ctx.synth_database(3, 300_000_000, 3001)
g = make_guides(100000, 3002)
pinned = torch.from_numpy(g.view(np.int64)).pin_memory(); gh = pinned.numpy().view(np.uint64)
gp = gh.ctypes.data_as(C.POINTER(C.c_uint64)); hp = C.POINTER(N.FFHits)()
for cuts, mb in (("50,80", "20000"), ("60,85", "20000"), ("65,90", "20000"), ("55,85", "20000"), ("70,90","20000"), ("50,80", "45000")):
    c1, c2 = cuts.split(","); ctx.set_option("subbatch_c1", int(c1)); ctx.set_option("subbatch_c2", int(c2)); ctx.set_option("subbatch_min", int(mb))
    for _ in range(3):
        N.check(N.lib().ff_discover(ctx._h, gp, len(g), 4, 2000, 0, C.byref(hp))); N.lib().ff_hits_free(hp)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        N.check(N.lib().ff_discover(ctx._h, gp, len(g), 4, 2000, 0, C.byref(hp))); N.lib().ff_hits_free(hp)
    dt = (time.perf_counter() - t0) / 20
    print(cuts, mb, "%.2f ms" % (dt * 1e3), flush=True)
