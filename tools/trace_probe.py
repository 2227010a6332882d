import os, sys, time, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import flashfry_b200.api as ff
from flashfry_b200 import _native as N
from bench import make_guides
ctx = ff.Context(0); ctx.synth_database(3, 300_000_000, 3001)
G = 100000
g = make_guides(G, 3002)
pinned = torch.from_numpy(g.view(np.int64)).pin_memory(); gh = pinned.numpy().view(np.uint64)
gp = gh.ctypes.data_as(C.POINTER(C.c_uint64)); hp = C.POINTER(N.FFHits)()
ctx.set_option("subbatch_c1", int(sys.argv[1]) if len(sys.argv) > 1 else 60); ctx.set_option("subbatch_c2", int(sys.argv[2]) if len(sys.argv) > 2 else 99); ctx.set_option("subbatch_min", int(sys.argv[3]) if len(sys.argv) > 3 else 45000)
for i in range(4):
    if i == 3: ctx.set_option("trace", 1)
    N.check(N.lib().ff_discover(ctx._h, gp, G, 4, 2000, 0, C.byref(hp))); N.lib().ff_hits_free(hp)
