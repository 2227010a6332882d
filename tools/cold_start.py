#!/usr/bin/env python3
"""Cold-start timing on the chr22 quick-start database: ff_load_database (BGZF inflate + block decode + H2D + index build)
against ff_load_image (the SoA side-car: one read + H2D + index build)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import flashfry_b200.api as ff
db = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "_chr22", "chr22_cas9ngg_database")
ctx = ff.Context(0)
for i in range(3):
    t0 = time.perf_counter(); ctx.load_database(db); dt = time.perf_counter() - t0
    info = ctx.info()
    print("ff_load_database %d: %.3f s  (%d targets, %d positions, %.1f MB on disk, %.2f GB in HBM)" %
          (i, dt, info.n_targets, info.n_positions, os.path.getsize(db) / 1e6, info.device_bytes / 1e9))
img = os.path.join(tempfile.gettempdir(), "chr22.ffimg")
t0 = time.perf_counter(); ctx.save_image(img); print("ff_save_image: %.3f s, %.1f MB" % (time.perf_counter() - t0, os.path.getsize(img) / 1e6))
for i in range(3):
    t0 = time.perf_counter(); ctx.load_image(img); dt = time.perf_counter() - t0
    print("ff_load_image %d: %.3f s" % (i, dt))
