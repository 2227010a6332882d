"""Cold start at human-genome size, in FlashFry's OWN on-disk format (SURVEY.md 8(d) cfg 3, 8(f1)):

  1. generate the synthetic 3e8-target index in HBM, copy the targets out;
  2. write them as a real FlashFry database (BGZF body + .header, all bins indexed) with oracle/big_db.py;
  3. ff_load_database it into a fresh context: `cold_load_s` (what replaces the reference's ~40 s of seek + inflate,
     reference/traverser/SeekTraverser.scala:113-121), check copy_targets() == the generator's array and the positions;
  4. ff_save_image / ff_load_image: `image_load_s` (the side-car skips inflate and block walk).

    python tools/cold_start.py [n_targets] [directory]      -> one JSON object on stdout
"""
import json
import os
import resource
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def measure(n_targets=300_000_000, where=None, threads=None, keep=False):
    import flashfry_b200.api as ff
    from oracle import big_db, ff_oracle as o
    import bench
    threads = threads or min(32, os.cpu_count() or 1)
    where = where or tempfile.mkdtemp(prefix="ff_cold_")
    path = os.path.join(where, "synth_hg38_cas9ngg_database")
    out = {"n_targets_requested": n_targets, "host_threads": threads}
    with ff.Context(0) as gen:
        gen.synth_database(bench.ENZYME, n_targets, bench.SEED_DB)
        targets = gen.copy_targets()
    t0 = time.perf_counter()
    st = big_db.write_big_database(path, o.PACK_BY_INDEX[bench.ENZYME], targets, threads=threads)
    out["write_s (oracle writer, not part of the product)"] = time.perf_counter() - t0
    out.update({"targets": st["targets"], "positions": st["positions"], "file_bytes": st["file_bytes"],
                "inflated_bytes": st["inflated_bytes"], "indexed_bins": st["indexed_bins"], "bgzf_members": st["members"]})
    rss0 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    with ff.Context(0) as ctx:
        t0 = time.perf_counter()
        ctx.load_database(path)
        out["cold_load_s"] = time.perf_counter() - t0
        out["peak_host_rss_gb"] = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6
        out["peak_host_rss_before_load_gb"] = rss0 / 1e6
        info = ctx.info()
        got = ctx.copy_targets()
        out["targets_equal_generator"] = bool(len(got) == len(targets) and (got == targets).all())
        out["n_positions_loaded"] = int(info.n_positions)
        # positions: a few guides with position output, checked against the writer's formula
        g = (targets[[5, len(targets) // 2, len(targets) - 7]] & np.uint64(0xFFFFFFFFFFFF)) | (np.uint64(1) << np.uint64(48))
        h = ctx.discover(g, 0, 10 ** 6, positions=True)
        pos_off = None
        ok = True
        counts = (targets >> np.uint64(48)).astype(np.int64)
        for row, ti in enumerate([5, len(targets) // 2, len(targets) - 7]):
            lo, hi = int(h.row_ptr[row]), int(h.row_ptr[row + 1])
            ok = ok and hi - lo == 1 and int(h.targets[lo]) == int(targets[ti])
            first = int(counts[:ti].sum())
            want = big_db.synthetic_positions(first, int(counts[ti]))
            gotp = h.positions[int(h.pos_ptr[lo]):int(h.pos_ptr[lo + 1])]
            ok = ok and len(gotp) == len(want) and (np.sort(gotp) == np.sort(want)).all()
        out["positions_match_writer_formula"] = bool(ok)
        img = os.path.join(where, "synth.ffimage")
        t0 = time.perf_counter()
        ctx.save_image(img)
        out["image_save_s"] = time.perf_counter() - t0
        out["image_bytes"] = os.path.getsize(img)
    with ff.Context(0) as ctx:
        t0 = time.perf_counter()
        ctx.load_image(img)
        out["image_load_s"] = time.perf_counter() - t0
        out["image_targets_equal"] = bool((ctx.copy_targets(0, 1000) == targets[:1000]).all() and int(ctx.info().n_targets) == len(targets))
    if not keep:
        for f in (path, path + ".header", img):
            try:
                os.remove(f)
            except OSError:
                pass
    out["reference_note"] = ("FlashFry pays ~40 s of BGZF seek + inflate + byte[]->long[] per run on hg38 "
                             "(reference/traverser/SeekTraverser.scala:113-121; BASELINE.md: 44 s for one guide)")
    return out


if __name__ == "__main__":
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 300_000_000
    print(json.dumps(measure(n, sys.argv[2] if len(sys.argv) > 2 else None)))
