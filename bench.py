#!/usr/bin/env python3
"""bench.py -- guides/sec of FlashFry's off-target discovery hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N --steps K --warmup W]             # this repo's CUDA path
    python bench.py --impl reference [...]                      # the reference's CPU algorithm (oracle port)

Workload (BASELINE.json configs[2], named in config.workload): 100 000 synthetic 20-bp NGG guides (+10 % planted
near real targets) against a synthetic human-genome-sized (3e8 distinct targets) spCas9-NGG index, <= 4 mismatches,
maximumOffTargets 2000.  The index is generated in HBM from a seed (it would be ~5 GB on disk); one replica per GPU.
A "step" = one discover call over one batch of guides per GPU; with N GPUs every rank processes its own batch
(guide-sharded, no data-path collective) and the ranks all-gather the per-guide hit counts at the end of the step.

`value`  : whole-job guides/s with guides already in HBM and results left in HBM (CUDA events, max over ranks).
`e2e`    : the same metric through the C ABI with HOST buffers (ff_discover: H2D of the guides, D2H of the hit lists).
`roofline`: the dominant kernel (k_cell_scan, the cell-major seed scan: one launch per index half, timed together with CUDA
            events on the launching stream) against the measured HBM copy bandwidth.  Algorithmic bytes per call = for
            every (guide, seed) the two 4-byte index entries + 4 bytes per index entry streamed from the seed's bucket
            + 8 bytes per guide + 8 bytes per candidate hit written; `traffic` = the HBM bytes ncu
            measured for the same call (L2 serves the repeated bucket reads, so it is below the algorithmic bytes).
`bulge_mode`, `fused_discover_score`: BASELINE.json configs[3] (an extension, no reference semantics) and configs[4] on
            the same batch.
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
sys.path.insert(0, ROOT)

SEED_DB, SEED_GUIDES, SEED_PLANTED = 3001, 3002, 3003
ENZYME = 3  # spCas9-NGG


def make_guides(n, seed, planted_from=None, planted_seed=0, planted_frac=0.10):
    """20 uniform bases + uniform N + GG, de-duplicated on the protospacer, count=1."""
    rng = np.random.default_rng(seed)
    n_pl = int(n * planted_frac) if planted_from is not None and len(planted_from) else 0
    n_rand = n - n_pl
    vals = np.zeros(0, np.uint64)
    while len(vals) < n_rand:
        v = rng.integers(0, 1 << 42, size=int((n_rand - len(vals)) * 1.05) + 16, dtype=np.uint64)
        vals = np.concatenate([vals, v])
        _, idx = np.unique(vals >> np.uint64(2), return_index=True)
        vals = vals[np.sort(idx)]
    g = (vals[:n_rand] << np.uint64(4)) | np.uint64(0xA)
    if n_pl:
        prng = np.random.default_rng(planted_seed)
        t = planted_from[prng.integers(0, len(planted_from), n_pl)] & np.uint64(0xFFFFFFFFFFFF)
        nsub = prng.integers(0, 5, n_pl)
        for j in range(4):
            pos = prng.integers(0, 20, n_pl)
            x = prng.integers(1, 4, n_pl).astype(np.uint64) << (np.uint64(2) * (np.uint64(22) - pos.astype(np.uint64)))
            t = np.where(nsub > j, t ^ x, t)
        g = np.concatenate([g, t])
        g = g[rng.permutation(len(g))]
    return (g | (np.uint64(1) << np.uint64(48))).astype(np.uint64)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark(self):
        """perf_counter timestamp, to delimit the timed region in stop()."""
        return time.perf_counter()

    def stop(self, t_from=None, t_to=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if t_from is not None and not (t_from <= ts <= t_to + 0.06):
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per k_scan launch from the committed ncu --set full summary, if one exists for this workload."""
    p = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def cell_major_expected(G, n_t):
    """Mirror of the library's choice (ff_discover.cu discover_plain): cell-major when part-one buckets are re-read."""
    return G * 529 / float(4 ** 11) >= 1.0 and n_t * 8.0 > 96e6


def run_native(args):
    import torch
    import torch.distributed as dist
    import flashfry_b200.api as ff

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    ctx = ff.Context(local)
    t0 = time.perf_counter()
    ctx.synth_database(ENZYME, args.targets, SEED_DB)
    torch.cuda.synchronize()
    db_s = time.perf_counter() - t0
    info = ctx.info()
    n_t = int(info.n_targets)
    # planted guides are drawn from a few slices of the resident database
    rng = np.random.default_rng(17)
    slices = [ctx.copy_targets(int(s), 2048) for s in rng.integers(0, max(1, n_t - 2048), 16)]
    pool = np.concatenate(slices)
    guides = make_guides(args.guides, SEED_GUIDES + 1000 * rank, pool, SEED_PLANTED + rank)
    G = len(guides)

    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    d_guides = torch.from_numpy(guides.view(np.int64)).to(dev)
    counts_all = torch.empty(world * G, dtype=torch.int32, device=dev) if world > 1 else None

    class _DevView:  # wrap a context-owned device pointer for torch without copying
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}

    def step():
        r = ctx.discover_device(d_guides.data_ptr(), G, args.k, args.max_ot, 0)
        if world > 1:  # the one collective of the path: every rank learns every guide's total count
            mine = torch.as_tensor(_DevView(r.d_total_count, G), device=dev)
            dist.all_gather_into_tensor(counts_all, mine)
        return r

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # nvidia-smi needs ~0.2 s to deliver its first sample: start it before the warm-up
    for _ in range(args.warmup):
        r = step()
    sync_all()
    t_region0 = sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_ms, launches, cand, compares, scan_bytes = [], 0, 0, 0, 0
    sync_all()
    e0.record(stream)
    for _ in range(args.steps):
        r = step()
        tm = ctx.timings()
        scan_ms.append(tm.scan_ms); launches += tm.kernel_launches + (1 if world > 1 else 0)
        cand, compares, scan_bytes = int(r.n_candidate_hits), int(r.n_compares), int(tm.scan_bytes_read)
        scan_launches = tm.scan_launches
    e1.record(stream)
    sync_all()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tmax = torch.tensor([ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    ms_per_step = ms / args.steps
    value = world * G / (ms_per_step / 1e3)
    n_hits = int(r.n_hits)

    # ---- e2e through the C ABI with host buffers (H2D guides, D2H hit lists), same steps
    pinned = torch.from_numpy(guides.view(np.int64)).pin_memory()
    g_host = pinned.numpy().view(np.uint64)
    import ctypes as C
    from flashfry_b200 import _native as N
    hp = C.POINTER(N.FFHits)()
    gp = g_host.ctypes.data_as(C.POINTER(C.c_uint64))

    def e2e_step():
        N.check(N.lib().ff_discover(ctx._h, gp, G, args.k, args.max_ot, 0, C.byref(hp)))
        nh = int(hp.contents.n_hits)
        N.lib().ff_hits_free(hp)
        return nh
    for _ in range(max(1, args.warmup // 2 + 1)):
        nh = e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        nh = e2e_step()
    sync_all()
    e2e_s = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop(t_region0, sampler.mark()) if rank == 0 else None  # samples taken inside the two timed regions
    if world > 1:
        tmax = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())
    e2e = {"value": world * G / e2e_s, "unit": "guides/s", "h2d_bytes_per_step": 8 * G,
           "d2h_bytes_per_step": (G + 1) * 8 + nh * 9 + G * 5, "ms_per_step": e2e_s * 1e3}

    out = None
    if rank == 0:
        peak, peak_src = measured_peak()
        scan_avg_ms = float(np.mean(scan_ms))
        achieved = scan_bytes / (scan_avg_ms / 1e3) / 1e9
        scan_kernel = "k_cell_scan" if cell_major_expected(G, n_t) else "k_seed_scan"
        roof = {"bound": "hbm", "kernel": scan_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "alg_bytes_per_launch": scan_bytes, "kernel_ms": scan_avg_ms,
                "kernel_share_of_step": scan_avg_ms / ms_per_step, "traffic": None,
                "entries_streamed_per_launch": compares, "entries_per_guide": compares / max(G, 1),
                # the second ceiling -- one 32-bit POPC per compared entry, 16 lanes/clk/SM x 148 SMs x 1.9 GHz
                "compares_per_s": compares / (scan_avg_ms / 1e3), "compare_ceiling_per_s": 4.5e12,
                "compare_frac": compares / (scan_avg_ms / 1e3) / 4.5e12}
        tr = ncu_traffic()
        if (tr and tr.get("targets") == n_t and tr.get("max_mismatch") == args.k and tr.get("kernel") == scan_kernel
                and tr.get("guides_in_profiled_launch") == G):
            # dram__bytes_read + dram__bytes_write of one launch of the same kernel on the same workload (ncu --set full)
            roof["traffic"] = tr.get("dram_bytes_per_profiled_launch")
            roof["traffic_source"] = tr.get("source")
            roof["note"] = ("algorithmic bytes = what the kernel requests per (guide, seed): index entries + streamed bucket entries; the "
                            "cell-major order serves repeated bucket reads from L2, so HBM traffic is BELOW the algorithmic bytes")
        out = {"metric": "guides/sec at <=4 mismatches vs hg38-sized index", "value": value, "unit": "guides/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": value / 53.7, "dtype": "u64", "data": "synthetic",
               "config": {"workload": "configs[2]: %d synthetic NGG guides per GPU (+10%% planted) vs synthetic %d-target spCas9-NGG index, "
                                      "k<=%d, maxOT %d" % (G, n_t, args.k, args.max_ot),
                          "guides_per_gpu": G, "targets": n_t, "max_mismatch": args.k, "maximum_off_targets": args.max_ot,
                          "parallelism": "guide-sharded x%d, one index replica per GPU" % world,
                          "l2": "index (%.1f GB) is larger than L2, re-streamed every step" % (info.device_bytes / 1e9),
                          "seed_split": "first %d | last %d protospacer bases" % (int(info.seed_split_a), 20 - int(info.seed_split_a)),
                          "db_build_s": db_s},
               "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof,
               "hits_per_step": n_hits, "candidate_hits_per_step": cand, "scan_launches_last_step": scan_launches,
               "baseline_note": "vs_baseline = value / 53.7 guides/s (published single-core JVM, 100k guides, hg38, k<=4; BASELINE.md)"}
        # the two scan kernels (guide-major k_seed_scan / cell-major k_cell_scan) must agree on the whole timed batch
        def _totals(kernel):
            ctx.set_option("scan_kernel", kernel)
            rr = ctx.discover_device(d_guides.data_ptr(), G, args.k, args.max_ot, 0)
            t = torch.as_tensor(_DevView(rr.d_total_count, G), device=dev).clone()
            ctx.set_option("scan_kernel", 0)
            return int(rr.n_hits), t
        h0, t0_ = _totals(1)
        h1, t1_ = _totals(2)
        out["scan_kernels_agree_on_full_batch"] = bool(h0 == h1 == n_hits and torch.equal(t0_, t1_))
        # BASELINE.json configs[4] flavour on the same batch: discover + CFD + Hsu2013 fused on the device
        fs = []
        for _ in range(3):
            ctx.discover_device(d_guides.data_ptr(), G, args.k, args.max_ot, 3)
            t = ctx.timings()
            fs.append((t.total_ms, t.score_ms))
        out["fused_discover_score"] = {"guides": G, "total_ms": float(np.median([a for a, _ in fs])),
                                       "score_kernel_ms": float(np.median([b for _, b in fs])),
                                       "guides_per_s": G / (float(np.median([a for a, _ in fs])) / 1e3),
                                       "metrics": "DoenchCFD_maxOT, DoenchCFD_specificityscore, Hsu2013 (FP64, bit-identical to the oracle)"}
        # BASELINE.json configs[3]: <= 5 mismatches + one 1-bp RNA/DNA bulge on the same batch.  An EXTENSION: the reference
        # has no bulge search, so this line is never part of a parity claim (semantics: include/flashfry_b200.h).
        if not args.no_bulge:
            bs = []
            for _ in range(3):
                rb = ctx.discover_bulge_device(d_guides.data_ptr(), G, 5, args.max_ot, 3)
                t = ctx.timings()
                bs.append((t.total_ms, t.scan_ms, t.scan_launches))
            bms = float(np.median([a for a, _, _ in bs]))
            for _ in range(2):  # the first call allocates the pinned result buffers (pooled afterwards)
                t0 = time.perf_counter()
                N.check(N.lib().ff_discover_bulge(ctx._h, gp, G, 5, args.max_ot, 3, 0, C.byref(hp)))
                bh = int(hp.contents.n_hits)
                N.lib().ff_hits_free(hp)
                be2e = time.perf_counter() - t0
            out["bulge_mode"] = {"workload": "configs[3]: %d guides, <=5 mismatches + one 1-bp RNA or DNA bulge, maxOT %d, same index" % (G, args.max_ot),
                                 "total_ms": bms, "scan_ms": float(np.median([b for _, b, _ in bs])), "windows": int(bs[-1][2]),
                                 "guides_per_s": G / (bms / 1e3), "hits": int(rb.n_hits), "candidate_hits": int(rb.n_candidate_hits),
                                 "overflowed_guides_note": "nearly every guide reaches maximumOffTargets: the scan walks database-order windows and drops full guides",
                                 "e2e_guides_per_s": G / be2e, "e2e_d2h_bytes": (G + 1) * 8 + bh * 10 + G * 5,
                                 "parity": "extension, no reference semantics: checked against a brute-force definition in tests/test_gpu_bulge.py"}
        # BASELINE.json configs[1]: 1 000 synthetic + 100 planted guides vs the chr22 quick-start database, when the database
        # built by __graft_entry__.build() travelled with the tree (33.5 MB of targets: L2-resident after the first pass)
        chr22 = os.path.join(ROOT, "tests", "golden", "_chr22", "chr22_cas9ngg_database")
        if os.path.exists(chr22) and os.path.exists(chr22 + ".header"):
            try:
                c2 = ff.Context(local)
                t0 = time.perf_counter()
                c2.load_database(chr22)
                load_s = time.perf_counter() - t0
                t22 = c2.copy_targets()
                g22 = make_guides(1100, 1001, t22[:: max(1, len(t22) // 4096)], 1002, planted_frac=100.0 / 1100.0)
                d22 = torch.from_numpy(g22.view(np.int64)).to(dev)
                ts = []
                for _ in range(8):
                    r22 = c2.discover_device(d22.data_ptr(), len(g22), args.k, args.max_ot, 0)
                    ts.append(c2.timings().total_ms)
                out["chr22_1000_guides"] = {"workload": "configs[1]: 1 000 synthetic + 100 planted guides vs the chr22 quick-start index (%d targets), k<=%d" % (len(t22), args.k),
                                            "total_ms": float(np.median(ts[2:])), "guides_per_s": len(g22) / (float(np.median(ts[2:])) / 1e3),
                                            "hits": int(r22.n_hits), "cold_load_s": load_s,
                                            "parity": "tests/test_gpu_parity.py::test_chr22_1000_guides_config (bit-exact vs the oracle)"}
                c2.close()
            except Exception as e:  # noqa: BLE001 -- a side measurement must not take the bench line down
                out["chr22_1000_guides"] = {"error": str(e)}
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(ctx, guides, args, threads=1, budget_guides=args.cpu_guides)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        _emit(out)


def cpu_baseline(ctx, guides, args, threads, budget_guides):
    """The oracle's restatement of the reference loop order (bin -> 256 sub-bin filters -> target x guide), timed on
    this box's host cores on a bounded sample of the same workload."""
    from oracle import ff_oracle as o
    pack = o.PACK_BY_INDEX[ENZYME]
    t = ctx.copy_targets()
    bin_off = o.bin_offsets_from_sorted(pack, 7, t)
    sample = guides[:budget_guides]
    t0 = time.perf_counter()
    ref = o.discover_soa(pack, 7, t, bin_off, sample, args.k, args.max_ot, n_threads=threads)
    dt = time.perf_counter() - t0
    # the sample doubles as a parity check of the timed GPU path
    ok = True
    for kernel in (1, 2):  # both scan kernels against the oracle, at the full index size
        ctx.set_option("scan_kernel", kernel)
        got = ctx.discover(sample, args.k, args.max_ot)
        ok = ok and bool((got.row_ptr == ref.row_ptr).all() and (got.targets == ref.targets).all() and (got.mismatches == ref.mismatches).all()
                         and (got.overflowed == ref.overflowed).all())
    ctx.set_option("scan_kernel", 0)
    _h, cmax, cspec, hsu = ctx.discover_score(sample[:64], args.k, args.max_ot)
    score_ok = True
    for g in range(min(64, len(sample))):
        ots = ref.targets[ref.row_ptr[g]:ref.row_ptr[g + 1]]
        mx, sp, _ = o.cfd_guide(int(sample[g]), ots)
        score_ok = score_ok and cmax[g] == mx and cspec[g] == sp and hsu[g] == o.hsu_guide(pack, int(sample[g]), ots)
    return {"value": len(sample) / dt, "unit": "guides/s", "cores": threads, "kind": "port",
            "sample": "%d of the %d guides vs the full %d-target index, oracle/ff_oracle.c ffo_discover_soa (C restatement of the "
                      "reference loop order, not the JVM), %.1f s" % (len(sample), len(guides), len(t), dt),
            "parity_with_gpu_on_sample": ok, "score_parity_on_64_guides": bool(score_ok), "reference_compares": int(ref.n_compares)}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the JVM cannot run here) on all host threads.
    Needs the GPU only to materialise the same synthetic index the native arm uses."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import flashfry_b200.api as ff
    from oracle import ff_oracle as o
    threads = os.cpu_count() or 1
    try:
        ctx = ff.Context(int(os.environ.get("LOCAL_RANK", "0")))
        ctx.synth_database(ENZYME, args.targets, SEED_DB)
    except Exception as e:  # noqa: BLE001
        _emit({"impl": "reference", "unavailable": "cannot materialise the synthetic index: %s" % e})
        return
    t = ctx.copy_targets()
    n_t = len(t)
    rng = np.random.default_rng(17)
    pool = np.concatenate([t[int(s):int(s) + 2048] for s in rng.integers(0, max(1, n_t - 2048), 16)])
    guides = make_guides(args.guides, SEED_GUIDES, pool, SEED_PLANTED)
    ctx.close()
    pack = o.PACK_BY_INDEX[ENZYME]
    bin_off = o.bin_offsets_from_sorted(pack, 7, t)
    n = min(len(guides), args.ref_guides)
    times = []
    for i in range(args.warmup + args.steps):
        sample = guides[(i * n) % max(1, len(guides) - n):][:n]
        t0 = time.perf_counter()
        o.discover_soa(pack, 7, t, bin_off, sample, args.k, args.max_ot, n_threads=threads)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    ms = float(np.mean(times)) * 1e3
    v = n / (ms / 1e3)
    sample_txt = ("%d-guide batches of the same %d-guide workload vs the full %d-target index per step, oracle port of the "
                  "reference loop order on %d host threads" % (n, len(guides), n_t, threads))
    _emit({"impl": "reference", "metric": "guides/sec at <=4 mismatches vs hg38-sized index", "value": v, "unit": "guides/s",
                      "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": v / 53.7, "dtype": "u64", "data": "synthetic",
                      "config": {"workload": "configs[2] bounded sample: " + sample_txt, "targets": n_t, "max_mismatch": args.k,
                                 "maximum_off_targets": args.max_ot},
                      "cpu_baseline": {"value": v, "unit": "guides/s", "cores": threads, "kind": "port", "sample": sample_txt},
                      "e2e": {"value": v, "unit": "guides/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def _emit(obj):
    """The ONE JSON line goes to the real stdout; everything else a library prints (NCCL's version banner ...) was sent
    to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--targets", type=int, default=300_000_000)
    ap.add_argument("--guides", type=int, default=100_000)
    ap.add_argument("--k", type=int, default=4)
    ap.add_argument("--max-ot", dest="max_ot", type=int, default=2000)
    ap.add_argument("--cpu-guides", type=int, default=2048, help="guides in the cpu_baseline sample (single thread)")
    ap.add_argument("--ref-guides", type=int, default=2048, help="guides per step of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bulge", action="store_true", help="skip the configs[3] (bulge extension) measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
