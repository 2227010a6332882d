#!/usr/bin/env python3
"""bench.py -- guides/sec of FlashFry's off-target discovery hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N --steps K --warmup W]             # this repo's CUDA path (N > 1: under torchrun, one rank per GPU)
    python bench.py --impl reference [...]                      # the reference's CPU algorithm (oracle port)
    python bench.py --workload fused|bulge                      # BASELINE.json configs[4] / configs[3] as first-class lines
    python bench.py --scaling weak                              # the full batch on every GPU instead of one sharded batch
    python bench.py --single-process --gpus N                   # every GPU behind ONE process through the C ABI's ff_multi
    python bench.py --shard guides|database                     # N > 1: what is sharded (auto: the index work from 8 GPUs on)

Workload (BASELINE.json configs[2], named in config.workload): 100 000 synthetic 20-bp NGG guides IN TOTAL (+10 % planted
near real targets), sharded over the ranks ("scaling": "strong"), against a synthetic human-genome-sized (3e8 distinct
targets) spCas9-NGG index, <= 4 mismatches, maximumOffTargets 2000.  The index is generated in HBM from a seed; one
replica per GPU.  A "step" = one discover call over the rank's shard + one all-gather of the per-guide totals; with
--shard database every rank scans 1/N of the index for ALL guides and the candidates, the barriers and the all-gather of the
totals go through NVLink peer memory (flashfry_b200/csrc/ff_shard.inl), each rank ending with the rows of its own guides.

`value`  : whole-job guides/s with guides already in HBM and results left in HBM (CUDA events, max over ranks).
`e2e`    : the same metric through the C ABI with HOST buffers (ff_discover: H2D of the guides, D2H of the hit lists);
           `e2e.with_compact_hits`: the same with 32-bit database indices instead of target longs.
`roofline`: the two scan kernels (k_bin_scan + k_pair_scan, one launch per index half, timed with CUDA events on the
            launching stream).  `achieved` / `frac` use SURVEY.md 8(d)'s ALGORITHMIC bytes (8 N_t + 8 G + 16 H per call)
            against the measured HBM copy bandwidth; `traffic` / `dram_frac` = the HBM bytes ncu measured for the same call
            (profiles/scan_traffic.json); `issue_frac`, `alu_pipe_frac` from the same capture; `compare_frac` against SURVEY's
            one-POPC-per-entry ceiling; `requested_bytes` = what the kernels ask of L2 / shared memory (NOT HBM); `bound`
            names what binds ("issue" for the bin-major kernels); `launches` has one entry per kernel.
Side measurements on rank 0 at N = 1 (never inside the timed regions): both scan kernels agree on the whole batch,
`fused_discover_score`, `bulge_mode`, `batch_ladder` (1 .. 100 000 guides x k = 3, 4, 5 next to the published JVM table),
`chr22_1000_guides`, `skewed_index`, `e2e_compact_hits`, `cold_start_flashfry_format` (the index written as a real
FlashFry database and loaded back), `cpu_baseline` (the oracle port on one host thread, bounded sample, doubles as a
parity check of the timed path).
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


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def family_consensus(seed, f):
    """The 21-base (protospacer + N) consensus of repeat family f of ff_synth_database_skewed (ff_db.cu k_synth_families)."""
    with np.errstate(over="ignore"):
        f = np.asarray(f, dtype=np.uint64)
        return _splitmix64(np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0xFA111E5) + f) >> np.uint64(64 - 42)


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


BASELINE_JVM = {  # FlashFry 1.8.1 JVM, ONE core, real hg38 (paper/timing_data/bwa_flashfry/*/runtime_set*.tar.gz),
    # median wall-clock seconds of a whole `discover` run (database load included), by max mismatches and guide count
    3: {1: 7.6, 100: 36.4, 1000: 44.1, 10000: 81.8, 100000: 514.0},
    4: {1: 15.9, 100: 41.7, 1000: 61.6, 10000: 210.6, 100000: 1861.0},
    5: {1: 24.5, 100: 50.1, 1000: 106.0, 10000: 628.2, 100000: 7924.0},
}


def _metric_name(workload, k):
    if workload == "fused":
        return "guides/sec, discover (<=%d mismatches) + CFD + Hsu2013 fused on the GPU, vs hg38-sized index" % k
    if workload == "bulge":
        return "guides/sec at <=5 mismatches + one 1-bp RNA/DNA bulge vs hg38-sized index (extension: no reference semantics)"
    return "guides/sec at <=%d mismatches vs hg38-sized index" % k


def run_native(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import flashfry_b200.api as ff
    from flashfry_b200 import _native as N
    from flashfry_b200.sharding import shard_range, all_gather_counts

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    wl = args.workload
    n_total = args.guides if args.guides else (50_000 if wl == "fused" else 100_000)
    k = 5 if wl == "bulge" else args.k

    ctx = ff.Context(local)
    t0 = time.perf_counter()
    ctx.synth_database(ENZYME, args.targets, SEED_DB)
    torch.cuda.synchronize()
    db_s = time.perf_counter() - t0
    info = ctx.info()
    n_t = int(info.n_targets)
    # planted guides are drawn from a few slices of the resident database (identical replicas: identical on every rank)
    rng = np.random.default_rng(17)
    pool = np.concatenate([ctx.copy_targets(int(s), 2048) for s in rng.integers(0, max(1, n_t - 2048), 16)])
    if args.scaling == "strong":   # configs[2]: ONE guide set, sharded over the ranks
        all_guides = make_guides(n_total, SEED_GUIDES, pool, SEED_PLANTED)
        lo, hi = shard_range(len(all_guides), rank, world)
        guides, G_job = all_guides[lo:hi], len(all_guides)
    else:                          # weak: every rank its own batch of the full size
        guides = make_guides(n_total, SEED_GUIDES + 1000 * rank, pool, SEED_PLANTED + rank)
        G_job = world * len(guides)
    G = len(guides)

    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    d_guides = torch.from_numpy(guides.view(np.int64)).to(dev)

    # N > 1, one guide set: shard the INDEX WORK instead of the guides (ff_shard.inl) -- every rank scans 1/N of the index
    # for all guides and the scan kernels push the candidates to the guide's owner over NVLink peer memory.  The exchange
    # blocks are mapped with CUDA IPC handles exchanged once through torch.distributed; any rank that cannot map them
    # sends every rank back to guide sharding.
    db_sharded, d_all, shard_note = False, None, None
    # auto: measured on 8 x B200 (profiles/r3_scale_*.jsonl) the sharded index work wins on device time from 4 GPUs on (+3 % at 4,
    # +16 % at 8, fused +46 % at 8) but loses end to end (its barriers make all ranks copy their rows to the host at the same
    # moment: 106 MB into one socket); it is the default from 8 GPUs on.
    want_db = args.shard == "database" or (args.shard == "auto" and world >= 8)
    if world > 1 and args.scaling == "strong" and wl in ("discover", "fused") and want_db:
        ok = 1
        try:
            hit_cap = max(1 << 24, int(2.0 * 130.0 * G_job / world))
            handle = ctx.peer_export(hit_cap, max(G_job, 1 << 20))
        except Exception as e:  # noqa: BLE001
            ok, shard_note, handle = 0, "peer_export failed: %s" % e, b"\0" * 64
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        if ok:
            try:
                ctx.peer_attach(rank, world, handles)
            except Exception as e:  # noqa: BLE001
                ok, shard_note = 0, "peer_attach failed: %s" % e
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        db_sharded = bool(flag.item())
        if db_sharded:
            d_all = torch.from_numpy(all_guides.view(np.int64)).to(dev)
            try:  # one trial step: a rank that fails (its peers then leave their barriers after 4 s) sends everyone to guide sharding
                ctx.discover_sharded_device(d_all.data_ptr(), G_job, k, args.max_ot, 0)
            except Exception as e:  # noqa: BLE001
                ok, shard_note = 0, "trial step failed: %s" % e
            flag = torch.tensor([ok], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            db_sharded = bool(flag.item())
            if not db_sharded:
                ctx.peer_detach()
                shard_note = shard_note or "the trial step failed on another rank"
        if not db_sharded and args.shard == "database":
            raise SystemExit("--shard database: %s" % (shard_note or "a peer rank could not map the exchange blocks"))

    class _DevView:  # wrap a context-owned device pointer for torch without copying
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}

    counts_weak = torch.empty(world * G, dtype=torch.int32, device=dev) if (world > 1 and args.scaling == "weak") else None

    def step():
        if db_sharded:  # candidates, barriers and the all-gather of the totals all go through peer memory: no NCCL call
            return ctx.discover_sharded_device(d_all.data_ptr(), G_job, k, args.max_ot, 3 if wl == "fused" else 0)
        if wl == "bulge":
            r = ctx.discover_bulge_device(d_guides.data_ptr(), G, k, args.max_ot, 3)
        else:
            r = ctx.discover_device(d_guides.data_ptr(), G, k, args.max_ot, 3 if wl == "fused" else 0)
        if world > 1:  # the one collective of the path: every rank learns every guide's total count
            mine = torch.as_tensor(_DevView(r.d_total_count, G), device=dev) if G else torch.zeros(0, dtype=torch.int32, device=dev)
            if args.scaling == "strong":
                all_gather_counts(mine, G_job)
            else:
                dist.all_gather_into_tensor(counts_weak, mine)
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
    tms, launches = [], 0
    sync_all()
    e0.record(stream)
    for _ in range(args.steps):
        r = step()
        tm = ctx.timings()
        tms.append((tm.scan_ms, tm.scan_part1_ms, tm.scan_part2_ms, tm.prep_ms, tm.order_ms, tm.cut_ms, tm.score_ms))
        launches += tm.kernel_launches + (1 if world > 1 and not db_sharded else 0)
    e1.record(stream)
    sync_all()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tmax = torch.tensor([ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    ms_per_step = ms / args.steps
    value = G_job / (ms_per_step / 1e3)
    n_hits, cand = int(r.n_hits), int(r.n_candidate_hits)
    ent1, ent2, req_bytes, scan_launches = int(tm.entries_part1), int(tm.entries_part2), int(tm.scan_bytes_read), int(tm.scan_launches)
    sharded_ok = None
    if db_sharded:  # the sharded rows of this rank's guides against the single-GPU call on the same shard, all ranks
        def _rows(res):
            v = lambda ptr, n, ts: torch.as_tensor(_Dev(ptr, n, ts), device=dev).clone()  # noqa: E731
            return (v(res.d_row_ptr, G + 1, "<i8"), v(res.d_targets, max(int(res.n_hits), 1), "<i8")[:int(res.n_hits)],
                    v(res.d_mismatches, max(int(res.n_hits), 1), "|u1")[:int(res.n_hits)], v(res.d_total_count, max(G, 1), "<i4")[:G])

        class _Dev:
            def __init__(self, ptr, n, ts):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": ts, "data": (ptr, False), "version": 2}
        a_rows = _rows(r)
        tot_all = torch.as_tensor(_Dev(ctx.peer_totals_ptr(), G_job, "<i4"), device=dev).clone()
        b_rows = _rows(ctx.discover_device(d_guides.data_ptr(), G, k, args.max_ot, 0))
        same = all(x.shape == y.shape and torch.equal(x, y) for x, y in zip(a_rows, b_rows)) and torch.equal(tot_all[lo:hi], b_rows[3])
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        sharded_ok = bool(flag.item())

    # ---- e2e through the C ABI with HOST buffers (pinned guides in, hit lists out), same shard, same steps
    pinned = torch.from_numpy(guides.view(np.int64)).pin_memory()
    g_host = pinned.numpy().view(np.uint64)
    hp = C.POINTER(N.FFHits)()
    gp = g_host.ctypes.data_as(C.POINTER(C.c_uint64))
    if db_sharded:
        pinned_all = torch.from_numpy(all_guides.view(np.int64)).pin_memory()
        gp_all = pinned_all.numpy().view(np.uint64).ctypes.data_as(C.POINTER(C.c_uint64))
    sc = [np.zeros(max(G, 1)) for _ in range(3)]
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731

    def e2e_step():
        if db_sharded and wl == "discover":  # H2D of ALL guides on every rank, D2H of the rank's own rows
            N.check(N.lib().ff_discover_sharded(ctx._h, gp_all, G_job, k, args.max_ot, C.byref(hp)))
        elif wl == "fused":
            N.check(N.lib().ff_discover_score(ctx._h, gp, G, k, args.max_ot, 0, 3, C.byref(hp), dp(sc[0]), dp(sc[1]), dp(sc[2])))
        elif wl == "bulge":
            N.check(N.lib().ff_discover_bulge(ctx._h, gp, G, k, args.max_ot, 3, 0, C.byref(hp)))
        else:
            N.check(N.lib().ff_discover(ctx._h, gp, G, k, args.max_ot, 0, C.byref(hp)))
        nh = int(hp.contents.n_hits)
        N.lib().ff_hits_free(hp)
        return nh

    def time_e2e(steps):
        for _ in range(max(2, args.warmup // 2 + 1)):
            nh = e2e_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(steps):
            nh = e2e_step()
        sync_all()
        s = (time.perf_counter() - t0) / steps
        if world > 1:
            tmax = torch.tensor([s], device=dev, dtype=torch.float64)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            s = float(tmax.item())
        return s, nh
    e2e_steps = args.steps if wl != "bulge" else max(2, args.steps // 5)
    e2e_s, nh = time_e2e(e2e_steps)
    e2e_compact_s = None
    if wl == "discover":  # the same with 32-bit database indices instead of target longs (5 instead of 9 bytes per hit)
        ctx.set_option("compact_hits", 1)
        e2e_compact_s, _ = time_e2e(e2e_steps)
        ctx.set_option("compact_hits", 0)
    clocks = sampler.stop(t_region0, sampler.mark()) if rank == 0 else None  # samples taken inside the two timed regions
    per_hit = 10 if wl == "bulge" else 9
    e2e = {"value": G_job / e2e_s, "unit": "guides/s", "h2d_bytes_per_step": 8 * (G_job if db_sharded and wl == "discover" else G),
           "d2h_bytes_per_step": (G + 1) * 8 + nh * per_hit + G * 5 + (24 * G if wl == "fused" else 0), "ms_per_step": e2e_s * 1e3,
           "bytes_are": "per rank", "over_device_step": e2e_s * 1e3 / ms_per_step,
           "call": ("ff_discover_sharded (all guides in, this rank's rows out)" if db_sharded and wl == "discover" else
                    "ff_discover_score on the rank's guides" if wl == "fused" else
                    "ff_discover_bulge on the rank's guides" if wl == "bulge" else "ff_discover on the rank's guides")}
    if e2e_compact_s:
        e2e["with_compact_hits"] = {"value": G_job / e2e_compact_s, "ms_per_step": e2e_compact_s * 1e3,
                                    "d2h_bytes_per_step": (G + 1) * 8 + nh * 5 + G * 5,
                                    "note": "option compact_hits: ff_hits.target_index (u32) instead of ff_hits.targets (u64); longs looked up on "
                                            "demand in the host mirror (ff_db_host_targets), as ff_hits_write_tsv and the JVM glue do"}

    out = None
    if rank == 0:
        peak, peak_src = measured_peak()
        a = np.asarray(tms, dtype=np.float64).mean(axis=0)
        scan_ms, p1_ms, p2_ms = float(a[0]), float(a[1]), float(a[2])
        bin_major = p2_ms > 0.0
        # ALG_BYTES = 8 N_t (every target word once) + 8 G (guides) + 16 H (compact hit records)
        alg_bytes = 8 * n_t + 8 * G + 16 * cand
        achieved = alg_bytes / (scan_ms / 1e3) / 1e9
        tr = ncu_traffic()
        tr_ok = bool(tr and tr.get("targets") == n_t and tr.get("max_mismatch") == k and tr.get("guides_in_profiled_launch") == G
                     and tr.get("kernel") == ("k_bin_scan+k_pair_scan" if bin_major else "k_seed_scan"))
        traffic = tr.get("dram_bytes_per_profiled_launch") if tr_ok else None
        roof = {"bound": (tr.get("bound") if tr_ok else None) or ("issue" if bin_major else "hbm"),
                "kernel": "k_bin_scan + k_pair_scan (the bin-major scan: one launch per index half)" if bin_major else "k_seed_scan",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "alg_bytes_8d": alg_bytes, "alg_bytes_are": "SURVEY.md 8(d): 8*N_t + 8*G + 16*H per call",
                "kernel_ms": scan_ms, "kernel_share_of_step": scan_ms / (float(a[[0, 3, 4, 5, 6]].sum())),
                "traffic": traffic, "dram_frac": (traffic / (scan_ms / 1e3) / 1e9 / peak) if traffic else None,
                "requested_bytes": req_bytes, "requested_bytes_are": "index lookups + bit-sliced planes streamed + guides + hit keys (L2 / shared-memory level, not HBM)",
                "entries_compared_per_launch": ent1 + ent2, "entries_per_guide": (ent1 + ent2) / max(G, 1),
                # SURVEY 8(d)'s second ceiling -- one 32-bit POPC per compared entry, 16 lanes/clk/SM x 148 SMs x 1.9 GHz.  The
                # bit-sliced kernels do not use the POPC pipe (40 LOP3 per 32 entries), so this fraction may exceed 1.
                "compares_per_s": (ent1 + ent2) / (scan_ms / 1e3), "compare_ceiling_per_s": 4.5e12,
                "compare_frac": (ent1 + ent2) / (scan_ms / 1e3) / 4.5e12,
                "launches": ([{"kernel": "k_bin_scan<9> (index A through shared memory, TMA-staged bins)", "ms": p1_ms, "entries": ent1,
                               "compares_per_s": ent1 / (p1_ms / 1e3)},
                              {"kernel": "k_pair_scan<11> (index B, pairs sorted by bucket)", "ms": p2_ms, "entries": ent2,
                               "compares_per_s": ent2 / (p2_ms / 1e3)}] if bin_major else
                             [{"kernel": "k_seed_scan", "ms": scan_ms, "entries": ent1 + ent2}]),
                "step_breakdown_ms": {"prep": float(a[3]), "scan": scan_ms, "order": float(a[4]), "cut": float(a[5]), "score": float(a[6])}}
        if tr_ok:
            roof["traffic_source"] = tr.get("source")
            for key in ("issue_frac", "alu_pipe_frac", "per_launch"):
                if key in tr:
                    roof[key] = tr[key]
        out = {"metric": _metric_name(wl, k), "value": value, "unit": "guides/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
               "scaling": args.scaling, "vs_baseline": (value / 53.7) if wl == "discover" and k == 4 else None, "dtype": "u64", "data": "synthetic",
               "config": {"workload": "%s: %d synthetic NGG guides (+10%% planted) %s vs synthetic %d-target spCas9-NGG index, k<=%d%s, maxOT %d" % (
                              {"discover": "configs[2]", "fused": "configs[4] (discover + CFD + Hsu2013 fused)", "bulge": "configs[3] (1-bp bulge extension)"}[wl],
                              G_job, ("in total, over %d GPU(s)" % world) if args.scaling == "strong" else "= %d per GPU" % G,
                              n_t, k, " + one 1-bp RNA/DNA bulge" if wl == "bulge" else "", args.max_ot),
                          "guides_total": G_job, "guides_per_gpu": G, "targets": n_t, "max_mismatch": k, "maximum_off_targets": args.max_ot,
                          "parallelism": ("database-sharded x%d: one index replica per GPU, every rank scans 1/%d of the index for ALL guides, the scan kernels push "
                                          "candidates to the guide's owner over NVLink peer memory (P2P stores + remote atomics), barriers and the all-gather of "
                                          "the totals through the same exchange blocks (no NCCL on the data path)" % (world, world)) if db_sharded else
                                         "guide-sharded x%d, one index replica per GPU, one NCCL all-gather of int32 totals per step" % world,
                          "sharded_rows_equal_single_gpu_rows": sharded_ok, "shard_fallback_reason": shard_note,
                          "l2": "index (%.1f GB) is larger than L2, re-streamed every step" % (info.device_bytes / 1e9),
                          "seed_split": "first %d | last %d protospacer bases" % (int(info.seed_split_a), 20 - int(info.seed_split_a)),
                          "db_build_s": db_s},
               "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof,
               "hits_per_step": n_hits, "candidate_hits_per_step": cand, "scan_launches_last_step": scan_launches,
               "baseline_note": "vs_baseline = value / 53.7 guides/s (published single-core JVM, 100k guides, hg38, k<=4; BASELINE.md)"}
        if wl == "discover" and not args.no_extras:
            _extras(out, ctx, ff, N, C, torch, dev, args, guides, d_guides, g_host, gp, hp, G, k, n_t, n_hits, world, _DevView)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        _emit(out)


def run_single_process(args):
    """--single-process: every GPU behind ONE host process through ff_multi (what a JVM host would do): guide-sharded
    ff_multi_discover with host buffers, one ncclAllGather of the totals inside the library.  End-to-end only."""
    import ctypes as C
    import torch
    import flashfry_b200.api as ff
    n = args.gpus
    if torch.cuda.device_count() < n:
        raise SystemExit("--single-process --gpus %d needs %d visible GPUs" % (n, n))
    mc = ff.MultiContext(list(range(n)))
    if n > 1 and args.shard != "guides":
        mc.set_option("shard_mode", 1)  # shard the index work; candidates reach the guide's owner over NVLink peer memory
    t0 = time.perf_counter()
    mc.synth_database(ENZYME, args.targets, SEED_DB)
    db_s = time.perf_counter() - t0
    with ff.Context(0) as c0:  # the guides of the bench line (pool drawn from an identical replica)
        c0.synth_database(ENZYME, args.targets, SEED_DB)
        n_t = int(c0.info().n_targets)
        rng = np.random.default_rng(17)
        pool = np.concatenate([c0.copy_targets(int(s), 2048) for s in rng.integers(0, max(1, n_t - 2048), 16)])
    guides = make_guides(args.guides or 100_000, SEED_GUIDES, pool, SEED_PLANTED)
    G = len(guides)
    pinned = torch.from_numpy(guides.view(np.int64)).pin_memory()
    gp = pinned.numpy().view(np.uint64).ctypes.data_as(C.POINTER(C.c_uint64))
    totals = np.zeros(G, np.int32)
    sampler = ClockSampler(0)
    sampler.start()
    for _ in range(max(3, args.warmup)):
        hits = mc.discover_raw(gp, G, args.k, args.max_ot, totals)
    t_region0 = sampler.mark()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hits = mc.discover_raw(gp, G, args.k, args.max_ot, totals)
    s = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop(t_region0, sampler.mark())
    dev_ms = max(mc.rank_timings(r).total_ms for r in range(n))
    _emit({"metric": _metric_name("discover", args.k), "value": G / s, "unit": "guides/s", "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": s * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": G / s / 53.7, "dtype": "u64", "data": "synthetic",
           "config": {"workload": "configs[2]: %d guides in total through ff_multi_discover (ONE host process, one thread per GPU, host buffers in "
                                  "and out, one ncclAllGather of the totals inside the library) vs %d targets, k<=%d, maxOT %d" % (G, n_t, args.k, args.max_ot),
                      "launcher": "single process (ff_multi)", "targets": n_t, "db_build_s_all_devices": db_s,
                      "shard_mode": "database (NVLink peer memory)" if (n > 1 and args.shard != "guides") else "guides (ncclAllGather of totals)"},
           "e2e": {"value": G / s, "unit": "guides/s", "h2d_bytes_per_step": 8 * G, "d2h_bytes_per_step": (G + n) * 8 + hits * 9 + G * 5,
                   "ms_per_step": s * 1e3, "bytes_are": "whole job"},
           "slowest_rank_device_ms_last_step": dev_ms, "hits_per_step": hits, "totals_sum": int(totals.sum()), "clocks": clocks,
           "note": "value == e2e here: this mode only exists end to end"})
    mc.close()


def _extras(out, ctx, ff, N, C, torch, dev, args, guides, d_guides, g_host, gp, hp, G, k, n_t, n_hits, world, _DevView):
    """Side measurements on rank 0's shard: never inside the timed regions, never allowed to take the bench line down."""
    def guard(name, fn):
        try:
            out[name] = fn()
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": "%s: %s" % (type(e).__name__, e)}

    # the two scan kernels (guide-major k_seed_scan / bin-major k_bin_scan + k_pair_scan) must agree on the whole timed batch
    def _totals(kernel):
        ctx.set_option("scan_kernel", kernel)
        rr = ctx.discover_device(d_guides.data_ptr(), G, k, args.max_ot, 0)
        t = torch.as_tensor(_DevView(rr.d_total_count, G), device=dev).clone()
        ctx.set_option("scan_kernel", 0)
        return int(rr.n_hits), t
    h0, t0_ = _totals(1)
    h1, t1_ = _totals(2)
    out["scan_kernels_agree_on_full_batch"] = bool(h0 == h1 == n_hits and torch.equal(t0_, t1_))

    def time_call(fn, reps):
        fn(); fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def e2e_call(kk=k, n=G):
        N.check(N.lib().ff_discover(ctx._h, gp, n, kk, args.max_ot, 0, C.byref(hp)))
        N.lib().ff_hits_free(hp)

    def compact():
        # the hit list as 32-bit database indices (5 instead of 9 bytes per hit over PCIe); the host mirrors the target array
        # once and looks longs up on demand (ff_db_host_targets / ff_hits_resolve)
        t0 = time.perf_counter()
        N.lib().ff_db_host_targets(ctx._h)
        mirror_s = time.perf_counter() - t0
        ctx.set_option("compact_hits", 1)
        try:
            s = time_call(e2e_call, max(3, args.steps // 2))

            def resolved():
                N.check(N.lib().ff_discover(ctx._h, gp, G, k, args.max_ot, 0, C.byref(hp)))
                N.check(N.lib().ff_hits_resolve(ctx._h, hp))
                N.lib().ff_hits_free(hp)
            s_res = time_call(resolved, 3)
        finally:
            ctx.set_option("compact_hits", 0)
        return {"ms_per_step": s * 1e3, "guides_per_s": G / s, "d2h_bytes_per_step": (G + 1) * 8 + n_hits * 5 + G * 5,
                "with_ff_hits_resolve_ms": s_res * 1e3, "host_mirror_build_s": mirror_s,
                "note": "option compact_hits: ff_hits.target_index instead of ff_hits.targets; resolve = all longs looked up on the host"}
    guard("e2e_compact_hits", compact)

    def fused():
        fs = []
        for _ in range(3):
            ctx.discover_device(d_guides.data_ptr(), G, k, args.max_ot, 3)
            t = ctx.timings()
            fs.append((t.total_ms, t.score_ms))
        return {"guides": G, "total_ms": float(np.median([a for a, _ in fs])), "score_kernel_ms": float(np.median([b for _, b in fs])),
                "guides_per_s": G / (float(np.median([a for a, _ in fs])) / 1e3),
                "metrics": "DoenchCFD_maxOT, DoenchCFD_specificityscore, Hsu2013 (FP64, bit-identical to the oracle)",
                "first_class_line": "python bench.py --workload fused  (configs[4]: 50 000 guides)"}
    guard("fused_discover_score", fused)

    # BASELINE.json configs[3]: <= 5 mismatches + one 1-bp RNA/DNA bulge on the same batch.  An EXTENSION: the reference
    # has no bulge search, so this line is never part of a parity claim (semantics: include/flashfry_b200.h).
    def bulge():
        bs = []
        for _ in range(3):
            rb = ctx.discover_bulge_device(d_guides.data_ptr(), G, 5, args.max_ot, 3)
            t = ctx.timings()
            bs.append((t.total_ms, t.scan_ms, t.scan_launches))
        bms = float(np.median([a for a, _, _ in bs]))
        return {"workload": "configs[3]: %d guides, <=5 mismatches + one 1-bp RNA or DNA bulge, maxOT %d, same index" % (G, args.max_ot),
                "total_ms": bms, "scan_ms": float(np.median([b for _, b, _ in bs])), "windows": int(bs[-1][2]),
                "guides_per_s": G / (bms / 1e3), "hits": int(rb.n_hits), "candidate_hits": int(rb.n_candidate_hits),
                "parity": "extension, no reference semantics: checked against a brute-force definition in tests/test_gpu_bulge.py",
                "first_class_line": "python bench.py --workload bulge"}
    if not args.no_bulge:
        guard("bulge_mode", bulge)

    # the batch-size ladder the reference publishes: device-resident and end-to-end, per k
    def ladder():
        rows = []
        for kk in (3, 4, 5):
            for n in (1, 100, 1000, 10000, 100000):
                if n > G:
                    continue
                reps = 20 if n <= 10000 else (5 if kk < 5 else 2)
                dv = time_call(lambda: ctx.discover_device(d_guides.data_ptr(), n, kk, args.max_ot, 0), reps)
                ee = time_call(lambda: e2e_call(kk, n), reps)
                jvm = BASELINE_JVM[kk].get(n)
                rows.append({"k": kk, "guides": n, "device_ms": dv * 1e3, "e2e_ms": ee * 1e3, "e2e_guides_per_s": n / ee,
                             "jvm_published_whole_run_s": jvm})
        return {"rows": rows, "one_guide_e2e_latency_ms": [r["e2e_ms"] for r in rows if r["guides"] == 1 and r["k"] == 4][0],
                "note": "here the database is resident (cold start: cold_start_flashfry_format); the published JVM column is a whole run incl. its "
                        "database load, one core, real hg38 (BASELINE.md section 1)"}
    if not args.no_ladder:
        guard("batch_ladder", ladder)

    # BASELINE.json configs[1]: 1 000 synthetic + 100 planted guides vs the chr22 quick-start database, when the database
    # built by __graft_entry__.build() travelled with the tree (33.5 MB of targets: L2-resident after the first pass)
    def chr22():
        path = os.path.join(ROOT, "tests", "golden", "_chr22", "chr22_cas9ngg_database")
        if not (os.path.exists(path) and os.path.exists(path + ".header")):
            return {"skipped": "chr22 database not in the tree"}
        c2 = ff.Context(dev.index)
        t0 = time.perf_counter()
        c2.load_database(path)
        load_s = time.perf_counter() - t0
        t22 = c2.copy_targets()
        g22 = make_guides(1100, 1001, t22[:: max(1, len(t22) // 4096)], 1002, planted_frac=100.0 / 1100.0)
        d22 = torch.from_numpy(g22.view(np.int64)).to(dev)
        ts = []
        for _ in range(8):
            r22 = c2.discover_device(d22.data_ptr(), len(g22), k, args.max_ot, 0)
            ts.append(c2.timings().total_ms)
        res = {"workload": "configs[1]: 1 000 synthetic + 100 planted guides vs the chr22 quick-start index (%d targets), k<=%d" % (len(t22), k),
               "total_ms": float(np.median(ts[2:])), "guides_per_s": len(g22) / (float(np.median(ts[2:])) / 1e3),
               "hits": int(r22.n_hits), "cold_load_s": load_s,
               "parity": "tests/test_gpu_parity.py::test_chr22_1000_guides_config (bit-exact vs the oracle)"}
        c2.close()
        return res
    guard("chr22_1000_guides", chr22)

    # Genome-like skew: 64 repeat neighbourhoods of 16 384 targets (consensus + 0..3 substitutions) inside the same-sized
    # index, 1 % of the guides planted in them (thousands of candidates each: long segments, hot buckets, buffer growth)
    def skewed():
        c3 = ff.Context(dev.index)
        c3.synth_database_skewed(ENZYME, args.targets, SEED_DB, 64, 16384, 3)
        n3 = int(c3.info().n_targets)
        rng = np.random.default_rng(5)
        pool3 = np.concatenate([c3.copy_targets(int(s), 2048) for s in rng.integers(0, max(1, n3 - 2048), 16)])
        g3 = make_guides(G, SEED_GUIDES, pool3, SEED_PLANTED, planted_frac=0.09)
        # 1 % of the guides sit inside a neighbourhood: the family's consensus (the generator's formula, ff_db.cu
        # k_synth_families) with 0..2 substitutions
        n_pl = max(1, G // 100)
        cons = family_consensus(SEED_DB, rng.integers(0, 64, n_pl))
        for _ in range(2):
            pos = rng.integers(0, 20, n_pl).astype(np.uint64)
            x = rng.integers(0, 4, n_pl).astype(np.uint64) << (np.uint64(2) * (np.uint64(20) - pos))
            cons = cons ^ x
        g3[:n_pl] = (cons << np.uint64(4)) | np.uint64(0xA) | (np.uint64(1) << np.uint64(48))
        d3 = torch.from_numpy(g3.view(np.int64)).to(dev)
        ts = []
        for _ in range(5):
            r3 = c3.discover_device(d3.data_ptr(), len(g3), k, args.max_ot, 0)
            t = c3.timings()
            ts.append((t.total_ms, t.scan_ms, t.order_ms + t.cut_ms, t.scan_launches))
        res = {"workload": "%d guides (1 %% planted inside repeat neighbourhoods) vs %d targets with 64 neighbourhoods of 16 384 (consensus + 0..3 substitutions)" % (len(g3), n3),
               "total_ms": float(np.median([x[0] for x in ts[1:]])), "scan_ms": float(np.median([x[1] for x in ts[1:]])),
               "order_cut_ms": float(np.median([x[2] for x in ts[1:]])), "scan_launches_steady_state": int(ts[-1][3]),
               "hits": int(r3.n_hits), "candidate_hits": int(r3.n_candidate_hits), "planted_in_neighbourhoods": int(n_pl),
               "guides_per_s": len(g3) / (float(np.median([x[0] for x in ts[1:]])) / 1e3),
               "parity": "tests/test_gpu_scale.py::test_skewed_index_at_benchmark_size (2 048-guide oracle sample)"}
        c3.close()
        return res
    if not args.no_skew:
        guard("skewed_index", skewed)

    # cold start at this size, in FlashFry's OWN on-disk format (BGZF body + .header, all bins indexed) and from the image side-car
    def cold():
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import cold_start
        return cold_start.measure(args.targets)
    if not args.no_cold_start and world == 1:
        guard("cold_start_flashfry_format", cold)

    if not args.no_cpu_baseline and world == 1:
        guard("cpu_baseline", lambda: cpu_baseline(ctx, guides, args, threads=1, budget_guides=args.cpu_guides))


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
    guides = make_guides(args.guides or 100_000, SEED_GUIDES, pool, SEED_PLANTED)
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
                      "higher_is_better": True, "scaling": args.scaling, "vs_baseline": v / 53.7, "dtype": "u64", "data": "synthetic",
                      "config": {"workload": "configs[2] bounded sample: " + sample_txt, "targets": n_t, "max_mismatch": args.k,
                                 "maximum_off_targets": args.max_ot,
                                 "same_config_note": "rate over %d-guide batches (a 100 000-guide CPU step would take ~40 s); always ALL host threads, "
                                                     "whatever --gpus says; the index is materialised by the product's generator, the timed region is the oracle only" % n},
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
    ap.add_argument("--guides", type=int, default=0, help="guides in the job (default: 100 000; 50 000 for --workload fused)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: ONE guide set sharded over the GPUs (BASELINE configs[2]); weak: the full batch on every GPU")
    ap.add_argument("--workload", default="discover", choices=["discover", "fused", "bulge"],
                    help="discover = configs[2]; fused = configs[4] (discover + CFD + Hsu2013 on the GPU, 50 000 guides); bulge = configs[3]")
    ap.add_argument("--no-extras", action="store_true", help="only the headline line (no side measurements)")
    ap.add_argument("--shard", default="auto", choices=["auto", "guides", "database"],
                    help="N > 1, strong scaling: shard the guides (NCCL all-gather of totals) or the index work (NVLink peer memory); auto = database from 8 GPUs on")
    ap.add_argument("--single-process", action="store_true",
                    help="all --gpus behind ONE process through the C ABI's ff_multi (not the torchrun contract): end-to-end line only")
    ap.add_argument("--no-ladder", action="store_true")
    ap.add_argument("--no-skew", action="store_true")
    ap.add_argument("--no-cold-start", action="store_true", help="skip writing + loading the index as a real FlashFry database (~40 s)")
    ap.add_argument("--k", type=int, default=4)
    ap.add_argument("--max-ot", dest="max_ot", type=int, default=2000)
    ap.add_argument("--cpu-guides", type=int, default=2048, help="guides in the cpu_baseline sample (single thread)")
    ap.add_argument("--ref-guides", type=int, default=2048, help="guides per step of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bulge", action="store_true", help="skip the configs[3] (bulge extension) measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.single_process:
        run_single_process(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
