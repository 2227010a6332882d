// The general discover path (included by ff_discover.cu, inside namespace ff): bulge patterns and database-order windows.
//
// 1. 1-bp bulge mode -- an EXTENSION: FlashFry has no gap / bulge / edit-distance search (BASELINE.json
//    configs[3]); the semantics are defined in include/flashfry_b200.h (ff_discover_bulge); the tests compare this path with a
//    brute-force statement of that definition.  Every alignment with one looped-out base is a 20-position TEMPLATE over the
//    stored protospacer with one wildcard position:
//        RNA bulge at q : t[0] = *, t[j] = g[j-1] (1 <= j <= q), t[j] = g[j] (j > q)
//        DNA bulge at q : t[j] = g[j+1] (j < q), t[q] = *, t[j] = g[j] (j > q)
//    so the bulge search is the mismatch search of 1 + 18 + 18 templates per guide through the same two-part seed index:
//    a wildcard inside the seeded key multiplies the seeds by four (mask over the other key bases x value of the wildcard
//    base), a wildcard inside the verified part is dropped from the count by an AND mask.  Hits of different templates
//    are merged (sort + unique) and the best alignment of each surviving pair is recomputed when the rows are gathered.
// 2. Database-order windows -- what OrderedBinTraversalFactory's overflow callback does for the reference
//    (reference/traversal/OrderedBinTraversalFactory.scala:107-119: overflowed guides leave the traversal): when a
//    guide is expected to collect more than maximumOffTargets occurrences, the database is scanned in windows of
//    index-A cells in database order, the overflow cut is applied after every window with the running totals, and
//    guides that are full are dropped from the later windows.  Results are identical to a whole-database scan.

constexpr int kMaxPatterns = 40;

struct PatternPlan {
  int n_patterns;
  int items_per_guide;
  int item0[kMaxPatterns + 1];                                   // first work item of every pattern (prefix sums)
  uint8_t type[kMaxPatterns], q[kMaxPatterns], cls[kMaxPatterns]; // type 0 none / 1 RNA / 2 DNA; class = where the wildcard sits
  struct Cls { int hA, nA, nB, itemsA, itemsB, spiB; } c[3];     // 0: no wildcard, 1: in the part-one key, 2: in the part-two key
};

struct GeneralParams {
  ScanParams sp;
  PatternPlan pl;
  const uint32_t *active;  // guide indices still collecting (nullptr = all guides)
  int64_t n_active;
  const uint32_t *a_masks_w1, *b_masks_w1;
  int a_bases, P;
  int c0, c1, cell_shift;
  const uint32_t *cell_off;  // nullptr = the whole database in one window
};

// The template a pattern lays over the stored protospacer (wildcard digit = 0) and the wildcard position (-1: none).
__device__ __forceinline__ uint64_t pattern_template(uint64_t proto, int type, int q, int P, int *w) {
  if (type == 0) { *w = -1; return proto; }
  const uint64_t lo_mask = (1ull << (2 * (P - 1 - q))) - 1ull;  // positions q+1 .. P-1 pair with g[j]
  if (type == 1) {                                               // positions 1 .. q pair with g[j-1]; position 0 is free
    *w = 0;
    return ((proto >> 2) & ~lo_mask & ((1ull << (2 * (P - 1))) - 1ull)) | (proto & lo_mask);
  }
  *w = q;                                                        // positions 0 .. q-1 pair with g[j+1]; position q is free
  const uint64_t hi_mask = ((1ull << (2 * P)) - 1ull) & ~((1ull << (2 * (P - q))) - 1ull);
  return ((proto << 2) & hi_mask) | (proto & lo_mask);
}

__global__ void __launch_bounds__(kScanThreads, FF_SCAN_MIN_BLOCKS) k_pattern_scan(GeneralParams gp) {
  __shared__ uint64_t s_hits[kScanWarps * kHW];
  __shared__ unsigned int s_hitn[kScanWarps];
  const ScanParams &p = gp.sp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpHits wh{s_hits + warp * kHW, &s_hitn[warp], p.hits, p.hit_count, p.hit_cap, 0};
  if (lane == 0) s_hitn[warp] = 0;
  __syncwarp();
  unsigned long long compares = 0;
  const int ipg = gp.pl.items_per_guide;
  const long long n_items = gp.n_active * (long long)ipg;
  const long long n_warps = (long long)gridDim.x * kScanWarps;
  const int a = gp.a_bases, P = gp.P;
  for (long long item = (long long)blockIdx.x * kScanWarps + warp; item < n_items; item += n_warps) {
    const long long gi = item / ipg;
    const int rem = (int)(item - gi * ipg);
    const long long g = gp.active ? (long long)gp.active[gi] : gi;
    int pt = 0;
    while (rem >= gp.pl.item0[pt + 1]) ++pt;  // <= 37 patterns, warp-uniform
    const int bi = rem - gp.pl.item0[pt];
    const PatternPlan::Cls c = gp.pl.c[gp.pl.cls[pt]];
    const uint64_t guide = p.guides[g];
    int w;
    const uint64_t T = pattern_template((guide >> p.proto_shift) & p.proto_mask, gp.pl.type[pt], gp.pl.q[pt], P, &w);
    const uint32_t key_a = (uint32_t)(T >> p.b_bits);
    const uint32_t key_b = (uint32_t)(T & ((1ull << p.b_bits) - 1ull));
    const uint64_t guide_key = (uint64_t)g << p.tbits;
    const bool wild_a = w >= 0 && w < a, wild_b = w >= a;
    PatItem pi;
    pi.c0 = gp.c0; pi.c1 = gp.c1; pi.cell_shift = gp.cell_shift; pi.cell_off = gp.cell_off;
    if (bi < c.itemsA) {
      const int seed0 = bi * 32;
      pi.masks = wild_a ? gp.a_masks_w1 : p.A.masks;
      pi.wild_bit = wild_a ? 2 * (a - 1 - w) : -1;
      pi.pmask = wild_b ? ~(3u << (2 * (P - 1 - w))) : 0xFFFFFFFFu;
      pi.lo_d = -1;
      scan_seeds<false, true>(p, p.A, wh, lane, key_a, key_b, seed0, min(32, c.nA - seed0), guide_key, compares, pi);
    } else {
      const int seed0 = (bi - c.itemsA) * c.spiB;
      pi.masks = wild_b ? gp.b_masks_w1 : p.B.masks;
      pi.wild_bit = wild_b ? 2 * (P - 1 - w) : -1;
      pi.pmask = wild_a ? ~(3u << (2 * (a - 1 - w))) : 0xFFFFFFFFu;
      pi.lo_d = c.hA;
      scan_seeds<true, true>(p, p.B, wh, lane, key_b, key_a, seed0, min(c.spiB, c.nB - seed0), guide_key, compares, pi);
    }
  }
  flush_warp_hits(wh, lane);
  for (int o = 16; o > 0; o >>= 1) compares += __shfl_down_sync(0xffffffffu, compares, o);
  if (lane == 0 && compares) atomicAdd(p.n_compares, compares);
}

// ---- per-window bookkeeping
// segment of every active guide inside the window's sorted unique keys
__global__ void k_segments_active(const uint64_t *__restrict__ keys, int64_t n_keys, const uint32_t *__restrict__ active, int64_t n_active,
                                  int tbits, int64_t *__restrict__ seg_start, int64_t *__restrict__ seg_end) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= 2 * n_active) return;
  const int64_t i = t >> 1;
  const uint64_t g = active ? (uint64_t)active[i] : (uint64_t)i;
  const uint64_t want = (g + (uint64_t)(t & 1)) << tbits;
  int64_t lo = 0, hi = n_keys;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < want) lo = mid + 1; else hi = mid;
  }
  if (t & 1) seg_end[i] = lo; else seg_start[i] = lo;
}

// k_overflow_cut continued across windows: one warp per active guide appends this window's hits (database order) while
// its running total is below max_ot (ResultsAggregator.scala:61-69 / CRISPRSiteOT.scala:39-46).
__global__ void k_cut_window(const uint64_t *__restrict__ keys, const int64_t *__restrict__ seg_start, const int64_t *__restrict__ seg_end,
                             const uint32_t *__restrict__ active, int64_t n_active, const uint64_t *__restrict__ targets, int max_ot,
                             int tbits, long long *__restrict__ running_all, int64_t *__restrict__ n_keep) {
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n_active) return;
  const int64_t g = active ? (int64_t)active[i] : i;
  const int64_t s0 = seg_start[i], s1 = seg_end[i];
  long long running = running_all[g];
  int64_t kept = 0;
  for (int64_t base = s0; base < s1 && running < max_ot; base += 32) {
    const int64_t j = base + lane;
    long long c = 0;
    if (j < s1) c = (long long)(targets[keys[j] & ((1ull << tbits) - 1ull)] >> 48);
    long long incl = c;
    for (int o = 1; o < 32; o <<= 1) {
      const long long v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const bool keep = (j < s1) && (running + incl - c < max_ot);
    const int nk = __popc(__ballot_sync(0xffffffffu, keep));  // kept hits form a prefix of the chunk
    kept += nk;
    const long long chunk_total = __shfl_sync(0xffffffffu, incl, nk > 0 ? nk - 1 : 0);
    if (nk > 0) running += chunk_total;
    if (nk < 32) break;
  }
  if (lane == 0) {
    n_keep[i] = kept;
    running_all[g] = running;
  }
}

__global__ void k_copy_kept(const uint64_t *__restrict__ keys, const int64_t *__restrict__ seg_start, const int64_t *__restrict__ keep_off,
                            int64_t n_active, uint64_t *__restrict__ kept) {
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n_active) return;
  const int64_t s0 = seg_start[i], o0 = keep_off[i], n = keep_off[i + 1] - o0;
  for (int64_t j = lane; j < n; j += 32) kept[o0 + j] = keys[s0 + j];
}

__global__ void k_still_collecting(const uint32_t *__restrict__ active, int64_t n_active, const long long *__restrict__ running, int max_ot,
                                   uint32_t *__restrict__ ids, uint8_t *__restrict__ flags, unsigned long long *__restrict__ collected) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  long long mine = 0;
  if (i < n_active) {
    const uint32_t g = active ? active[i] : (uint32_t)i;
    ids[i] = g;
    const bool open = running[g] < max_ot;
    flags[i] = open ? 1 : 0;
    if (open) mine = running[g];
  }
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(collected, (unsigned long long)mine);  // what the open guides hold so far
}

__global__ void k_finish_totals(const long long *__restrict__ running, int64_t n_guides, int max_ot, int32_t *__restrict__ total_count,
                                uint8_t *__restrict__ overflowed) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= n_guides) return;
  total_count[g] = (int32_t)running[g];
  overflowed[g] = running[g] >= max_ot ? 1 : 0;
}

// Best alignment of a (guide, target) pair under the bulge flags: smallest (mismatches, type, position).
__device__ __forceinline__ void best_alignment(uint64_t gp, uint64_t tp, int P, int flags, int *mm_out, int *code_out) {
  const uint64_t full = (1ull << (2 * P)) - 1ull, odd = 0x5555555555555555ull & full;
  const uint64_t x = gp ^ tp;
  const uint64_t u = (x | (x >> 1)) & odd;  // per-position mismatch of g[j] vs t[j]
  int best = __popcll(u), code = 0;
  if (flags & FF_BULGE_RNA) {
    const uint64_t y = (gp >> 2) ^ tp;
    const uint64_t s = (y | (y >> 1)) & odd;  // g[j-1] vs t[j]
    for (int q = 1; q <= P - 2; ++q) {
      const uint64_t lo_mask = (1ull << (2 * (P - 1 - q))) - 1ull;
      const uint64_t hi_mask = ~lo_mask & ((1ull << (2 * (P - 1))) - 1ull);  // positions 1 .. q
      const int mm = __popcll(s & hi_mask) + __popcll(u & lo_mask);
      if (mm < best) { best = mm; code = 0x40 | q; }
    }
  }
  if (flags & FF_BULGE_DNA) {
    const uint64_t y = ((gp << 2) & full) ^ tp;
    const uint64_t s = (y | (y >> 1)) & odd;  // g[j+1] vs t[j]
    for (int q = 1; q <= P - 2; ++q) {
      const uint64_t lo_mask = (1ull << (2 * (P - 1 - q))) - 1ull;
      const uint64_t hi_mask = full & ~((1ull << (2 * (P - q))) - 1ull);     // positions 0 .. q-1
      const int mm = __popcll(s & hi_mask) + __popcll(u & lo_mask);
      if (mm < best) { best = mm; code = 0x80 | q; }
    }
  }
  *mm_out = best; *code_out = code;
}

// One warp per guide: copy the kept hits out (target long, mismatch count, bulge code, target index).
__global__ void k_gather_general(const uint64_t *__restrict__ keys, const int64_t *__restrict__ row_ptr, const uint64_t *__restrict__ targets,
                                 const uint64_t *__restrict__ guides, int proto_shift, int P, int bulge_flags, int64_t n_guides, int tbits,
                                 uint64_t *__restrict__ out_targets, uint8_t *__restrict__ out_mm, uint8_t *__restrict__ out_bulge,
                                 uint32_t *__restrict__ out_tidx) {
  const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n_guides) return;
  const int64_t r0 = row_ptr[g], r1 = row_ptr[g + 1];
  const uint64_t pm = (1ull << (2 * P)) - 1ull;
  const uint64_t gproto = (guides[g] >> proto_shift) & pm;
  for (int64_t i = r0 + lane; i < r1; i += 32) {
    const uint32_t t = (uint32_t)(keys[i] & ((1ull << tbits) - 1ull));
    const uint64_t tl = targets[t];
    int mm, code;
    best_alignment(gproto, (tl >> proto_shift) & pm, P, bulge_flags, &mm, &code);
    out_targets[i] = tl;
    out_mm[i] = (uint8_t)mm;
    out_bulge[i] = (uint8_t)code;
    out_tidx[i] = t;
  }
}

// grow a device buffer and keep its first `used` bytes
static int grow_keep(DevBuf &b, size_t used, size_t want, cudaStream_t st) {
  if (want <= b.cap && b.p) return FF_OK;
  DevBuf bigger;
  FF_TRY(bigger.reserve(want + want / 2));
  if (b.p && used) {
    FF_CUDA(cudaMemcpyAsync(bigger.p, b.p, used, cudaMemcpyDeviceToDevice, st));
    FF_CUDA(cudaStreamSynchronize(st));
  }
  b.release();
  b = bigger;
  return FF_OK;
}

static double ball_probability(int bases, int k) {  // P(a random `bases`-mer is within k mismatches)
  double prob = 0.0, term = 1.0;
  for (int i = 0; i <= std::min(k, bases); ++i) {
    prob += term;
    term = term * 3.0 * (double)(bases - i) / (double)(i + 1);
  }
  return prob / std::pow(4.0, (double)bases);
}

static int discover_general(ff_ctx *ctx, const uint64_t *d_guides, int64_t n_guides, int max_mm, int max_ot, bool want_positions,
                            int bulge_flags, int slot, DeviceResult *res) {
  Database &db = ctx->db;
  ff_ctx::OutSlot &os = ctx->out[slot & 1];
  cudaStream_t st = ctx->stream;
  const int64_t G = n_guides, Gp = G > 0 ? G : 1;
  const int P = db.proto_bases, a = db.A.key_bases, b = db.B.key_bases;
  if (bulge_flags && !(db.pack.scan_len == 23 && !db.pack.five_prime && P == 20)) {
    set_error("the bulge mode is defined for 23-bp Cas9 parameter packs (20-base protospacer, 3' PAM) only");
    return FF_EUNSUPPORTED;
  }
  const bool can_window = !db.pack.five_prime && !db.A.d_canon && a >= 3;
  int launches = 0;
  ff_timings tm = {};
  const int k_eff = std::min(max_mm, P);

  // ---- patterns and their seed plans
  GeneralParams gp;
  memset(&gp, 0, sizeof gp);
  PatternPlan &pl = gp.pl;
  {
    const double n = (double)db.n_targets;
    const double bucket_a = n / (double)(1ull << (2 * a)), bucket_b = n / (double)(1ull << (2 * b));
    const double per_seed = 24.0;
    for (int cls = 0; cls < 3; ++cls) {
      double best = -1.0;
      PatternPlan::Cls &c = pl.c[cls];
      for (int h = 0; h <= std::min(k_eff, a - 1); ++h) {
        const int hb = k_eff - h - 1;
        const double sa = cls == 1 ? 4.0 * db.A.cum_w1[std::min(h, a - 1)] : (double)db.A.cum[std::min(h, a)];
        const double sb = hb < 0 ? 0.0 : cls == 2 ? 4.0 * db.B.cum_w1[std::min(hb, b - 1)] : (double)db.B.cum[std::min(hb, b)];
        const double cost = sa * (bucket_a + per_seed) + sb * (bucket_b + per_seed);
        if (best < 0 || cost < best) { best = cost; c.hA = h; c.nA = (int)sa; c.nB = (int)sb; }
      }
      c.itemsA = (c.nA + 31) / 32;
    }
    int np = 0;
    auto add = [&](int type, int q, int cls) { pl.type[np] = (uint8_t)type; pl.q[np] = (uint8_t)q; pl.cls[np] = (uint8_t)cls; np++; };
    add(0, 0, 0);
    if (bulge_flags & FF_BULGE_RNA) for (int q = 1; q <= P - 2; ++q) add(1, q, 1);
    if (bulge_flags & FF_BULGE_DNA) for (int q = 1; q <= P - 2; ++q) add(2, q, q < a ? 1 : 2);
    pl.n_patterns = np;
  }
  // part-two seeds per work item follow the length of the bucket run inside a window
  auto plan_items = [&](int cells) {
    const double run = (double)db.n_targets / (double)(1ull << (2 * b)) * (double)cells / (double)kCells;
    int spi = run > 2048 ? 1 : run > 512 ? 4 : run > 128 ? 8 : 32;
    if (ctx->opt.b_spi > 0) spi = ctx->opt.b_spi;
    for (int cls = 0; cls < 3; ++cls) { pl.c[cls].spiB = spi; pl.c[cls].itemsB = (pl.c[cls].nB + spi - 1) / spi; }
    pl.item0[0] = 0;
    for (int i = 0; i < pl.n_patterns; ++i) pl.item0[i + 1] = pl.item0[i] + pl.c[pl.cls[i]].itemsA + pl.c[pl.cls[i]].itemsB;
    pl.items_per_guide = pl.item0[pl.n_patterns];
  };
  ScanParams &sp = gp.sp;
  sp.guides = d_guides; sp.n_guides = G;
  sp.A.off = db.A.d_off; sp.A.other = db.A.d_other; sp.A.canon = db.A.d_canon; sp.A.masks = db.A.d_masks;
  sp.B.off = db.B.d_off; sp.B.other = db.B.d_other; sp.B.canon = db.B.d_canon; sp.B.masks = db.B.d_masks;
  sp.proto_shift = db.proto_shift; sp.b_bits = 2 * b; sp.proto_mask = (1ull << (2 * P)) - 1ull;
  sp.k = k_eff; sp.hA = 0;
  int tbits = 1;
  while ((1ull << tbits) < db.n_targets + 1) tbits++;
  sp.tbits = tbits;
  int gbits = 1;
  while ((1ll << gbits) < Gp) gbits++;
  gp.a_masks_w1 = db.A.d_masks_w1; gp.b_masks_w1 = db.B.d_masks_w1;
  gp.a_bases = a; gp.P = P;

  // ---- expected hits per guide decide the window width (in cells of kCells)
  const int n_bulge_patterns = pl.n_patterns - 1;
  const double prob = ball_probability(P, k_eff) + n_bulge_patterns * ball_probability(P - 1, std::min(k_eff, P - 1));
  const double exp_hits = (double)db.n_targets * prob;  // candidates per guide (duplicates between patterns included)
  // Templates of neighbouring bulge positions differ in one base, so their hit sets overlap: ~0.45 of the candidates are
  // distinct targets (measured on the bench index).  The first window aims at 1.25 x maximumOffTargets occurrences for
  // a typical guide; later windows are sized from what the remaining guides have collected so far.
  const double exp_occ = exp_hits * (n_bulge_patterns ? 0.45 : 1.0) * 1.3;
  int cells_per_window = kCells;
  if (can_window && max_ot > 0 && exp_occ > 0.75 * (double)max_ot)
    cells_per_window = std::max(1, std::min(kCells, (int)std::ceil((double)kCells * 1.25 * (double)max_ot / exp_occ)));
  bool fixed_window = false;
  { const int v = ctx->opt.window_cells; if (v >= 1 && can_window) { cells_per_window = std::min(v, kCells); fixed_window = true; } }
  const bool windowed = cells_per_window < kCells;
  if (windowed) FF_TRY(db_build_cell_offsets(ctx));
  gp.cell_off = windowed ? db.d_cell_off : nullptr;
  gp.cell_shift = 2 * a - 6;  // kCells = 4^3

  // ---- workspaces
  FF_TRY(ctx->counters.reserve(64));
  FF_TRY(ctx->seg_start.reserve((Gp + 1) * 8));
  FF_TRY(ctx->seg_end.reserve((Gp + 1) * 8));
  FF_TRY(ctx->n_keep.reserve((Gp + 1) * 8));
  FF_TRY(ctx->running.reserve(Gp * 8));
  FF_TRY(ctx->active.reserve(Gp * 4));
  FF_TRY(ctx->active2.reserve(Gp * 4));
  FF_TRY(ctx->act_flags.reserve(Gp));
  FF_TRY(ctx->n_sel.reserve(16));
  FF_TRY(os.row_ptr.reserve((Gp + 1) * 8));
  FF_TRY(os.total_count.reserve(Gp * 4));
  FF_TRY(os.overflowed.reserve(Gp));
  FF_CUDA(cudaEventRecord(ctx->ev[0], st));
  FF_CUDA(cudaMemsetAsync(ctx->running.p, 0, Gp * 8, st));
  unsigned long long *d_cnt = ctx->counters.as<unsigned long long>();
  sp.hit_count = d_cnt; sp.n_compares = d_cnt + 1;
  if (ctx->hit_cap == 0) ctx->hit_cap = 1u << 22;

  const uint32_t *d_active = nullptr;  // nullptr = every guide
  int64_t n_active = G, n_kept = 0;
  uint64_t n_cand_total = 0, n_compares = 0, scan_bytes = 0;
  int windows = 0, scan_launches = 0;
  float scan_ms = 0.f;
  size_t tmp_bytes = 0;
  const int max_grid = ctx->sm_count * 8;
  for (int c0 = 0, c1 = 0; c0 < kCells && n_active > 0; c0 = c1) {
    c1 = std::min(kCells, c0 + cells_per_window);
    plan_items(c1 - c0);
    gp.c0 = c0; gp.c1 = c1; gp.active = d_active; gp.n_active = n_active;
    {
      const double want = (double)n_active * (exp_hits * 1.4 * (double)(c1 - c0) / (double)kCells + 64.0);
      const size_t cap = (size_t)std::min(want, 268435456.0);
      if (cap > ctx->hit_cap) ctx->hit_cap = cap;
    }
    const long long n_items = n_active * (long long)pl.items_per_guide;
    const int grid = (int)std::max<long long>(1, std::min<long long>(max_grid, (n_items + kScanWarps - 1) / kScanWarps));
    unsigned long long h_cnt[2] = {0, 0};
    for (;;) {
      FF_TRY(ctx->hit_keys.reserve(ctx->hit_cap * 8));
      FF_TRY(ctx->hit_keys_sorted.reserve(ctx->hit_cap * 8));
      sp.hits = ctx->hit_keys.as<uint64_t>(); sp.hit_cap = ctx->hit_cap;
      FF_CUDA(cudaMemsetAsync(d_cnt, 0, 16, st));
      FF_CUDA(cudaEventRecord(ctx->ev[1], st));
      k_pattern_scan<<<grid, kScanThreads, 0, st>>>(gp);
      launches++; scan_launches++;
      FF_CUDA(cudaEventRecord(ctx->ev[2], st));
      FF_TRY(fetch_words(ctx, st, d_cnt, &h_cnt[0], d_cnt + 1, &h_cnt[1]));  // (mapped status words: no copy-engine queue)
      if (h_cnt[0] <= ctx->hit_cap) break;
      ctx->hit_cap = (size_t)(h_cnt[0] + h_cnt[0] / 8 + 1024);
    }
    {
      float ms = 0.f;
      FF_CUDA(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]));
      scan_ms += ms;
    }
    windows++;
    const int64_t n_cand = (int64_t)h_cnt[0];
    n_cand_total += (uint64_t)n_cand; n_compares += h_cnt[1];
    {
      uint64_t seeds = 0;
      for (int i = 0; i < pl.n_patterns; ++i) seeds += (uint64_t)(pl.c[pl.cls[i]].nA + pl.c[pl.cls[i]].nB);
      scan_bytes += (uint64_t)n_active * seeds * 8ull + h_cnt[1] * 4ull + (uint64_t)n_active * 8ull + (uint64_t)n_cand * 8ull;
    }
    // order the window's candidates by (guide, database index) and drop the duplicates different templates produced
    const uint64_t *keys = ctx->hit_keys.as<uint64_t>();
    int64_t n_uniq = 0;
    if (n_cand > 0) {
      FF_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, ctx->hit_keys.as<uint64_t>(), ctx->hit_keys_sorted.as<uint64_t>(), n_cand, 0, tbits + gbits, st));
      FF_TRY(ctx->cub_tmp.reserve(tmp_bytes));
      FF_CUDA(cub::DeviceRadixSort::SortKeys(ctx->cub_tmp.p, tmp_bytes, ctx->hit_keys.as<uint64_t>(), ctx->hit_keys_sorted.as<uint64_t>(), n_cand, 0, tbits + gbits, st));
      launches += 2 + (tbits + gbits + 7) / 8;
      if (pl.n_patterns > 1) {
        FF_CUDA(cub::DeviceSelect::Unique(nullptr, tmp_bytes, ctx->hit_keys_sorted.as<uint64_t>(), ctx->hit_keys.as<uint64_t>(), ctx->n_sel.as<int64_t>(), n_cand, st));
        FF_TRY(ctx->cub_tmp.reserve(tmp_bytes));
        FF_CUDA(cub::DeviceSelect::Unique(ctx->cub_tmp.p, tmp_bytes, ctx->hit_keys_sorted.as<uint64_t>(), ctx->hit_keys.as<uint64_t>(), ctx->n_sel.as<int64_t>(), n_cand, st));
        launches += 2;
        { unsigned long long v = 0; FF_TRY(fetch_words(ctx, st, ctx->n_sel.p, &v)); n_uniq = (int64_t)v; }
        keys = ctx->hit_keys.as<uint64_t>();
      } else {
        n_uniq = n_cand;
        keys = ctx->hit_keys_sorted.as<uint64_t>();
      }
    }
    // overflow cut of this window, continued from the running totals
    int64_t n_keep_total = 0;
    if (n_active > 0) {
      k_segments_active<<<blocks_for(2 * n_active, 256), 256, 0, st>>>(keys, n_uniq, d_active, n_active, tbits, ctx->seg_start.as<int64_t>(), ctx->seg_end.as<int64_t>());
      k_cut_window<<<blocks_for(n_active * 32, 256), 256, 0, st>>>(keys, ctx->seg_start.as<int64_t>(), ctx->seg_end.as<int64_t>(), d_active, n_active, db.d_targets,
                                                                  max_ot, tbits, ctx->running.as<long long>(), ctx->n_keep.as<int64_t>());
      launches += 2;
      FF_CUDA(cudaMemsetAsync(ctx->n_keep.as<int64_t>() + n_active, 0, 8, st));
      // keep_off lives in seg_end (no longer needed after the cut)
      FF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ctx->n_keep.as<int64_t>(), ctx->seg_end.as<int64_t>(), n_active + 1, st));
      FF_TRY(ctx->cub_tmp.reserve(tmp_bytes));
      FF_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_bytes, ctx->n_keep.as<int64_t>(), ctx->seg_end.as<int64_t>(), n_active + 1, st));
      launches += 2;
      { unsigned long long v = 0; FF_TRY(fetch_words(ctx, st, ctx->seg_end.as<int64_t>() + n_active, &v)); n_keep_total = (int64_t)v; }
      if (n_keep_total > 0) {
        FF_TRY(grow_keep(ctx->kept_keys, (size_t)n_kept * 8, (size_t)(n_kept + n_keep_total) * 8, st));
        k_copy_kept<<<blocks_for(n_active * 32, 256), 256, 0, st>>>(keys, ctx->seg_start.as<int64_t>(), ctx->seg_end.as<int64_t>(), n_active,
                                                                   ctx->kept_keys.as<uint64_t>() + n_kept);
        launches++;
        n_kept += n_keep_total;
      }
    }
    // guides that are full leave the traversal
    if (c1 < kCells) {
      uint32_t *ids = ctx->active2.as<uint32_t>();
      uint32_t *next = ctx->active.as<uint32_t>();
      FF_CUDA(cudaMemsetAsync(d_cnt + 2, 0, 8, st));
      k_still_collecting<<<blocks_for(n_active, 256), 256, 0, st>>>(d_active, n_active, ctx->running.as<long long>(), max_ot, ids, ctx->act_flags.as<uint8_t>(), d_cnt + 2);
      FF_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, ids, ctx->act_flags.as<uint8_t>(), next, ctx->n_sel.as<int64_t>(), n_active, st));
      FF_TRY(ctx->cub_tmp.reserve(tmp_bytes));
      FF_CUDA(cub::DeviceSelect::Flagged(ctx->cub_tmp.p, tmp_bytes, ids, ctx->act_flags.as<uint8_t>(), next, ctx->n_sel.as<int64_t>(), n_active, st));
      launches += 3;
      int64_t n_next = 0;
      unsigned long long collected = 0;
      { unsigned long long v = 0; FF_TRY(fetch_words(ctx, st, ctx->n_sel.p, &v, d_cnt + 2, &collected)); n_next = (int64_t)v; }
      d_active = next;
      n_active = n_next;
      if (!fixed_window && n_active > 0) {
        // the open guides collected `mean` occurrences over c1 cells: size the next window for what they still need
        const double mean = (double)collected / (double)n_active;
        const double rate = mean / (double)c1;
        const double need = ((double)max_ot - mean) * 1.25;
        cells_per_window = rate > 0 ? (int)std::ceil(need / rate) : kCells;
        cells_per_window = std::max(1, std::min(kCells, cells_per_window));
      }
    }
  }
  FF_CUDA(cudaEventRecord(ctx->ev[2], st));

  // ---- rows: kept keys in (guide, database index) order
  const uint64_t *kept = ctx->kept_keys.as<uint64_t>();
  if (windows > 1 && n_kept > 0) {
    FF_TRY(ctx->kept_sorted.reserve((size_t)n_kept * 8));
    FF_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, ctx->kept_keys.as<uint64_t>(), ctx->kept_sorted.as<uint64_t>(), n_kept, 0, tbits + gbits, st));
    FF_TRY(ctx->cub_tmp.reserve(tmp_bytes));
    FF_CUDA(cub::DeviceRadixSort::SortKeys(ctx->cub_tmp.p, tmp_bytes, ctx->kept_keys.as<uint64_t>(), ctx->kept_sorted.as<uint64_t>(), n_kept, 0, tbits + gbits, st));
    launches += 2 + (tbits + gbits + 7) / 8;
    kept = ctx->kept_sorted.as<uint64_t>();
  }
  FF_CUDA(cudaEventRecord(ctx->ev[3], st));
  k_segments<<<blocks_for(G + 1, 256), 256, 0, st>>>(kept, n_kept, G, tbits, os.row_ptr.as<int64_t>());
  launches++;
  if (G > 0) {
    k_finish_totals<<<blocks_for(G, 256), 256, 0, st>>>(ctx->running.as<long long>(), G, max_ot, os.total_count.as<int32_t>(), os.overflowed.as<uint8_t>());
    launches++;
  }
  const int64_t n_hits = n_kept;
  const int64_t Hp = n_hits > 0 ? n_hits : 1;
  FF_TRY(os.out_targets.reserve(Hp * 8));
  FF_TRY(os.out_mm.reserve(Hp));
  FF_TRY(os.out_bulge.reserve(Hp));
  FF_TRY(os.out_tidx.reserve(Hp * 4));
  if (G > 0 && n_hits > 0) {
    k_gather_general<<<blocks_for(G * 32, 256), 256, 0, st>>>(kept, os.row_ptr.as<int64_t>(), db.d_targets, d_guides, db.proto_shift, P, bulge_flags, G, tbits,
                                                             os.out_targets.as<uint64_t>(), os.out_mm.as<uint8_t>(), os.out_bulge.as<uint8_t>(),
                                                             os.out_tidx.as<uint32_t>());
    launches++;
  }
  int64_t n_pos = 0;
  FF_TRY(gather_positions(ctx, os, want_positions, n_hits, res, &n_pos, &launches));
  FF_CUDA(cudaEventRecord(ctx->ev[4], st));
  FF_CUDA(cudaStreamSynchronize(st));
  FF_CUDA(cudaGetLastError());

  tm.prep_ms = 0.f;
  tm.scan_ms = scan_ms;
  FF_CUDA(cudaEventElapsedTime(&tm.order_ms, ctx->ev[2], ctx->ev[3]));
  FF_CUDA(cudaEventElapsedTime(&tm.cut_ms, ctx->ev[3], ctx->ev[4]));
  FF_CUDA(cudaEventElapsedTime(&tm.total_ms, ctx->ev[0], ctx->ev[4]));
  tm.score_ms = 0.f;
  tm.scan_launches = scan_launches;
  tm.kernel_launches = launches;
  tm.scan_bytes_read = scan_bytes;
  ctx->last = tm;

  res->n_guides = G; res->n_hits = n_hits; res->n_positions = n_pos;
  res->n_candidate_hits = n_cand_total; res->n_compares = n_compares;
  res->d_row_ptr = os.row_ptr.as<int64_t>(); res->d_targets = os.out_targets.as<uint64_t>();
  res->d_mismatches = os.out_mm.as<uint8_t>(); res->d_bulge = os.out_bulge.as<uint8_t>();
  res->d_total_count = os.total_count.as<int32_t>(); res->d_overflowed = os.overflowed.as<uint8_t>();
  res->d_tidx = os.out_tidx.as<uint32_t>();
  return FF_OK;
}

int discover_on_device(ff_ctx *ctx, const uint64_t *d_guides, int64_t n_guides, int max_mm, int max_ot, bool want_positions,
                       int bulge_flags, int slot, DeviceResult *res) {
  Database &db = ctx->db;
  if (!db.resident) { set_error("no database resident in this context"); return FF_ENODB; }
  if (n_guides < 0 || max_mm < 0 || max_ot < 0 || (n_guides > 0 && !d_guides)) { set_error("bad discover argument"); return FF_EINVAL; }
  if (n_guides >= (1ll << 31)) { set_error("too many guides in one call"); return FF_EINVAL; }
  if (bulge_flags & ~(FF_BULGE_RNA | FF_BULGE_DNA)) { set_error("unknown bulge flag"); return FF_EINVAL; }
  bool general = bulge_flags != 0;
  if (!general && !db.pack.five_prime && !db.A.d_canon && max_ot > 0) {
    // mismatch-only search: windows pay off once a typical guide overflows (e.g. k = 6 on a human-sized index)
    const double exp_occ = (double)db.n_targets * ball_probability(db.proto_bases, std::min(max_mm, db.proto_bases)) * 1.3;
    general = exp_occ > 2.0 * (double)max_ot;
  }
  general = general || ctx->opt.force_general != 0;
  if (general) return discover_general(ctx, d_guides, n_guides, max_mm, max_ot, want_positions, bulge_flags, slot, res);
  return discover_plain(ctx, d_guides, n_guides, max_mm, max_ot, want_positions, slot, res);
}
