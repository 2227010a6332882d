// Cell-major seed scan (included by ff_discover.cu, inside namespace ff).
//
// The guide-major kernel (k_seed_scan) reads every bucket from HBM each time a guide's seed lands in it; with 100 000
// guides against a human-sized index a part-one bucket is read ~12 times and a part-two bucket ~10 times per call --
// 28 GB of traffic over a 2.4 GB index.  Here the SAME (guide, seed) pairs are visited in the order of the bucket they
// land in, coarsely: the key space of each index half is cut into 64 cells (the value of the first three key bases,
// 19 MB of index per cell on a human-sized database), and the whole grid works through one cell at a time, so after
// the first touch a bucket is served from the 126 MB L2.  No sort of the 5.6e7 pairs is needed:
//   * a seed mask whose first three bases are m3 moves a guide whose key starts with t3 into cell t3 ^ m3, so with the
//     masks grouped by m3 (SeedIndex::d_gmasks) the seeds of (guide class t3, cell c) are one contiguous run of group
//     c ^ t3, the same run for every guide of the class;
//   * guides are listed by class (a 6-bit radix sort of the guide indices), and the pairs of a (cell, class) segment
//     are numbered guide-major, 32 per warp (part one) or a few per warp (part two, whose buckets are long).
// A lane looks up its own pair's bucket; the warp then streams the buckets exactly like k_seed_scan does, except that
// the guide's probe and index travel with the bucket (shuffles) because neighbouring lanes may belong to different
// guides.  Results are the same hit keys; everything after the scan is unchanged.

constexpr int kSegs = 2 * kCells * kCells;  // (phase, cell, class)

struct CellParams {
  ScanParams sp;
  const uint32_t *gmasks_a, *gmasks_b;  // grouped masks
  int goff_a[kCells], goff_b[kCells];   // first mask of every group
  int ng_a[kCells], ng_b[kCells];       // seeds of every group within this call's budgets (hA / k - hA - 1)
  int ppi_b;                            // pairs per work item in the part-two phase
  const uint32_t *perm_a, *perm_b;      // guide indices listed by class
  const int *cls_off;                   // [2][kCells + 1] first guide of every class
  const long long *seg_item0;           // [kSegs + 1] first work item of every segment; [kSegs] = number of items
  unsigned long long *next_item;        // [2] work counters (one per index half): warps claim items in order, so the whole
                                        // grid stays inside ~one cell
};

__global__ void k_guide_classes(const uint64_t *__restrict__ guides, int64_t n, int proto_shift, uint64_t proto_mask, int b_bits, int a_bits,
                                uint32_t *__restrict__ cls_a, uint32_t *__restrict__ cls_b, uint32_t *__restrict__ iota, int *__restrict__ hist) {
  __shared__ int s_hist[2 * kCells];
  if (threadIdx.x < 2 * kCells) s_hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g < n) {
    const uint64_t proto = (guides[g] >> proto_shift) & proto_mask;
    const uint32_t ca = (uint32_t)(proto >> b_bits) >> (a_bits - 6);
    const uint32_t cb = (uint32_t)(proto & ((1ull << b_bits) - 1ull)) >> (b_bits - 6);
    cls_a[g] = ca; cls_b[g] = cb; iota[g] = (uint32_t)g;
    atomicAdd(s_hist + ca, 1);
    atomicAdd(s_hist + kCells + cb, 1);
  }
  __syncthreads();
  if (threadIdx.x < 2 * kCells && s_hist[threadIdx.x]) atomicAdd(hist + threadIdx.x, s_hist[threadIdx.x]);
}

// one block: class offsets and the work-item prefix over the (phase, cell, class) segments
__global__ void __launch_bounds__(1024) k_build_segments(const int *__restrict__ hist, CellParams cp, int *__restrict__ cls_off,
                                                         long long *__restrict__ seg_item0) {
  __shared__ long long s_part[1024];
  __shared__ int s_size[2 * kCells];
  const int t = threadIdx.x;
  if (t < 2 * kCells) s_size[t] = hist[t];
  __syncthreads();
  if (t < 2) {
    int acc = 0;
    for (int c = 0; c < kCells; ++c) { cls_off[t * (kCells + 1) + c] = acc; acc += s_size[t * kCells + c]; }
    cls_off[t * (kCells + 1) + kCells] = acc;
  }
  constexpr int per = kSegs / 1024;
  long long mine[per], sum = 0;
  for (int i = 0; i < per; ++i) {
    const int seg = t * per + i;
    const int phase = seg / (kCells * kCells), cell = (seg / kCells) % kCells, cls = seg % kCells;
    const int n = phase ? cp.ng_b[cell ^ cls] : cp.ng_a[cell ^ cls];
    const long long pairs = (long long)s_size[phase * kCells + cls] * n;
    const int ppi = phase ? cp.ppi_b : 32;
    mine[i] = (pairs + ppi - 1) / ppi;
    sum += mine[i];
  }
  s_part[t] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {  // inclusive scan of the per-thread sums
    const long long v = t >= o ? s_part[t - o] : 0;
    __syncthreads();
    s_part[t] += v;
    __syncthreads();
  }
  long long acc = s_part[t] - sum;
  for (int i = 0; i < per; ++i) { seg_item0[t * per + i] = acc; acc += mine[i]; }
  if (t == 1023) seg_item0[kSegs] = acc;
}

// Stream the (<= 32) buckets held by the lanes: the grouped first-chunk loads and tail loop of scan_seeds, with the
// probe and the guide index taken from the bucket's lane.
template <bool PASS_B>
__device__ __forceinline__ void stream_pairs(const ScanParams &p, const SeedSide &sd, const WarpHits &wh, int lane, uint32_t lo, uint32_t hi,
                                             int budget, uint32_t probe, uint32_t gid, int n) {
  const uint32_t lane4 = 4u * lane;
  const bool any_long = __any_sync(0xffffffffu, hi - (lo & ~3u) > 128u);
  for (int l0 = 0; l0 < n; l0 += FF_GROUP) {
    uint32_t blo[FF_GROUP], bhi[FF_GROUP];
    uint4 v[FF_GROUP];
#pragma unroll
    for (int j = 0; j < FF_GROUP; ++j) {
      blo[j] = __shfl_sync(0xffffffffu, lo, l0 + j);
      bhi[j] = __shfl_sync(0xffffffffu, hi, l0 + j);
      v[j] = make_uint4(0, 0, 0, 0);
      if ((blo[j] & ~3u) + lane4 < bhi[j]) v[j] = ldg128(sd.other + (blo[j] & ~3u) + lane4);
    }
#pragma unroll
    for (int j = 0; j < FF_GROUP; ++j) {
      const int bud = __shfl_sync(0xffffffffu, budget, l0 + j);
      const uint32_t pr = __shfl_sync(0xffffffffu, probe, l0 + j);
      const uint32_t gj = __shfl_sync(0xffffffffu, gid, l0 + j);
      const uint32_t base = (blo[j] & ~3u) + lane4;
      if (base < bhi[j]) verify_chunk<PASS_B, false>(p, wh, sd.canon, v[j], base, blo[j], bhi[j], pr, bud, (uint64_t)gj << p.tbits, 0u, 0);
    }
    if (!any_long) continue;
#pragma unroll 1
    for (int j = 0; j < FF_GROUP; ++j) {
      const uint32_t jlo = __shfl_sync(0xffffffffu, lo, l0 + j), jhi = __shfl_sync(0xffffffffu, hi, l0 + j);
      if (jhi - (jlo & ~3u) <= 128u) continue;  // warp-uniform
      const int bud = __shfl_sync(0xffffffffu, budget, l0 + j);
      const uint32_t pr = __shfl_sync(0xffffffffu, probe, l0 + j);
      const uint64_t gk = (uint64_t)__shfl_sync(0xffffffffu, gid, l0 + j) << p.tbits;
      for (uint32_t c2 = (jlo & ~3u) + lane4 + 128u; c2 < jhi; c2 += 128u * FF_TAIL) {
        uint4 w[FF_TAIL];
#pragma unroll
        for (int c = 0; c < FF_TAIL; ++c)  // chunks past the bucket end re-read its first chunk (unconditional loads, see stream_flat)
          w[c] = ldg128(sd.other + (c2 + 128u * c < jhi ? c2 + 128u * c : (jlo & ~3u) + lane4));
        int best = 64;  // (a chance match in a re-read chunk only costs a trip through the rare path, which skips it)
#pragma unroll
        for (int c = 0; c < FF_TAIL; ++c) {
          const int d0 = base_dist32(w[c].x ^ pr), d1 = base_dist32(w[c].y ^ pr), d2 = base_dist32(w[c].z ^ pr), d3 = base_dist32(w[c].w ^ pr);
          best = min(best, min(min(d0, d1), min(d2, d3)));
        }
        if (best <= bud) {
#pragma unroll
          for (int c = 0; c < FF_TAIL; ++c)
            if (c2 + 128u * c < jhi) verify_chunk<PASS_B, false>(p, wh, sd.canon, w[c], c2 + 128u * c, jlo, jhi, pr, bud, gk, 0u, 0);
        }
      }
    }
  }
  __syncwarp();
  if (*wh.count >= kHW / 2) flush_warp_hits(wh, lane);
}

// The same streaming with every lane busy.  The 16-byte chunks of the batch's buckets are numbered consecutively
// (inclusive prefix sums P of the per-bucket chunk counts, one per lane), and in every round lane l takes chunk
// x = round * 32 + l of that numbering, whichever bucket it falls into.  A part-one bucket of a human-sized index is
// ~18 chunks, so the per-bucket loop above leaves 14 of 32 lanes idle and pays its fixed cost per bucket; here every
// round compares 128 entries.  The bucket of a lane = (# buckets that end before the round: one ballot) + (# bucket
// ends at or below the lane inside the round: the owners' end positions OR-reduced into one 32-bit mask) -- empty
// buckets are squeezed out first so that ends are distinct.  Only the chunk address (A = start - 4 * first chunk
// number, so that address = A + 4 * x) and probe | budget travel by shuffle; bounds and guide index, needed only when
// an entry is within budget, wait in shared memory.
struct BucketRec { uint32_t lo, hi, gid, pad; };

template <bool PASS_B>
__device__ __forceinline__ void verify_flat(const ScanParams &p, const WarpHits &wh, const uint32_t *canon, uint4 v, uint32_t base,
                                            const BucketRec *recs, int j, uint32_t probe, int budget) {
  const int d0 = base_dist32(v.x ^ probe), d1 = base_dist32(v.y ^ probe), d2 = base_dist32(v.z ^ probe), d3 = base_dist32(v.w ^ probe);
  if (min(min(d0, d1), min(d2, d3)) <= budget) {
    const BucketRec r = recs[j];
    const int lo_d = PASS_B ? wh.hA : -1;
    unsigned int ok = 0;
    ok |= (d0 <= budget && d0 > lo_d && base + 0 >= r.lo && base + 0 < r.hi) ? 1u : 0u;
    ok |= (d1 <= budget && d1 > lo_d && base + 1 >= r.lo && base + 1 < r.hi) ? 2u : 0u;
    ok |= (d2 <= budget && d2 > lo_d && base + 2 >= r.lo && base + 2 < r.hi) ? 4u : 0u;
    ok |= (d3 <= budget && d3 > lo_d && base + 3 >= r.lo && base + 3 < r.hi) ? 8u : 0u;
    while (ok) {
      const uint32_t idx = base + (uint32_t)(__ffs((int)ok) - 1);
      ok &= ok - 1u;
      emit_hit(wh, ((uint64_t)r.gid << p.tbits) | (canon ? canon[idx] : idx));
    }
  }
}

constexpr int kFlatChunks = 2;  // 16-byte chunks a lane takes per slot: 2 = one whole, aligned 32-byte sector per lane
constexpr int kFlatRounds = 2;  // slots in flight per lane
template <bool PASS_B, int U>
__device__ __forceinline__ void stream_flat(const ScanParams &p, const SeedSide &sd, const WarpHits &wh, BucketRec *recs, int lane, uint32_t lo,
                                            uint32_t hi, int budget, uint32_t probe, uint32_t gid) {
  constexpr uint32_t kSlot = 4u * kFlatChunks;  // entries per lane slot
  {  // squeeze out empty buckets
    const unsigned int nz = __ballot_sync(0xffffffffu, hi > lo);
    const int n = __popc(nz);
    if (n == 0) return;
    if (nz != (n == 32 ? 0xffffffffu : (1u << n) - 1u)) {  // warp-uniform
      const unsigned int src = __fns(nz, 0, lane + 1) & 31u;
      lo = __shfl_sync(0xffffffffu, lo, src);
      hi = __shfl_sync(0xffffffffu, hi, src);
      budget = __shfl_sync(0xffffffffu, budget, src);
      probe = __shfl_sync(0xffffffffu, probe, src);
      gid = __shfl_sync(0xffffffffu, gid, src);
      if (lane >= n) { lo = 0; hi = 0; }
    }
  }
  const uint32_t s0 = lo & ~(kSlot - 1u);  // slots are aligned to their own size (32 bytes for kFlatChunks = 2)
  const uint32_t c = hi > lo ? (hi - s0 + kSlot - 1u) / kSlot : 0u;
  uint32_t P = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, P, o);
    if (lane >= o) P += t;
  }
  const uint32_t T = __shfl_sync(0xffffffffu, P, 31);
  const uint32_t safe = __shfl_sync(0xffffffffu, s0, 0);  // first slot of the batch (lane 0 holds a non-empty bucket)
  const uint32_t A = s0 - kSlot * (P - c);
  const uint32_t pb = probe | ((uint32_t)budget << 24);
  __syncwarp();
  recs[lane] = BucketRec{lo, hi, gid, 0u};
  __syncwarp();
  const uint32_t le_mask = 0xffffffffu >> (31 - lane);  // lanes <= lane
  for (uint32_t x0 = 0; x0 < T; x0 += 32u * U) {
    uint4 v[U][kFlatChunks];
    uint32_t base[U], pbj[U];
    int jj[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t r0 = x0 + 32u * u;
      const int jb = __popc(__ballot_sync(0xffffffffu, P <= r0));
      const uint32_t rel = P - r0;  // a bucket that ends inside this round ends before lane `rel`
      const uint32_t ends = __reduce_or_sync(0xffffffffu, (rel - 1u < 31u) ? (1u << rel) : 0u);
      const int j = jb + __popc(ends & le_mask);
      jj[u] = j;
      const uint32_t Aj = __shfl_sync(0xffffffffu, A, j);
      pbj[u] = __shfl_sync(0xffffffffu, pb, j);
      base[u] = Aj + kSlot * (r0 + lane);
      // Lanes past the last slot re-read the batch's first slot instead of being predicated off: an unconditional load
      // needs neither the eight register initialisations nor the eight conditional moves a predicated one costs per
      // slot, and the compare below skips those lanes anyway.
      const uint32_t from = r0 + lane < T ? base[u] : safe;
#pragma unroll
      for (int w = 0; w < kFlatChunks; ++w)  // (past a bucket's last chunk the padded array holds neighbouring entries: the range check rejects them)
        v[u][w] = ldg128(sd.other + from + 4u * w);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (x0 + 32u * u + lane < T) {
        const uint32_t pr = pbj[u] & 0xFFFFFFu;
        const int bud = (int)(pbj[u] >> 24);
        int best = 64;
#pragma unroll
        for (int w = 0; w < kFlatChunks; ++w) {
          const int d0 = base_dist32(v[u][w].x ^ pr), d1 = base_dist32(v[u][w].y ^ pr), d2 = base_dist32(v[u][w].z ^ pr), d3 = base_dist32(v[u][w].w ^ pr);
          best = min(best, min(min(d0, d1), min(d2, d3)));
        }
        if (best <= bud) {  // ONE branch per slot (A/B: 5.1 -> 4.8 ms); rare: redo the slot's chunks with the full checks
#pragma unroll
          for (int w = 0; w < kFlatChunks; ++w) verify_flat<PASS_B>(p, wh, sd.canon, v[u][w], base[u] + 4u * w, recs, jj[u], pr, bud);
        }
      }
  }
  __syncwarp();
  if (*wh.count >= kHW / 2) flush_warp_hits(wh, lane);
}

// 6 CTAs per SM (40 registers, ~35 spilled words) beat 4 CTAs with no spills: 5.2 vs 5.6 ms (5 CTAs: 5.3) -- the kernel
// waits on L2 / HBM round trips and on its issue slots, and more resident warps cover both.
#ifndef FF_CELL_MIN_BLOCKS
#define FF_CELL_MIN_BLOCKS 6
#endif
// One launch per index half (PHASE 0: part one, 1: part two): each instantiation carries only its own streaming loop,
// which keeps it inside the 40-register budget with fewer spills than a kernel that holds both.
#ifndef FF_CELL_MIN_BLOCKS_B
#define FF_CELL_MIN_BLOCKS_B FF_CELL_MIN_BLOCKS
#endif
template <int PHASE>
__global__ void __launch_bounds__(kScanThreads, PHASE ? FF_CELL_MIN_BLOCKS_B : FF_CELL_MIN_BLOCKS) k_cell_scan(CellParams cp) {
  __shared__ uint64_t s_hits[kScanWarps * kHW];
  __shared__ unsigned int s_hitn[kScanWarps];
  __shared__ BucketRec s_recs[kScanWarps * 32];
  const ScanParams &p = cp.sp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  BucketRec *recs = s_recs + warp * 32;
  WarpHits wh{s_hits + warp * kHW, &s_hitn[warp], p.hits, p.hit_count, p.hit_cap, p.hA};
  if (lane == 0) s_hitn[warp] = 0;
  __syncwarp();
  unsigned long long compares = 0;
  const long long item_lo = cp.seg_item0[PHASE * kCells * kCells], n_items = cp.seg_item0[(PHASE + 1) * kCells * kCells];
  constexpr int kClaim = 2;
  // Items are claimed from a global counter, kClaim at a time: a static stride lets fast warps run cells ahead
  // of slow ones (measured: L2 hit rate 24 %, 23 GB of HBM reads), claiming in order keeps all resident warps within
  // a fraction of a cell.
  for (;;) {
    unsigned long long first = 0;
    if (lane == 0) first = atomicAdd(cp.next_item + PHASE, (unsigned long long)kClaim);
    first = __shfl_sync(0xffffffffu, first, 0) + (unsigned long long)item_lo;
    if ((long long)first >= n_items) break;
    int seg = PHASE * kCells * kCells;
    {  // the segment that holds the first claimed item: the last one whose first item is <= item
      int hi_s = (PHASE + 1) * kCells * kCells - 1;
      while (seg < hi_s) {
        const int mid = (seg + hi_s + 1) >> 1;
        if (cp.seg_item0[mid] <= (long long)first) seg = mid; else hi_s = mid - 1;
      }
    }
#pragma unroll 1
   for (long long item = (long long)first; item < (long long)first + kClaim && item < n_items; ++item) {
    while (cp.seg_item0[seg + 1] <= item) ++seg;  // the next claimed item may open the next (non-empty) segment
    constexpr int phase = PHASE;
    const int cell = (seg / kCells) % kCells, cls = seg % kCells;
    const int grp = cell ^ cls;
    const int n = phase ? cp.ng_b[grp] : cp.ng_a[grp];
    const int ppi = phase ? cp.ppi_b : 32;
    const int c0 = cp.cls_off[phase * (kCells + 1) + cls];
    const long long n_pairs = (long long)(cp.cls_off[phase * (kCells + 1) + cls + 1] - c0) * n;
    const long long pair = (item - cp.seg_item0[seg]) * ppi + lane;
    uint32_t lo = 0, hi = 0, probe = 0, gid = 0;
    int budget = -1;
    if (lane < ppi && pair < n_pairs) {
      int gl, s;
      if (n_pairs <= 0xFFFFFFFFll) { gl = (int)((uint32_t)pair / (uint32_t)n); s = (int)((uint32_t)pair - (uint32_t)gl * (uint32_t)n); }
      else { gl = (int)(pair / n); s = (int)(pair - (long long)gl * n); }
      gid = (phase ? cp.perm_b : cp.perm_a)[c0 + gl];
      const uint32_t m = phase ? cp.gmasks_b[cp.goff_b[grp] + s] : cp.gmasks_a[cp.goff_a[grp] + s];
      const uint64_t proto = (p.guides[gid] >> p.proto_shift) & p.proto_mask;
      const uint32_t key_a = (uint32_t)(proto >> p.b_bits), key_b = (uint32_t)(proto & ((1ull << p.b_bits) - 1ull));
      const uint32_t kk = (phase ? key_b : key_a) ^ (m & 0xFFFFFFu);
      probe = phase ? key_a : key_b;
      const uint32_t *off = phase ? p.B.off : p.A.off;
      lo = off[kk];
      hi = off[kk + 1];
      budget = p.k - (int)(m >> 24);
    }
    compares += hi - lo;
    const int n_here = (int)min((long long)ppi, n_pairs - (item - cp.seg_item0[seg]) * ppi);
    // part one: short buckets, flattened so that every lane is busy; part two: long buckets, the per-bucket loop
    // (A/B on the GPU: flattening the long buckets gains nothing)
    if (PHASE) stream_pairs<true>(p, p.B, wh, lane, lo, hi, budget, probe, gid, n_here);
    else stream_flat<false, kFlatRounds>(p, p.A, wh, recs, lane, lo, hi, budget, probe, gid);
   }
  }
  flush_warp_hits(wh, lane);
  for (int o = 16; o > 0; o >>= 1) compares += __shfl_down_sync(0xffffffffu, compares, o);
  if (lane == 0 && compares) atomicAdd(p.n_compares, compares);
}

// Fill CellParams for this call and launch the class / segment set-up kernels (no host synchronisation).
static int cell_scan_prepare(ff_ctx *ctx, const ScanParams &sp, int nA_h, int nB_h, CellParams *cp, int *launches) {
  Database &db = ctx->db;
  cudaStream_t st = ctx->stream;
  const int64_t G = sp.n_guides;
  cp->sp = sp;
  cp->gmasks_a = db.A.d_gmasks; cp->gmasks_b = db.B.d_gmasks;
  for (int g = 0; g < kCells; ++g) {
    cp->goff_a[g] = db.A.goff[g]; cp->goff_b[g] = db.B.goff[g];
    cp->ng_a[g] = db.A.gcum[g][std::min(15, nA_h)];
    cp->ng_b[g] = nB_h < 0 ? 0 : db.B.gcum[g][std::min(15, nB_h)];
  }
  cp->ppi_b = sp.B.seeds_per_item >= 32 ? 32 : std::max(4, sp.B.seeds_per_item * 2);
  if (const char *e = getenv("FF_CELL_PPI_B")) cp->ppi_b = std::max(1, std::min(32, atoi(e)));
  FF_TRY(ctx->cell_ws.reserve((size_t)G * 4 * 6 + (2 * kCells + 2 * (kCells + 1)) * 4 + (kSegs + 1) * 8 + 512));
  uint8_t *w = ctx->cell_ws.as<uint8_t>();
  uint32_t *cls_a = (uint32_t *)w; w += G * 4;
  uint32_t *cls_b = (uint32_t *)w; w += G * 4;
  uint32_t *iota = (uint32_t *)w; w += G * 4;
  uint32_t *perm_a = (uint32_t *)w; w += G * 4;
  uint32_t *perm_b = (uint32_t *)w; w += G * 4;
  uint32_t *sorted_cls = (uint32_t *)w; w += G * 4;
  w = (uint8_t *)(((uintptr_t)w + 15) & ~(uintptr_t)15);
  long long *seg_item0 = (long long *)w; w += (kSegs + 1) * 8;
  int *hist = (int *)w; w += 2 * kCells * 4;
  int *cls_off = (int *)w; w += 2 * (kCells + 1) * 4;
  w = (uint8_t *)(((uintptr_t)w + 15) & ~(uintptr_t)15);
  cp->next_item = (unsigned long long *)w;
  FF_CUDA(cudaMemsetAsync(hist, 0, 2 * kCells * 4, st));
  k_guide_classes<<<blocks_for(G, 256), 256, 0, st>>>(sp.guides, G, sp.proto_shift, sp.proto_mask, sp.b_bits, 2 * db.A.key_bases, cls_a, cls_b, iota, hist);
  size_t tmp = 0;
  FF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, cls_a, sorted_cls, iota, perm_a, G, 0, 6, st));
  FF_TRY(ctx->cub_tmp.reserve(tmp));
  FF_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, cls_a, sorted_cls, iota, perm_a, G, 0, 6, st));
  FF_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, cls_b, sorted_cls, iota, perm_b, G, 0, 6, st));
  cp->perm_a = perm_a; cp->perm_b = perm_b; cp->cls_off = cls_off; cp->seg_item0 = seg_item0;
  k_build_segments<<<1, 1024, 0, st>>>(hist, *cp, cls_off, seg_item0);
  *launches += 2 + 2 * 3;
  FF_CUDA(cudaGetLastError());
  return FF_OK;
}
