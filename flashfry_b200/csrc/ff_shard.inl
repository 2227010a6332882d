// Database-sharded discover over NVLink peer memory (included by ff_discover.cu, inside namespace ff).
//
// The reference is a single process (modules/OffTargetDiscovery.scala:117) and has nothing to compare with; the unit of
// work it hands a traverser is still "all guides against the whole database" (reference/traverser/Traverser.scala:52-59).
// Guide sharding (one index replica per GPU, each rank scans the whole index for 1/N of the guides) stops scaling
// when the shard gets small: the index is streamed once per call whatever the batch.  Here the INDEX WORK is split
// instead: every rank scans bins / buckets [rank/N, (rank+1)/N) of both index halves for ALL guides, and the drain of
// the scan kernels pushes every candidate straight into the exchange block of the rank that owns the guide -- fire-and-
// forget P2P stores over NVLink / NVSwitch into the region the owner keeps for this source (positions from LOCAL
// atomics: no remote round trip in the scan), no separate all-to-all step.  After a barrier (arrival counters in the
// same blocks: the path's only remote atomics) every owner orders and cuts its guides' candidates
// with the single-GPU pipeline, and the per-guide totals are written into every rank's block (the path's one
// all-gather, also as peer stores).  Results are identical to the single-GPU call by construction: the candidate SET of
// a guide does not depend on who found it, and the ordering pipeline sorts by database index.

struct PeerPtrs { uint8_t *base[kMaxPeers]; };

// Arrive at every rank's counter (own included), then wait until all `world` ranks have arrived at ours `epoch` times.
__global__ void k_peer_barrier(PeerPtrs pp, int world, int rank, unsigned int target, unsigned int *flag, unsigned int flag_bit) {
  if ((int)threadIdx.x < world) {
    __threadfence_system();
    atomicAdd_system(&reinterpret_cast<PeerCtr *>(pp.base[threadIdx.x])->arrive, 1u);
  }
  if (threadIdx.x == 0) {
    volatile unsigned int *a = &reinterpret_cast<PeerCtr *>(pp.base[rank])->arrive;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int)(*a - target) < 0) {
      __nanosleep(100);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 4000000000ull) { atomicOr(flag, flag_bit); break; }  // a peer never came (failed call): give up after 4 s
    }
    __threadfence_system();
  }
}

// End of a rank's scan: tell every owner how many keys its region `rank` holds (one remote store each).
__global__ void k_peer_counts(PeerPtrs pp, int world, int rank) {
  const int o = threadIdx.x;
  if (o < world) reinterpret_cast<PeerCtr *>(pp.base[o])->recv[rank] = reinterpret_cast<const PeerCtr *>(pp.base[rank])->sent[o];
  __threadfence_system();
}

// The owner: its `world` regions, each filled by one source, into one contiguous key array; the total into *n_out.  A region
// that was sent more keys than it holds leaves the total above `cap` (the call then fails: nothing is truncated silently).
__global__ void k_peer_compact(const uint8_t *__restrict__ block, int world, size_t region_cap, uint64_t *__restrict__ out, size_t cap,
                               unsigned long long *__restrict__ n_out) {
  const PeerCtr *pc = reinterpret_cast<const PeerCtr *>(block);
  const uint64_t *keys = reinterpret_cast<const uint64_t *>(block + kPeerHead);
  size_t base = 0;
  bool over = false;
  for (int s = 0; s < world; ++s) {
    const size_t n = pc->recv[s];
    over = over || n > region_cap;
    const size_t m = n < region_cap ? n : region_cap;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x)
      if (base + i < cap) out[base + i] = keys[(size_t)s * region_cap + i];
    base += m;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_out = over ? (unsigned long long)cap + 1ull : (unsigned long long)base;
}

// The owner's per-guide totals into every rank's totals region (the all-gather of the path, as peer stores).
__global__ void k_peer_totals(PeerPtrs pp, int world, const int32_t *__restrict__ total, int64_t n_own, int64_t first, size_t totals_off) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= n_own) return;
  const int32_t v = total[g];
  for (int r = 0; r < world; ++r) reinterpret_cast<int32_t *>(pp.base[r] + totals_off)[first + g] = v;
}

static int peer_barrier(ff_ctx *ctx, unsigned int *d_flag, int which) {
  PeerLink &pl = ctx->peer;
  PeerPtrs pp;
  for (int r = 0; r < kMaxPeers; ++r) pp.base[r] = r < pl.world ? pl.base[r] : nullptr;
  pl.epoch += 1;
  k_peer_barrier<<<1, 32, 0, ctx->stream>>>(pp, pl.world, pl.rank, pl.epoch * (unsigned int)pl.world, d_flag, 0x100u << which);
  FF_CUDA(cudaGetLastError());
  return FF_OK;
}

// reserve_only: make every workspace of the call and launch nothing.  Ranks that share ONE device (tests) must all have
// done this before any of them queues a barrier: a cudaMalloc waits for the device, i.e. for the other rank's spinning
// barrier kernel, which waits for this rank.  Ranks on different devices need not bother.
static int discover_sharded_impl(ff_ctx *ctx, const uint64_t *d_guides_all, int64_t n_all, int max_mm, int max_ot, int slot, DeviceResult *res,
                                 bool reserve_only) {
  Database &db = ctx->db;
  PeerLink &pl = ctx->peer;
  ff_ctx::OutSlot &os = ctx->out[slot & 1];
  if (!db.resident) { set_error("no database resident in this context"); return FF_ENODB; }
  if (!pl.ready) { set_error("ff_peer_attach has not been called on this context"); return FF_EINVAL; }
  if (n_all < 0 || max_mm < 0 || max_ot < 0 || (n_all > 0 && !d_guides_all && !reserve_only)) { set_error("bad discover argument"); return FF_EINVAL; }
  if (n_all > pl.g_cap) { set_error("more guides (%lld) than the exchange blocks were sized for (%lld)", (long long)n_all, (long long)pl.g_cap); return FF_EINVAL; }
  cudaStream_t st = ctx->stream;
  int launches = 0;
  ff_timings tm = {};
  const int world = pl.world, rank = pl.rank;
  ScanShard shard;
  shard.rank = rank; shard.world = world;
  const size_t region_cap = pl.hit_cap / (size_t)world;
  shard.sink.world = world; shard.sink.rank = rank; shard.sink.hit_cap = region_cap;
  shard.sink.owner_scale = n_all > 0 ? (float)world / (float)n_all : 0.f;
  for (int r = 0; r <= world; ++r) shard.sink.first[r] = (unsigned int)(n_all * r / world);  // ff_shard_range
  for (int r = 0; r < world; ++r) shard.sink.peer[r] = ctx->opt.peer_local_only ? pl.base[rank] : pl.base[r];  // (diagnostic: no NVLink stores, wrong rows)
  const int64_t first = shard.sink.first[rank], G = (int64_t)shard.sink.first[rank + 1] - first, Gp = G > 0 ? G : 1;
  const uint64_t *d_guides = d_guides_all + first;

  ScanParams sp;
  int hA = 0, nA = 1, nB = 0;
  fill_scan_params(ctx, d_guides_all, n_all, max_mm, &sp, &hA, &nA, &nB);
  if (!bin_scan_supported(db, hA, n_all)) { set_error("the database-sharded discover needs the bin-major scan (seed budget / index layout not supported)"); return FF_EUNSUPPORTED; }
  FF_TRY(ctx->counters.reserve(256));
  FF_TRY(ctx->seg_start.reserve((Gp + 1) * 8));
  FF_TRY(ctx->n_keep.reserve((Gp + 1) * 8));
  FF_TRY(os.row_ptr.reserve((Gp + 1) * 8));
  FF_TRY(os.total_count.reserve(Gp * 4));
  FF_TRY(os.overflowed.reserve(Gp));
  const size_t cap = pl.hit_cap;
  FF_TRY(os.out_targets.reserve((cap + 1) * 8));
  FF_TRY(os.out_mm.reserve(cap + 1));
  FF_TRY(os.out_tidx.reserve((cap + 1) * 4));
  FF_TRY(ctx->running.reserve((size_t)(Gp + 1) * 4 * 2 + kLongCap * 4));
  static bool long_attr[64] = {false};
  if (!long_attr[ctx->device & 63]) {
    FF_CUDA(cudaFuncSetAttribute(k_sort_long, cudaFuncAttributeMaxDynamicSharedMemorySize, kLongMax * 4));
    long_attr[ctx->device & 63] = true;
  }
  PlainStatus *d_stt = ctx->counters.as<PlainStatus>();
  PlainStatus *h_stt = static_cast<PlainStatus *>(ctx->h_status);
  PlainStatus *h_stt_dev = static_cast<PlainStatus *>(ctx->h_status_dev);
  PeerCtr *own = reinterpret_cast<PeerCtr *>(pl.base[rank]);
  FF_TRY(ctx->hit_keys.reserve(cap * 8));
  uint64_t *own_hits = ctx->hit_keys.as<uint64_t>();
  sp.hit_count = &d_stt->n_cand; sp.n_compares = &d_stt->n_compares;
  PeerPtrs ptrs;
  for (int r = 0; r < kMaxPeers; ++r) ptrs.base[r] = r < world ? pl.base[r] : nullptr;

  if (reserve_only) {
    BinScanPlan plan;
    FF_TRY(bin_scan_prepare(ctx, sp, hA, nB, &d_stt->n_compares_b, &plan, &launches, &shard, true));
    FF_TRY(ctx->idx32.reserve((cap + 1) * 4));
    FF_TRY(ctx->st_targets.reserve((cap + 1) * 8));
    FF_TRY(ctx->st_mm.reserve(cap + 1));
    FF_TRY(ctx->cub_tmp.reserve(1 << 20));
    FF_CUDA(cudaStreamSynchronize(st));
    return FF_OK;
  }
  FF_CUDA(cudaEventRecord(ctx->ev[0], st));
  FF_CUDA(cudaMemsetAsync(d_stt, 0, sizeof(PlainStatus), st));
  FF_CUDA(cudaMemsetAsync(own->sent, 0, sizeof(own->sent) + sizeof(own->recv), st));
  // (every allocation of the scan happens before the first barrier is queued: with several ranks on ONE device -- the
  //  tests -- a cudaMalloc of one rank would wait for the other rank's spinning barrier kernel)
  BinScanPlan bpl;
  FF_TRY(bin_scan_prepare(ctx, sp, hA, nB, &d_stt->n_compares_b, &bpl, &launches, &shard));
  FF_TRY(ctx->idx32.reserve((cap + 1) * 4));
  FF_TRY(ctx->st_targets.reserve((cap + 1) * 8));
  FF_TRY(ctx->st_mm.reserve(cap + 1));
  FF_TRY(ctx->cub_tmp.reserve(1 << 20));
  FF_TRY(peer_barrier(ctx, &d_stt->flag, 0));  // every block is empty (and the previous step's keys have been consumed)
  FF_CUDA(cudaEventRecord(ctx->ev[1], st));
  if (n_all > 0) FF_TRY(bin_scan_launch(ctx, &bpl, sp, nullptr, &launches));
  else FF_CUDA(cudaEventRecord(ctx->ev[7], st));
  k_peer_counts<<<1, 32, 0, st>>>(ptrs, world, rank);
  FF_TRY(peer_barrier(ctx, &d_stt->flag, 1));  // every rank has scanned its part: this block holds all candidates of the own guides
  FF_CUDA(cudaEventRecord(ctx->ev[2], st));
  k_peer_compact<<<ctx->sm_count * 4, 256, 0, st>>>(pl.base[rank], world, region_cap, own_hits, cap, &d_stt->n_cand);
  launches += 2;
  unsigned int *cnt = ctx->running.as<unsigned int>(), *cursor = cnt + (Gp + 1);
  FF_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(Gp + 1) * 4 * 2, st));
  if (G > 0) FF_TRY(order_grouped(ctx, os, own_hits, d_stt, cap, sp.tbits, cnt, cursor, true, d_guides, G, max_ot, &launches));
  else { FF_CUDA(cudaEventRecord(ctx->ev[3], st)); FF_CUDA(cudaMemsetAsync(os.row_ptr.p, 0, 8, st)); FF_CUDA(cudaEventRecord(ctx->ev[4], st)); }
  if (G > 0) k_peer_totals<<<blocks_for(G, 256), 256, 0, st>>>(ptrs, world, os.total_count.as<int32_t>(), G, first, kPeerHead + pl.hit_cap * 8);
  launches++;
  FF_TRY(peer_barrier(ctx, &d_stt->flag, 2));  // every rank's totals region is complete
  FF_CUDA(cudaEventRecord(ctx->ev[4], st));
  launches += 3;
  k_publish_status<<<1, 32, 0, st>>>(d_stt, nullptr, h_stt_dev, ++ctx->status_seq);
  FF_TRY(wait_status(ctx, st, ctx->status_seq));
  if (h_stt->flag & 0x700u) {
    set_error("database-sharded discover: rank %d waited 4 s for its peers at barrier(s) 0x%x (1 = before, 2 = after the scan, 4 = totals); epoch %u",
              rank, (h_stt->flag >> 8) & 7u, pl.epoch);
    return FF_ECUDA;
  }
  const int64_t n_cand = (int64_t)h_stt->n_cand;
  if ((size_t)n_cand > cap) {
    set_error("database-sharded discover: more candidates for this rank's guides than the exchange block holds (%zu keys, %zu per source rank): "
              "ff_peer_export with a larger hit_cap", cap, region_cap);
    return FF_ENOMEM;
  }
  if ((h_stt->flag & 0xFFu) && !ctx->opt.peer_local_only) { set_error("database-sharded discover: a guide with more candidates than the per-guide sort takes; use the guide-sharded call"); return FF_EUNSUPPORTED; }
  const int64_t n_hits = G > 0 ? h_stt->n_hits : 0;
  FF_CUDA(cudaEventSynchronize(ctx->ev[4]));
  FF_CUDA(cudaEventElapsedTime(&tm.prep_ms, ctx->ev[0], ctx->ev[1]));
  FF_CUDA(cudaEventElapsedTime(&tm.scan_ms, ctx->ev[1], ctx->ev[2]));
  FF_CUDA(cudaEventElapsedTime(&tm.order_ms, ctx->ev[2], ctx->ev[3]));
  FF_CUDA(cudaEventElapsedTime(&tm.cut_ms, ctx->ev[3], ctx->ev[4]));
  FF_CUDA(cudaEventElapsedTime(&tm.total_ms, ctx->ev[0], ctx->ev[4]));
  tm.scan_launches = 1; tm.kernel_launches = launches;
  tm.entries_part1 = h_stt->n_compares; tm.entries_part2 = h_stt->n_compares_b;
  tm.scan_part1_ms = tm.scan_ms; tm.scan_part2_ms = 0.f;
  if (nB > 0 && n_all > 0) {
    FF_CUDA(cudaEventElapsedTime(&tm.scan_part1_ms, ctx->ev[1], ctx->ev[7]));
    FF_CUDA(cudaEventElapsedTime(&tm.scan_part2_ms, ctx->ev[7], ctx->ev[2]));
  }
  ctx->last = tm;
  res->n_guides = G; res->n_hits = n_hits; res->n_positions = 0;
  res->n_candidate_hits = (uint64_t)n_cand; res->n_compares = h_stt->n_compares + h_stt->n_compares_b;
  res->d_row_ptr = os.row_ptr.as<int64_t>(); res->d_targets = os.out_targets.as<uint64_t>();
  res->d_mismatches = os.out_mm.as<uint8_t>(); res->d_total_count = os.total_count.as<int32_t>();
  res->d_overflowed = os.overflowed.as<uint8_t>(); res->d_bulge = nullptr;
  res->d_tidx = os.out_tidx.as<uint32_t>();
  res->d_pos_ptr = nullptr; res->d_positions = nullptr;
  return FF_OK;
}

int discover_sharded(ff_ctx *ctx, const uint64_t *d_guides_all, int64_t n_all, int max_mm, int max_ot, int slot, DeviceResult *res) {
  return discover_sharded_impl(ctx, d_guides_all, n_all, max_mm, max_ot, slot, res, false);
}

// warm: also make the driver LOAD every kernel of the path now (CUDA loads kernels lazily, at their first launch, and such
// a load waits for the device like a cudaMalloc does): a tiny single-rank discover through the same scan and ordering
// kernels + the attributes of the exchange kernels.  The caller runs this for one rank at a time.
int discover_sharded_reserve(ff_ctx *ctx, int64_t n_all, int max_mm, bool warm) {
  DeviceResult r;
  FF_TRY(discover_sharded_impl(ctx, nullptr, n_all, max_mm, 0, 0, &r, true));
  if (!warm) return FF_OK;
  const Options saved = ctx->opt;
  ctx->opt.scan_kernel = 2;
  int rc = FF_OK;
  const int64_t n = (int64_t)std::min<uint64_t>(8, ctx->db.n_targets);
  for (int pk = 1; pk <= 2 && rc == FF_OK && n > 0; ++pk) {
    ctx->opt.pair_kernel = pk;
    rc = discover_plain(ctx, ctx->db.d_targets, n, max_mm, 2000, false, 0, &r);  // (the first targets serve as guides)
  }
  ctx->opt = saved;
  FF_TRY(rc);
  cudaFuncAttributes fa;
  FF_CUDA(cudaFuncGetAttributes(&fa, k_peer_barrier));
  FF_CUDA(cudaFuncGetAttributes(&fa, k_peer_counts));
  FF_CUDA(cudaFuncGetAttributes(&fa, k_peer_compact));
  FF_CUDA(cudaFuncGetAttributes(&fa, k_peer_totals));
  FF_CUDA(cudaFuncGetAttributes(&fa, k_guide_hist));
  FF_CUDA(cudaStreamSynchronize(ctx->stream));
  return FF_OK;
}
