// Database-sharded discover over NVLink peer memory (included by ff_discover.cu, inside namespace ff).
//
// The reference is a single process (modules/OffTargetDiscovery.scala:117) and has nothing to compare with; the unit of
// work it hands a traverser is still "all guides against the whole database" (reference/traverser/Traverser.scala:52-59).
// Guide sharding (one index replica per GPU, each rank scans the whole index for 1/N of the guides) stops scaling
// when the shard gets small: the index is streamed once per call whatever the batch.  Here the INDEX WORK is split
// instead: every rank scans bins / buckets [rank/N, (rank+1)/N) of both index halves for ALL guides, and the drain of
// the scan kernels pushes every candidate straight into the exchange block of the rank that owns the guide -- P2P
// stores and one remote atomic per (warp drain, owner) over NVLink / NVSwitch, no separate all-to-all step.  After a
// barrier (arrival counters in the same blocks, remote atomics again) every owner orders and cuts its guides' candidates
// with the single-GPU pipeline, and the per-guide totals are written into every rank's block (the path's one
// all-gather, also as peer stores).  Results are identical to the single-GPU call by construction: the candidate SET of
// a guide does not depend on who found it, and the ordering pipeline sorts by database index.

struct PeerPtrs { uint8_t *base[kMaxPeers]; };

// Arrive at every rank's counter (own included), then wait until all `world` ranks have arrived at ours `epoch` times.
__global__ void k_peer_barrier(PeerPtrs pp, int world, int rank, unsigned int target, unsigned int *flag) {
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
      if (t1 - t0 > 4000000000ull) { atomicOr(flag, 0x100u); break; }  // a peer never came (failed call): give up after 4 s
    }
    __threadfence_system();
  }
}

// The owner's per-guide totals into every rank's totals region (the all-gather of the path, as peer stores).
__global__ void k_peer_totals(PeerPtrs pp, int world, const int32_t *__restrict__ total, int64_t n_own, int64_t first, size_t totals_off) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= n_own) return;
  const int32_t v = total[g];
  for (int r = 0; r < world; ++r) reinterpret_cast<int32_t *>(pp.base[r] + totals_off)[first + g] = v;
}

static int peer_barrier(ff_ctx *ctx, unsigned int *d_flag) {
  PeerLink &pl = ctx->peer;
  PeerPtrs pp;
  for (int r = 0; r < kMaxPeers; ++r) pp.base[r] = r < pl.world ? pl.base[r] : nullptr;
  pl.epoch += 1;
  k_peer_barrier<<<1, 32, 0, ctx->stream>>>(pp, pl.world, pl.rank, pl.epoch * (unsigned int)pl.world, d_flag);
  FF_CUDA(cudaGetLastError());
  return FF_OK;
}

int discover_sharded(ff_ctx *ctx, const uint64_t *d_guides_all, int64_t n_all, int max_mm, int max_ot, int slot, DeviceResult *res) {
  Database &db = ctx->db;
  PeerLink &pl = ctx->peer;
  ff_ctx::OutSlot &os = ctx->out[slot & 1];
  if (!db.resident) { set_error("no database resident in this context"); return FF_ENODB; }
  if (!pl.ready) { set_error("ff_peer_attach has not been called on this context"); return FF_EINVAL; }
  if (n_all < 0 || max_mm < 0 || max_ot < 0 || (n_all > 0 && !d_guides_all)) { set_error("bad discover argument"); return FF_EINVAL; }
  if (n_all > pl.g_cap) { set_error("more guides (%lld) than the exchange blocks were sized for (%lld)", (long long)n_all, (long long)pl.g_cap); return FF_EINVAL; }
  cudaStream_t st = ctx->stream;
  int launches = 0;
  ff_timings tm = {};
  const int world = pl.world, rank = pl.rank;
  ScanShard shard;
  shard.rank = rank; shard.world = world;
  shard.sink.world = world; shard.sink.hit_cap = pl.hit_cap;
  for (int r = 0; r <= world; ++r) shard.sink.first[r] = (unsigned int)(n_all * r / world);  // ff_shard_range
  for (int r = 0; r < world; ++r) shard.sink.peer[r] = pl.base[r];
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
  uint64_t *own_hits = reinterpret_cast<uint64_t *>(pl.base[rank] + kPeerHead);
  sp.hit_count = &d_stt->n_cand; sp.n_compares = &d_stt->n_compares;

  FF_CUDA(cudaEventRecord(ctx->ev[0], st));
  FF_CUDA(cudaMemsetAsync(d_stt, 0, sizeof(PlainStatus), st));
  FF_CUDA(cudaMemsetAsync(&own->hit_count, 0, 8, st));
  FF_TRY(peer_barrier(ctx, &d_stt->flag));  // every block is empty (and the previous step's rows have been consumed)
  BinScanPlan bpl;
  FF_TRY(bin_scan_prepare(ctx, sp, hA, nB, &d_stt->n_compares_b, &bpl, &launches, &shard));
  FF_CUDA(cudaEventRecord(ctx->ev[1], st));
  if (n_all > 0) FF_TRY(bin_scan_launch(ctx, &bpl, sp, nullptr, &launches));
  else FF_CUDA(cudaEventRecord(ctx->ev[7], st));
  FF_TRY(peer_barrier(ctx, &d_stt->flag));  // every rank has scanned its part: this block holds all candidates of the own guides
  FF_CUDA(cudaEventRecord(ctx->ev[2], st));
  FF_CUDA(cudaMemcpyAsync(&d_stt->n_cand, &own->hit_count, 8, cudaMemcpyDeviceToDevice, st));
  unsigned int *cnt = ctx->running.as<unsigned int>(), *cursor = cnt + (Gp + 1);
  FF_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(Gp + 1) * 4 * 2, st));
  if (G > 0) FF_TRY(order_grouped(ctx, os, own_hits, d_stt, cap, sp.tbits, cnt, cursor, true, d_guides, G, max_ot, &launches));
  else { FF_CUDA(cudaEventRecord(ctx->ev[3], st)); FF_CUDA(cudaMemsetAsync(os.row_ptr.p, 0, 8, st)); FF_CUDA(cudaEventRecord(ctx->ev[4], st)); }
  {
    PeerPtrs pp;
    for (int r = 0; r < kMaxPeers; ++r) pp.base[r] = r < world ? pl.base[r] : nullptr;
    if (G > 0) k_peer_totals<<<blocks_for(G, 256), 256, 0, st>>>(pp, world, os.total_count.as<int32_t>(), G, first, kPeerHead + pl.hit_cap * 8);
    launches++;
  }
  FF_TRY(peer_barrier(ctx, &d_stt->flag));  // every rank's totals region is complete
  FF_CUDA(cudaEventRecord(ctx->ev[4], st));
  launches += 3;
  k_publish_status<<<1, 32, 0, st>>>(d_stt, nullptr, h_stt_dev, ++ctx->status_seq);
  FF_TRY(wait_status(ctx, st, ctx->status_seq));
  if (h_stt->flag & 0x100u) { set_error("database-sharded discover: a peer rank did not reach the barrier"); return FF_ECUDA; }
  const int64_t n_cand = (int64_t)h_stt->n_cand;
  if ((size_t)n_cand > cap) {
    set_error("database-sharded discover: %lld candidates for this rank's guides, the exchange block holds %zu (ff_peer_export with a larger hit_cap)",
              (long long)n_cand, cap);
    return FF_ENOMEM;
  }
  if (h_stt->flag) { set_error("database-sharded discover: a guide with more candidates than the per-guide sort takes; use the guide-sharded call"); return FF_EUNSUPPORTED; }
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
