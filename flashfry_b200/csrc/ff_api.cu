// C ABI of libflashfry_b200 (include/flashfry_b200.h): context, error plumbing, host <-> device staging.
#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <chrono>

#include "ff_common.cuh"
#include "ff_kernels.cuh"

namespace ff {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  set_error("CUDA error %s (%s) at %s:%d in %s", cudaGetErrorName(e), cudaGetErrorString(e), file, line, what);
  return e == cudaErrorMemoryAllocation ? FF_ENOMEM : FF_ECUDA;
}

int DevBuf::reserve(size_t bytes) {
  if (bytes <= cap && p) return FF_OK;
  if (p) cudaFree(p);
  p = nullptr; cap = 0;
  size_t want = bytes + bytes / 4 + 256;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) { want = bytes + 256; e = cudaMalloc(&p, want); }
  if (e != cudaSuccess) { p = nullptr; return cuda_fail(e, "cudaMalloc(workspace)", __FILE__, __LINE__); }
  cap = want;
  return FF_OK;
}
void DevBuf::release() { if (p) cudaFree(p); p = nullptr; cap = 0; }

int HostBuf::reserve(size_t bytes) {
  if (bytes <= cap && p) return FF_OK;
  if (p) cudaFreeHost(p);
  p = nullptr; cap = 0;
  const size_t want = bytes + bytes / 4 + 256;
  cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
  if (e != cudaSuccess) { p = nullptr; return cuda_fail(e, "cudaHostAlloc", __FILE__, __LINE__); }
  cap = want;
  return FF_OK;
}
void HostBuf::release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }

// Host-side owner of an ff_hits: pinned buffers recycled through a small process-wide free list so that steady-state
// discover calls do not pay cudaHostAlloc.
struct HitsOwner {
  ff_hits pub;
  HostBuf row_ptr, targets, mm, bulge, pos_ptr, positions, total, ovf, tidx;
};
static std::mutex g_pool_mu;
static std::vector<HitsOwner *> g_pool;

static HitsOwner *owner_get() {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (!g_pool.empty()) { HitsOwner *o = g_pool.back(); g_pool.pop_back(); return o; }
  return new (std::nothrow) HitsOwner();
}
static void owner_put(HitsOwner *o) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (g_pool.size() < 64) { g_pool.push_back(o); return; }  // (one owner per device of an ff_multi, and then some)
  o->row_ptr.release(); o->targets.release(); o->mm.release(); o->bulge.release(); o->pos_ptr.release(); o->positions.release();
  o->total.release(); o->ovf.release(); o->tidx.release();
  delete o;
}

// Grow a pinned host buffer while copies into it may still be in flight on the copy stream.
static int grow_pinned(ff_ctx *ctx, HostBuf &hb, size_t used_bytes, size_t want_bytes) {
  if (want_bytes <= hb.cap && hb.p) return FF_OK;
  FF_CUDA(cudaStreamSynchronize(ctx->copy_stream));  // nothing may be writing into the old block while it moves
  HostBuf bigger;
  FF_TRY(bigger.reserve(want_bytes + want_bytes / 2));
  if (hb.p && used_bytes) memcpy(bigger.p, hb.p, used_bytes);
  hb.release();
  hb = bigger;
  return FF_OK;
}

static int score_slot(ff_ctx *ctx, const uint64_t *d_guides, const DeviceResult &r, uint32_t metrics, int slot) {
  if (!metrics) return FF_OK;
  ff_ctx::OutSlot &os = ctx->out[slot & 1];
  const int64_t Gp = r.n_guides > 0 ? r.n_guides : 1;
  FF_TRY(os.cfd_max.reserve(Gp * 8));
  FF_TRY(os.cfd_spec.reserve(Gp * 8));
  FF_TRY(os.hsu.reserve(Gp * 8));
  cudaEvent_t e0 = ctx->ev[5], e1 = ctx->ev[6];
  FF_CUDA(cudaEventRecord(e0, ctx->stream));
  FF_TRY(score_on_device(ctx, d_guides, r.n_guides, r.d_row_ptr, r.d_targets, r.n_hits, metrics, os.cfd_max.as<double>(),
                         os.cfd_spec.as<double>(), os.hsu.as<double>(), nullptr));
  FF_CUDA(cudaEventRecord(e1, ctx->stream));
  FF_CUDA(cudaStreamSynchronize(ctx->stream));
  FF_CUDA(cudaEventElapsedTime(&ctx->last.score_ms, e0, e1));
  ctx->last.total_ms += ctx->last.score_ms;
  return FF_OK;
}

static void add_timings(ff_timings *acc, const ff_timings &t) {
  acc->prep_ms += t.prep_ms; acc->scan_ms += t.scan_ms; acc->order_ms += t.order_ms; acc->cut_ms += t.cut_ms;
  acc->score_ms += t.score_ms; acc->total_ms += t.total_ms; acc->scan_launches += t.scan_launches;
  acc->kernel_launches += t.kernel_launches; acc->scan_bytes_read += t.scan_bytes_read;
  acc->scan_part1_ms += t.scan_part1_ms; acc->scan_part2_ms += t.scan_part2_ms;
  acc->entries_part1 += t.entries_part1; acc->entries_part2 += t.entries_part2;
}

// The host-facing discover: guides come from host memory, results go back to pinned host memory.  Large guide sets are
// cut into a few sub-batches; the D2H of sub-batch i runs on the copy stream while sub-batch i+1 is being scanned.
static int discover_host(ff_ctx *c, const uint64_t *guides, int64_t n_guides, int max_mm, int max_ot, int want_positions,
                         uint32_t metrics, ff_hits **out, double *cfd_max, double *cfd_spec, double *hsu, int bulge_flags = 0,
                         bool bulge_api = false) {
  if (!c || !out || (n_guides > 0 && !guides)) { set_error("null argument"); return FF_EINVAL; }
  *out = nullptr;
  FF_CUDA(cudaSetDevice(c->device));
  if (!c->db.resident) { set_error("no database resident in this context"); return FF_ENODB; }
  if (n_guides < 0 || max_mm < 0 || max_ot < 0) { set_error("bad discover argument"); return FF_EINVAL; }
  FF_TRY(c->scratch_guides.reserve((n_guides > 0 ? n_guides : 1) * 8));
  if (n_guides > 0) FF_CUDA(cudaMemcpyAsync(c->scratch_guides.p, guides, n_guides * 8, cudaMemcpyHostToDevice, c->stream));
  const uint64_t *d_guides = c->scratch_guides.as<uint64_t>();

  // Sub-batches of DECREASING size: the D2H of a sub-batch hides behind the scan of the next one, so only the last --
  // smallest -- copy is exposed.  Every sub-batch pays the bin scan's fixed cost (the whole index is staged once per
  // call, ~0.25 ms on a human-sized index), which is why two sub-batches beat three at 100 000 guides.
  const int64_t min_batch = std::max(1, c->opt.subbatch_min);
  const int nb = want_positions ? 1 : (int)std::min<int64_t>(3, std::max<int64_t>(1, n_guides / min_batch));
  int kCut[4][4] = {{0, 0, 0, 0}, {0, 100, 100, 100}, {0, 70, 100, 100}, {0, 65, 90, 100}};  // cumulative % (measured, 100 000 guides: 70 / 30 -> 4.16 ms, 60 / 40 -> 4.25, three sub-batches 4.38)
  if (c->opt.subbatch_c1 > 0 && c->opt.subbatch_c1 < c->opt.subbatch_c2 && c->opt.subbatch_c2 < 100) {
    kCut[3][1] = c->opt.subbatch_c1; kCut[3][2] = c->opt.subbatch_c2;
  }
  if (c->opt.subbatch_two > 0 && c->opt.subbatch_two < 100) kCut[2][1] = c->opt.subbatch_two;

  const auto t_start = std::chrono::steady_clock::now();
  auto trace = [&](const char *what, int b) {
    if (!c->opt.trace) return;
    const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_start).count();
    fprintf(stderr, "[ff trace] %8.1f us  %s %d\n", us, what, b);
  };
  HitsOwner *o = owner_get();
  if (!o) { set_error("out of host memory"); return FF_ENOMEM; }
  int rc = FF_OK;
  auto fail = [&](int code) { cudaStreamSynchronize(c->copy_stream); owner_put(o); return code; };
  const int64_t G = n_guides;
  if ((rc = o->row_ptr.reserve((G + 1) * 8)) || (rc = o->total.reserve((G + 1) * 4)) || (rc = o->ovf.reserve(G + 1))) return fail(rc);
  ff_timings acc = {};
  const bool compact = c->opt.compact_hits != 0 && !bulge_api;  // ship database indices, not target longs
  uint64_t n_compares = 0, n_cand = 0;
  int64_t hit_off = 0, pos_total = 0;
  bool with_pos = false;
  std::vector<int64_t> batch_hit_off(nb + 1, 0), batch_g0(nb + 1, 0);
  cudaStream_t cs = c->copy_stream;
  for (int b = 0; b < nb; ++b) {
    const int64_t g0 = G * kCut[nb][b] / 100, g1 = G * kCut[nb][b + 1] / 100, gn = g1 - g0;
    const int slot = b & 1;
    batch_g0[b] = g0;
    if (b >= 2) { cudaError_t e = cudaEventSynchronize(c->slot_copied[slot]); if (e != cudaSuccess) return fail(cuda_fail(e, "event sync", __FILE__, __LINE__)); }
    DeviceResult r;
    trace("sub-batch start", b);
    if ((rc = discover_on_device(c, d_guides + g0, gn, max_mm, max_ot, want_positions != 0, bulge_flags, slot, &r)) != FF_OK) return fail(rc);
    trace("device done", b);
    add_timings(&acc, c->last);
    if (metrics) {
      if ((rc = score_slot(c, d_guides + g0, r, metrics, slot)) != FF_OK) return fail(rc);
      acc.score_ms += c->last.score_ms; acc.total_ms += c->last.score_ms; acc.kernel_launches += 1;
    }
    n_compares += r.n_compares; n_cand += r.n_candidate_hits;
    const int64_t H = r.n_hits;
    if ((rc = compact ? grow_pinned(c, o->tidx, (size_t)hit_off * 4, (size_t)(hit_off + H + 1) * 4)
                      : grow_pinned(c, o->targets, (size_t)hit_off * 8, (size_t)(hit_off + H + 1) * 8)) ||
        (rc = grow_pinned(c, o->mm, (size_t)hit_off, (size_t)(hit_off + H + 1))) ||
        (bulge_api && (rc = grow_pinned(c, o->bulge, (size_t)hit_off, (size_t)(hit_off + H + 1)))))
      return fail(rc);
    // the compute stream is idle here (discover_on_device synchronises), so the slot's contents are final
    cudaError_t e = cudaMemcpyAsync(o->row_ptr.as<int64_t>() + g0, r.d_row_ptr, (gn + 1) * 8, cudaMemcpyDeviceToHost, cs);
    if (compact && !r.d_tidx) { set_error("compact hit lists are not available on this path"); return fail(FF_EUNSUPPORTED); }
    if (e == cudaSuccess && H > 0)
      e = compact ? cudaMemcpyAsync(o->tidx.as<uint32_t>() + hit_off, r.d_tidx, H * 4, cudaMemcpyDeviceToHost, cs)
                  : cudaMemcpyAsync(o->targets.as<uint64_t>() + hit_off, r.d_targets, H * 8, cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess && H > 0) e = cudaMemcpyAsync(o->mm.as<uint8_t>() + hit_off, r.d_mismatches, H, cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess && bulge_api && H > 0) {
      if (r.d_bulge) e = cudaMemcpyAsync(o->bulge.as<uint8_t>() + hit_off, r.d_bulge, H, cudaMemcpyDeviceToHost, cs);
      else memset(o->bulge.as<uint8_t>() + hit_off, 0, (size_t)H);
    }
    if (e == cudaSuccess && gn > 0) e = cudaMemcpyAsync(o->total.as<int32_t>() + g0, r.d_total_count, gn * 4, cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess && gn > 0) e = cudaMemcpyAsync(o->ovf.as<uint8_t>() + g0, r.d_overflowed, gn, cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess && metrics && gn > 0) {
      ff_ctx::OutSlot &os = c->out[slot];
      if (cfd_max && (metrics & FF_METRIC_CFD)) e = cudaMemcpyAsync(cfd_max + g0, os.cfd_max.p, gn * 8, cudaMemcpyDeviceToHost, cs);
      if (e == cudaSuccess && cfd_spec && (metrics & FF_METRIC_CFD)) e = cudaMemcpyAsync(cfd_spec + g0, os.cfd_spec.p, gn * 8, cudaMemcpyDeviceToHost, cs);
      if (e == cudaSuccess && hsu && (metrics & FF_METRIC_HSU2013)) e = cudaMemcpyAsync(hsu + g0, os.hsu.p, gn * 8, cudaMemcpyDeviceToHost, cs);
    }
    if (e == cudaSuccess && r.d_pos_ptr) {  // positions: single batch only
      with_pos = true;
      pos_total = r.n_positions;
      if ((rc = o->pos_ptr.reserve((H + 1) * 8)) || (rc = o->positions.reserve((pos_total + 1) * 8))) return fail(rc);
      e = cudaMemcpyAsync(o->pos_ptr.p, r.d_pos_ptr, (H + 1) * 8, cudaMemcpyDeviceToHost, cs);
      if (e == cudaSuccess && pos_total > 0) e = cudaMemcpyAsync(o->positions.p, r.d_positions, pos_total * 8, cudaMemcpyDeviceToHost, cs);
    }
    if (e == cudaSuccess) e = cudaEventRecord(c->slot_copied[slot], cs);
    trace("copies queued", b);
    if (e != cudaSuccess) return fail(cuda_fail(e, "D2H of discover results", __FILE__, __LINE__));
    batch_hit_off[b] = hit_off;
    hit_off += H;
  }
  batch_hit_off[nb] = hit_off; batch_g0[nb] = G;
  { cudaError_t e = cudaStreamSynchronize(cs); if (e != cudaSuccess) return fail(cuda_fail(e, "D2H of discover results", __FILE__, __LINE__)); }
  trace("copies done", nb);
  // sub-batch row pointers are local: shift them by the batch's first hit
  int64_t *rp = o->row_ptr.as<int64_t>();
  for (int b = 1; b < nb; ++b)
    for (int64_t g = batch_g0[b]; g < batch_g0[b + 1]; ++g) rp[g] += batch_hit_off[b];
  rp[G] = hit_off;
  c->last = acc;
  ff_hits &h = o->pub;
  h.n_guides = G; h.n_hits = hit_off;
  h.row_ptr = rp; h.targets = compact ? nullptr : o->targets.as<uint64_t>(); h.mismatches = o->mm.as<uint8_t>();
  h.target_index = compact ? o->tidx.as<uint32_t>() : nullptr;
  h.pos_ptr = with_pos ? o->pos_ptr.as<int64_t>() : nullptr;
  h.positions = with_pos ? o->positions.as<uint64_t>() : nullptr;
  h.total_count = o->total.as<int32_t>(); h.overflowed = o->ovf.as<uint8_t>();
  h.n_compares = n_compares; h.n_candidate_hits = n_cand;
  h.bulge = bulge_api ? o->bulge.as<uint8_t>() : nullptr;
  h.opaque = o;
  *out = &h;
  return FF_OK;
}

// A caller-built CSR (ff_score / ff_hit_aggregates read row_ptr[G] targets): row_ptr must start at 0, be monotone and
// end at n_hits.
static int check_csr(const ff_hits *h) {
  const int64_t G = h->n_guides;
  if (G < 0 || h->n_hits < 0 || !h->row_ptr || (h->n_hits > 0 && !h->targets)) { set_error("malformed hit list"); return FF_EINVAL; }
  if (h->row_ptr[0] != 0 || h->row_ptr[G] != h->n_hits) { set_error("hit list: row_ptr does not span [0, n_hits]"); return FF_EINVAL; }
  for (int64_t g = 0; g < G; ++g)
    if (h->row_ptr[g + 1] < h->row_ptr[g]) { set_error("hit list: row_ptr is not monotone at row %lld", (long long)g); return FF_EINVAL; }
  return FF_OK;
}

// No exception crosses the C ABI: host allocations are sized from file contents and caller arguments.
template <typename F>
static int guarded(F &&f) noexcept {
  try {
    return f();
  } catch (const std::bad_alloc &) {
    set_error("out of host memory");
    return FF_ENOMEM;
  } catch (const std::exception &e) {
    set_error("internal error: %s", e.what());
    return FF_EIO;
  } catch (...) {
    set_error("internal error");
    return FF_EIO;
  }
}

}  // namespace ff

using namespace ff;

extern "C" {

int ff_abi_version(void) { return 2; }
const char *ff_last_error(void) { return g_err; }

int ff_create(ff_ctx **out, int device_id) {
  return guarded([&]() -> int {
    if (!out) { set_error("null out pointer"); return FF_EINVAL; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
      set_error("no CUDA device available (%s); libflashfry_b200 has no CPU fallback", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
      return FF_ENODEVICE;
    }
    if (device_id < 0 || device_id >= n) { set_error("device %d out of range (0..%d)", device_id, n - 1); return FF_EINVAL; }
    FF_CUDA(cudaSetDevice(device_id));
    ff_ctx *c = new (std::nothrow) ff_ctx();
    if (!c) { set_error("out of host memory"); return FF_ENOMEM; }
    c->device = device_id;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
    c->stream = c->own_stream;
    e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
    for (auto &ev : c->slot_copied) {
      e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
      if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaEventCreate", __FILE__, __LINE__); }
    }
    for (auto &ev : c->ev) {
      e = cudaEventCreate(&ev);
      if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaEventCreate", __FILE__, __LINE__); }
    }
    e = cudaHostAlloc(&c->h_status, 256, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer(&c->h_status_dev, c->h_status, 0);
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaHostAlloc(mapped status)", __FILE__, __LINE__); }
    memset(c->h_status, 0, 256);
    *out = c;
    return FF_OK;
  });
}

static void peer_unmap(ff_ctx *c);

void ff_destroy(ff_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  c->db.release();
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  DevBuf *bufs[] = {&c->cub_tmp, &c->hit_keys, &c->hit_keys_sorted, &c->counters, &c->hit_ranks, &c->seg_start, &c->n_keep, &c->pos_cnt,
                    &c->pos_ptr, &c->out_positions, &c->cfd_per_ot, &c->hsu_per_ot, &c->scratch_guides, &c->running, &c->active, &c->active2,
                    &c->act_flags, &c->seg_end, &c->kept_keys, &c->kept_sorted, &c->n_sel, &c->cell_ws, &c->idx32, &c->st_targets, &c->st_mm};
  for (DevBuf *b : bufs) b->release();
  peer_unmap(c);
  c->peer.block.release();
  if (c->h_status) cudaFreeHost(c->h_status);
  c->host_targets.release();
  for (auto &os : c->out) {
    DevBuf *ob[] = {&os.row_ptr, &os.total_count, &os.overflowed, &os.out_targets, &os.out_mm, &os.out_tidx, &os.out_bulge, &os.cfd_max, &os.cfd_spec, &os.hsu};
    for (DevBuf *b : ob) b->release();
  }
  for (auto &ev : c->slot_copied) if (ev) cudaEventDestroy(ev);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  for (auto &ev : c->ev) if (ev) cudaEventDestroy(ev);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

int ff_set_stream(ff_ctx *c, void *cuda_stream) {
  return guarded([&]() -> int {
    if (!c) { set_error("null context"); return FF_EINVAL; }
    FF_CUDA(cudaSetDevice(c->device));
    FF_CUDA(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return FF_OK;
  });
}

int ff_set_option(ff_ctx *c, const char *key, long long value) {
  return guarded([&]() -> int {
    if (!c || !key) { set_error("null argument"); return FF_EINVAL; }
    struct { const char *name; int *field; long long lo, hi; } table[] = {
        {"scan_kernel", &c->opt.scan_kernel, 0, 2},       {"force_general", &c->opt.force_general, 0, 1},
        {"window_cells", &c->opt.window_cells, 0, 64},    {"subbatch_min", &c->opt.subbatch_min, 1, 1 << 30},
        {"subbatch_c1", &c->opt.subbatch_c1, 1, 98},      {"subbatch_c2", &c->opt.subbatch_c2, 2, 99},
        {"group_sort", &c->opt.group_sort, 0, 1},         {"b_spi", &c->opt.b_spi, 0, 32},
        {"trace", &c->opt.trace, 0, 1},
        {"split_a", &c->opt.split_a, 0, 12},              {"compact_hits", &c->opt.compact_hits, 0, 1},
        {"pair_kernel", &c->opt.pair_kernel, 0, 2},       {"pair_segs", &c->opt.pair_segs, 0, 8},
        {"subbatch_two", &c->opt.subbatch_two, 1, 99},
        {"peer_local_only", &c->opt.peer_local_only, 0, 1}, {"debug_bin_div", &c->opt.debug_bin_div, 0, 64},
    };
    for (auto &t : table)
      if (strcmp(key, t.name) == 0) {
        if (value < t.lo || value > t.hi) { set_error("option %s: value %lld out of range [%lld, %lld]", key, value, t.lo, t.hi); return FF_EINVAL; }
        *t.field = (int)value;
        return FF_OK;
      }
    set_error("unknown option %s", key);
    return FF_EINVAL;
  });
}

int ff_load_database(ff_ctx *c, const char *db_path, const char *header_path) {
  return guarded([&]() -> int {
    if (!c || !db_path) { set_error("null argument"); return FF_EINVAL; }
    FF_CUDA(cudaSetDevice(c->device));
    std::string hp = header_path ? header_path : std::string(db_path) + ".header";  // BinaryHeader.headerExtension
    return db_load_files(c, db_path, hp.c_str());
  });
}

int ff_save_image(ff_ctx *c, const char *image_path) {
  return guarded([&]() -> int {
    if (!c || !image_path) { set_error("null argument"); return FF_EINVAL; }
    FF_CUDA(cudaSetDevice(c->device));
    FF_CUDA(cudaStreamSynchronize(c->stream));
    return db_save_image(c, image_path);
  });
}

int ff_load_image(ff_ctx *c, const char *image_path) {
  return guarded([&]() -> int {
    if (!c || !image_path) { set_error("null argument"); return FF_EINVAL; }
    FF_CUDA(cudaSetDevice(c->device));
    return db_load_image(c, image_path);
  });
}

int ff_load_database_arrays(ff_ctx *c, int enzyme_index, int bin_width, const uint64_t *targets, uint64_t n_targets,
                            const uint64_t *positions, uint64_t n_positions, const char *const *contigs, int n_contigs) {
  return guarded([&]() -> int {
    if (!c || (!targets && n_targets)) { set_error("null argument"); return FF_EINVAL; }
    FF_CUDA(cudaSetDevice(c->device));
    Pack pack;
    FF_TRY(pack_from_index(enzyme_index, &pack));
    std::vector<std::string> names;
    for (int i = 0; contigs && i < n_contigs; ++i) names.push_back(contigs[i] ? contigs[i] : "");
    return db_from_host_arrays(c, pack, bin_width, targets, n_targets, positions, n_positions, names);
  });
}

int ff_synth_database(ff_ctx *c, int enzyme_index, uint64_t n_targets, uint64_t seed) {
  return guarded([&]() -> int {
    if (!c) { set_error("null context"); return FF_EINVAL; }
    FF_CUDA(cudaSetDevice(c->device));
    Pack pack;
    FF_TRY(pack_from_index(enzyme_index, &pack));
    return db_synth(c, pack, n_targets, seed);
  });
}

int ff_synth_database_skewed(ff_ctx *c, int enzyme_index, uint64_t n_targets, uint64_t seed, uint64_t n_families, uint64_t family_size,
                             int family_subs) {
  return guarded([&]() -> int {
    if (!c || family_subs < 0 || family_subs > 8) { set_error("bad argument"); return FF_EINVAL; }
    FF_CUDA(cudaSetDevice(c->device));
    Pack pack;
    FF_TRY(pack_from_index(enzyme_index, &pack));
    return db_synth(c, pack, n_targets, seed, n_families, family_size, family_subs);
  });
}

int ff_db_info(const ff_ctx *c, ff_db_info_t *o) {
  return guarded([&]() -> int {
    if (!c || !o) { set_error("null argument"); return FF_EINVAL; }
    if (!c->db.resident) { set_error("no database resident in this context"); return FF_ENODB; }
    const Database &d = c->db;
    o->enzyme_index = d.pack.enzyme_index; o->bin_width = d.bin_width; o->scan_len = d.pack.scan_len; o->pam_len = d.pack.pam_len;
    o->five_prime_pam = d.pack.five_prime; o->cmp_mask = d.pack.cmp_mask; o->n_targets = d.n_targets; o->n_positions = d.n_positions;
    o->n_contigs = (int)d.contigs.size(); o->seed_split_a = d.A.key_bases; o->device_bytes = d.device_bytes;
    return FF_OK;
  });
}

const char *ff_db_contig(const ff_ctx *c, int contig_id) {
  if (!c || contig_id < 1 || contig_id > (int)c->db.contigs.size()) return nullptr;
  return c->db.contigs[contig_id - 1].c_str();
}

int ff_db_copy_targets(ff_ctx *c, uint64_t first, uint64_t n, uint64_t *out) {
  return guarded([&]() -> int {
    if (!c || !out) { set_error("null argument"); return FF_EINVAL; }
    if (!c->db.resident) { set_error("no database resident in this context"); return FF_ENODB; }
    if (first + n > c->db.n_targets) { set_error("target range out of bounds"); return FF_EINVAL; }
    FF_CUDA(cudaSetDevice(c->device));
    FF_CUDA(cudaMemcpy(out, c->db.d_targets + first, n * 8, cudaMemcpyDeviceToHost));
    return FF_OK;
  });
}

int ff_discover(ff_ctx *c, const uint64_t *guides, int64_t n_guides, int max_mm, int max_ot, int want_positions, ff_hits **out) {
  return guarded([&]() -> int {
    return discover_host(c, guides, n_guides, max_mm, max_ot, want_positions, 0, out, nullptr, nullptr, nullptr);
  });
}

int ff_discover_score(ff_ctx *c, const uint64_t *guides, int64_t n_guides, int max_mm, int max_ot, int want_positions,
                      uint32_t metrics, ff_hits **out, double *cfd_max, double *cfd_spec, double *hsu) {
  return guarded([&]() -> int {
    return discover_host(c, guides, n_guides, max_mm, max_ot, want_positions, metrics, out, cfd_max, cfd_spec, hsu);
  });
}

int ff_discover_bulge(ff_ctx *c, const uint64_t *guides, int64_t n_guides, int max_mm, int max_ot, int bulge_flags, int want_positions,
                      ff_hits **out) {
  return guarded([&]() -> int {
    if (bulge_flags & ~(FF_BULGE_RNA | FF_BULGE_DNA)) { set_error("unknown bulge flag"); return FF_EINVAL; }
    return discover_host(c, guides, n_guides, max_mm, max_ot, want_positions, 0, out, nullptr, nullptr, nullptr, bulge_flags, true);
  });
}

const uint64_t *ff_db_host_targets(ff_ctx *c) {
  if (!c || !c->db.resident) { set_error("no database resident in this context"); return nullptr; }
  if (c->host_targets.p && c->host_targets_n == c->db.n_targets) return c->host_targets.as<uint64_t>();
  if (cudaSetDevice(c->device) != cudaSuccess) return nullptr;
  if (c->host_targets.reserve((c->db.n_targets + 1) * 8) != FF_OK) return nullptr;
  if (cudaMemcpy(c->host_targets.p, c->db.d_targets, c->db.n_targets * 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
    set_error("copy of the target array to the host failed");
    return nullptr;
  }
  c->host_targets_n = c->db.n_targets;
  return c->host_targets.as<uint64_t>();
}

int ff_hits_resolve(ff_ctx *c, ff_hits *h) {
  return guarded([&]() -> int {
    if (!c || !h || !h->opaque) { set_error("null argument"); return FF_EINVAL; }
    if (h->targets || h->n_hits == 0) return FF_OK;
    if (!h->target_index) { set_error("hit list carries neither targets nor indices"); return FF_EINVAL; }
    const uint64_t *mirror = ff_db_host_targets(c);
    if (!mirror) return FF_ENOMEM;
    HitsOwner *o = static_cast<HitsOwner *>(h->opaque);
    FF_TRY(o->targets.reserve((size_t)(h->n_hits + 1) * 8));
    uint64_t *out = o->targets.as<uint64_t>();
    const uint32_t *ix = h->target_index;
    const int64_t H = h->n_hits;
    const uint64_t n_t = c->host_targets_n;
    unsigned nth = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (H < 200000) nth = 1;
    std::vector<int> bad(nth, 0);
    auto work = [&](unsigned w) {
      const int64_t lo = H * w / nth, hi = H * (w + 1) / nth;
      for (int64_t i = lo; i < hi; ++i) {
        if (i + 16 < hi) __builtin_prefetch(mirror + ix[i + 16]);
        if (ix[i] >= n_t) { bad[w] = 1; out[i] = 0; } else out[i] = mirror[ix[i]];
      }
    };
    if (nth == 1) work(0);
    else {
      std::vector<std::thread> pool;
      for (unsigned w = 0; w < nth; ++w) pool.emplace_back(work, w);
      for (auto &t : pool) t.join();
    }
    for (int b : bad) if (b) { set_error("hit list holds a database index out of range"); return FF_EINVAL; }
    h->targets = out;
    return FF_OK;
  });
}

int ff_hits_write_tsv(ff_ctx *c, const char *path, const ff_tsv_guide *guides, const ff_hits *h, int write_positions) {
  return guarded([&]() -> int {
    if (!c || !path || !h || (h->n_guides > 0 && !guides)) { set_error("null argument"); return FF_EINVAL; }
    if (!c->db.resident) { set_error("no database resident in this context"); return FF_ENODB; }
    const uint64_t *mirror = nullptr;
    if (!h->targets && h->n_hits > 0) {
      if (!h->target_index) { set_error("hit list carries neither targets nor indices"); return FF_EINVAL; }
      mirror = ff_db_host_targets(c);
      if (!mirror) return FF_ENOMEM;
    }
    if (write_positions && h->n_hits > 0 && !h->pos_ptr) { set_error("positions were not requested from ff_discover"); return FF_EINVAL; }
    FILE *f = fopen(path, "w");
    if (!f) { set_error("cannot write %s", path); return FF_EIO; }
    const int L = c->db.pack.scan_len;
    const uint64_t cmp_mask = c->db.pack.cmp_mask;
    std::string buf;
    buf.reserve(1 << 22);
    buf += "contig\tstart\tstop\ttarget\tcontext\toverflow\torientation\totCount\toffTargets\n";
    char num[32];
    auto put_int = [&](long long v) { buf.append(num, (size_t)snprintf(num, sizeof num, "%lld", v)); };
    bool ok = true;
    for (int64_t g = 0; g < h->n_guides && ok; ++g) {
      const ff_tsv_guide &gd = guides[g];
      buf += gd.contig ? gd.contig : "";
      buf += '\t'; put_int(gd.start);
      buf += '\t'; put_int((long long)gd.start + (long long)strlen(gd.bases ? gd.bases : ""));
      buf += '\t'; buf += gd.bases ? gd.bases : "";
      buf += '\t'; buf += gd.context ? gd.context : "NONE";
      buf += h->overflowed[g] ? "\tOVERFLOW\t" : "\tOK\t";
      buf += gd.forward ? "FWD\t" : "RVS\t";
      put_int(h->total_count[g]);  // == the summed counts of the listed hits (CRISPRSiteOT.currentTotal)
      buf += '\t';
      for (int64_t i = h->row_ptr[g]; i < h->row_ptr[g + 1]; ++i) {
        if (i > h->row_ptr[g]) buf += ',';
        uint64_t t;
        if (h->targets) t = h->targets[i];
        else if (h->target_index[i] < c->host_targets_n) t = mirror[h->target_index[i]];
        else { set_error("hit list holds a database index out of range"); ok = false; break; }
        char seq[25];
        for (int b = 0; b < L; ++b) seq[b] = "ACGT"[(t >> (2 * (L - 1 - b))) & 3];  // BitEncoding.bitDecodeString :85-99
        buf.append(seq, (size_t)L);
        buf += '_'; put_int((long long)(int16_t)(t >> 48));
        buf += '_'; put_int(h->mismatches[i]);
        (void)cmp_mask;
        if (h->bulge && h->bulge[i]) { buf += (h->bulge[i] & 0xC0) == 0x40 ? "_R" : "_D"; put_int(h->bulge[i] & 0x3F); }
        if (write_positions && h->pos_ptr[i + 1] > h->pos_ptr[i]) {
          buf += '<';
          for (int64_t p = h->pos_ptr[i]; p < h->pos_ptr[i + 1]; ++p) {  // BitPosition.decode :72-92
            const uint64_t pl = h->positions[p];
            const int contig = (int)((pl >> 32) & 0xFFFFF);
            if (p > h->pos_ptr[i]) buf += '|';
            if (contig >= 1 && contig <= (int)c->db.contigs.size()) buf += c->db.contigs[contig - 1];
            else { set_error("position refers to contig %d, the database has %zu", contig, c->db.contigs.size()); ok = false; break; }
            buf += ':'; put_int((long long)(pl & 0xFFFFFFFFull));
            buf += ((pl >> 60) & 0xF) == 0 ? "^F" : "^R";
          }
          buf += '>';
        }
      }
      buf += '\n';
      if (buf.size() > (1u << 22) - 65536) { ok = ok && fwrite(buf.data(), 1, buf.size(), f) == buf.size(); buf.clear(); }
    }
    if (ok && !buf.empty()) ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
    const bool closed = fclose(f) == 0;
    if (!ok) { if (!*ff_last_error()) set_error("short write on %s", path); return FF_EIO; }
    if (!closed) { set_error("short write on %s", path); return FF_EIO; }
    return FF_OK;
  });
}

void ff_hits_free(ff_hits *h) {
  if (!h || !h->opaque) return;
  owner_put(static_cast<HitsOwner *>(h->opaque));
}

int ff_score(ff_ctx *c, const uint64_t *guides, const ff_hits *hits, uint32_t metrics, double *cfd_max, double *cfd_spec,
             double *hsu, double *per_ot_cfd) {
  // the enzyme of the resident database; without one the caller vouches for 23-bp Cas9 longs (ff_score_enzyme says it)
  return ff_score_enzyme(c, c && c->db.resident ? c->db.pack.enzyme_index : 3, guides, hits, metrics, cfd_max, cfd_spec, hsu, per_ot_cfd);
}

int ff_score_enzyme(ff_ctx *c, int enzyme_index, const uint64_t *guides, const ff_hits *hits, uint32_t metrics, double *cfd_max,
                    double *cfd_spec, double *hsu, double *per_ot_cfd) {
  return guarded([&]() -> int {
    if (!c || !hits || (!guides && hits->n_guides > 0)) { set_error("null argument"); return FF_EINVAL; }
    Pack pack;
    FF_TRY(pack_from_index(enzyme_index, &pack));
    if (!(pack.scan_len == 23 && !pack.five_prime)) {  // validOverEnzyme (Doench2016CFDScore.scala:96-98)
      set_error("CFD / Hsu2013 are only valid for 23-bp Cas9 parameter packs");
      return FF_EUNSUPPORTED;
    }
    FF_TRY(check_csr(hits));
    FF_CUDA(cudaSetDevice(c->device));
    const int64_t G = hits->n_guides;
    if (G <= 0 || !metrics) return FF_OK;
    const int64_t H = hits->row_ptr[G];
    cudaStream_t st = c->stream;
    FF_TRY(c->scratch_guides.reserve(G * 8));
    ff_ctx::OutSlot &os = c->out[0];
    FF_TRY(os.row_ptr.reserve((G + 1) * 8));
    FF_TRY(os.out_targets.reserve((H + 1) * 8));
    FF_TRY(os.cfd_max.reserve(G * 8));
    FF_TRY(os.cfd_spec.reserve(G * 8));
    FF_TRY(os.hsu.reserve(G * 8));
    FF_TRY(c->cfd_per_ot.reserve((H + 1) * 8));
    FF_CUDA(cudaMemcpyAsync(c->scratch_guides.p, guides, G * 8, cudaMemcpyHostToDevice, st));
    FF_CUDA(cudaMemcpyAsync(os.row_ptr.p, hits->row_ptr, (G + 1) * 8, cudaMemcpyHostToDevice, st));
    if (H > 0) FF_CUDA(cudaMemcpyAsync(os.out_targets.p, hits->targets, H * 8, cudaMemcpyHostToDevice, st));
    FF_TRY(score_on_device(c, c->scratch_guides.as<uint64_t>(), G, os.row_ptr.as<int64_t>(), os.out_targets.as<uint64_t>(), H, metrics,
                           os.cfd_max.as<double>(), os.cfd_spec.as<double>(), os.hsu.as<double>(), c->cfd_per_ot.as<double>()));
    if (cfd_max && (metrics & FF_METRIC_CFD)) FF_CUDA(cudaMemcpyAsync(cfd_max, os.cfd_max.p, G * 8, cudaMemcpyDeviceToHost, st));
    if (cfd_spec && (metrics & FF_METRIC_CFD)) FF_CUDA(cudaMemcpyAsync(cfd_spec, os.cfd_spec.p, G * 8, cudaMemcpyDeviceToHost, st));
    if (hsu && (metrics & FF_METRIC_HSU2013)) FF_CUDA(cudaMemcpyAsync(hsu, os.hsu.p, G * 8, cudaMemcpyDeviceToHost, st));
    if (per_ot_cfd && (metrics & FF_METRIC_CFD) && H > 0) FF_CUDA(cudaMemcpyAsync(per_ot_cfd, c->cfd_per_ot.p, H * 8, cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    FF_CUDA(cudaGetLastError());
    return FF_OK;
  });
}

int ff_hit_aggregates(ff_ctx *c, int enzyme_index, const uint64_t *guides, const ff_hits *hits, int32_t *closest,
                      int32_t *closest_count, int32_t *hist, int32_t *in_genome) {
  return guarded([&]() -> int {
    if (!c || !hits || (!guides && hits->n_guides > 0)) { set_error("null argument"); return FF_EINVAL; }
    FF_CUDA(cudaSetDevice(c->device));
    Pack pack;
    FF_TRY(pack_from_index(enzyme_index, &pack));
    FF_TRY(check_csr(hits));
    const int64_t G = hits->n_guides;
    if (G <= 0) return FF_OK;
    const int64_t H = hits->row_ptr[G];
    cudaStream_t st = c->stream;
    ff_ctx::OutSlot &os = c->out[0];
    FF_TRY(c->scratch_guides.reserve(G * 8));
    FF_TRY(os.row_ptr.reserve((G + 1) * 8));
    FF_TRY(os.out_targets.reserve((H + 1) * 8));
    FF_TRY(os.total_count.reserve(G * 8 * 4));  // 8 int32 per guide
    FF_CUDA(cudaMemcpyAsync(c->scratch_guides.p, guides, G * 8, cudaMemcpyHostToDevice, st));
    FF_CUDA(cudaMemcpyAsync(os.row_ptr.p, hits->row_ptr, (G + 1) * 8, cudaMemcpyHostToDevice, st));
    if (H > 0) FF_CUDA(cudaMemcpyAsync(os.out_targets.p, hits->targets, H * 8, cudaMemcpyHostToDevice, st));
    FF_TRY(hit_aggregates_on_device(c, c->scratch_guides.as<uint64_t>(), G, os.row_ptr.as<int64_t>(), os.out_targets.as<uint64_t>(),
                                    pack.cmp_mask, os.total_count.as<int32_t>()));
    std::vector<int32_t> tmp((size_t)G * 8);
    FF_CUDA(cudaMemcpyAsync(tmp.data(), os.total_count.p, (size_t)G * 8 * 4, cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    for (int64_t g = 0; g < G; ++g) {
      const int32_t *o = tmp.data() + g * 8;
      if (closest) closest[g] = o[0];
      if (closest_count) closest_count[g] = o[1];
      if (hist) for (int m = 0; m < 5; ++m) hist[g * 5 + m] = o[2 + m];
      if (in_genome) in_genome[g] = o[7];
    }
    return FF_OK;
  });
}

int ff_discover_device(ff_ctx *c, const uint64_t *d_guides, int64_t n_guides, int max_mm, int max_ot, uint32_t metrics,
                       ff_device_result *out) {
  return guarded([&]() -> int {
    if (!c || !out) { set_error("null argument"); return FF_EINVAL; }
    FF_CUDA(cudaSetDevice(c->device));
    DeviceResult r;
    FF_TRY(discover_on_device(c, d_guides, n_guides, max_mm, max_ot, false, 0, 0, &r));
    FF_TRY(score_slot(c, d_guides, r, metrics, 0));
    out->d_bulge = nullptr;
    out->n_guides = r.n_guides; out->n_hits = r.n_hits; out->n_candidate_hits = r.n_candidate_hits; out->n_compares = r.n_compares;
    out->d_row_ptr = r.d_row_ptr; out->d_targets = r.d_targets; out->d_mismatches = r.d_mismatches;
    out->d_total_count = r.d_total_count; out->d_overflowed = r.d_overflowed;
    out->d_cfd_max = (metrics & FF_METRIC_CFD) ? c->out[0].cfd_max.as<double>() : nullptr;
    out->d_cfd_specificity = (metrics & FF_METRIC_CFD) ? c->out[0].cfd_spec.as<double>() : nullptr;
    out->d_hsu2013 = (metrics & FF_METRIC_HSU2013) ? c->out[0].hsu.as<double>() : nullptr;
    return FF_OK;
  });
}

int ff_discover_bulge_device(ff_ctx *c, const uint64_t *d_guides, int64_t n_guides, int max_mm, int max_ot, int bulge_flags,
                             ff_device_result *out) {
  return guarded([&]() -> int {
    if (!c || !out) { set_error("null argument"); return FF_EINVAL; }
    FF_CUDA(cudaSetDevice(c->device));
    DeviceResult r;
    FF_TRY(discover_on_device(c, d_guides, n_guides, max_mm, max_ot, false, bulge_flags, 0, &r));
    out->n_guides = r.n_guides; out->n_hits = r.n_hits; out->n_candidate_hits = r.n_candidate_hits; out->n_compares = r.n_compares;
    out->d_row_ptr = r.d_row_ptr; out->d_targets = r.d_targets; out->d_mismatches = r.d_mismatches;
    out->d_total_count = r.d_total_count; out->d_overflowed = r.d_overflowed;
    out->d_cfd_max = out->d_cfd_specificity = out->d_hsu2013 = nullptr;
    out->d_bulge = r.d_bulge;
    return FF_OK;
  });
}

// ---- database-sharded discover (ff_shard.inl) --------------------------------------------------------------
static void peer_unmap(ff_ctx *c) {
  PeerLink &pl = c->peer;
  for (int r = 0; r < kMaxPeers; ++r) {
    if (pl.ipc_opened[r] && pl.base[r]) cudaIpcCloseMemHandle(pl.base[r]);
    pl.ipc_opened[r] = false;
    pl.base[r] = nullptr;
  }
  pl.ready = false;
}

int ff_peer_export(ff_ctx *c, uint64_t hit_cap, int64_t guide_cap, void *handle_out, void **block_out) {
  return guarded([&]() -> int {
    if (!c || !handle_out) { set_error("null argument"); return FF_EINVAL; }
    static_assert(sizeof(cudaIpcMemHandle_t) == FF_PEER_HANDLE_BYTES, "IPC handle size");
    FF_CUDA(cudaSetDevice(c->device));
    PeerLink &pl = c->peer;
    peer_unmap(c);
    if (hit_cap == 0) hit_cap = 1ull << 24;
    if (guide_cap <= 0) guide_cap = 1ll << 20;
    if (hit_cap > (1ull << 31) || guide_cap > (1ll << 28)) { set_error("exchange block too large"); return FF_EINVAL; }
    pl.block.release();
    FF_TRY(pl.block.reserve(kPeerHeadBytes + (size_t)hit_cap * 8 + (size_t)guide_cap * 4));
    FF_CUDA(cudaMemset(pl.block.p, 0, kPeerHeadBytes));
    pl.hit_cap = (size_t)hit_cap; pl.g_cap = guide_cap; pl.epoch = 0; pl.fresh = true;
    cudaIpcMemHandle_t h;
    FF_CUDA(cudaIpcGetMemHandle(&h, pl.block.p));
    memcpy(handle_out, &h, sizeof(h));
    if (block_out) *block_out = pl.block.p;
    return FF_OK;
  });
}

int ff_peer_attach(ff_ctx *c, int rank, int world, const void *handles, void *const *blocks) {
  return guarded([&]() -> int {
    if (!c) { set_error("null context"); return FF_EINVAL; }
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) { set_error("bad rank / world (at most %d ranks)", kMaxPeers); return FF_EINVAL; }
    if (world > 1 && !handles && !blocks) { set_error("neither IPC handles nor block pointers given"); return FF_EINVAL; }
    PeerLink &pl = c->peer;
    if (!pl.block.p || !pl.fresh) { set_error("ff_peer_export first (every attach needs a freshly exported block: its counters start at 0)"); return FF_EINVAL; }
    FF_CUDA(cudaSetDevice(c->device));
    peer_unmap(c);
    for (int r = 0; r < world; ++r) {
      if (r == rank) { pl.base[r] = pl.block.as<uint8_t>(); continue; }
      if (blocks) {  // same process: enable peer access to the block's device (unless it is this device)
        cudaPointerAttributes at;
        FF_CUDA(cudaPointerGetAttributes(&at, blocks[r]));
        if (at.type != cudaMemoryTypeDevice) { set_error("block of rank %d is not device memory", r); return FF_EINVAL; }
        if (at.device != c->device) {
          int can = 0;
          FF_CUDA(cudaDeviceCanAccessPeer(&can, c->device, at.device));
          if (!can) { set_error("device %d cannot access device %d", c->device, at.device); return FF_EUNSUPPORTED; }
          const cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
          cudaGetLastError();
        }
        pl.base[r] = static_cast<uint8_t *>(blocks[r]);
      } else {
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const uint8_t *>(handles) + (size_t)r * FF_PEER_HANDLE_BYTES, sizeof(h));
        void *p = nullptr;
        FF_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        pl.base[r] = static_cast<uint8_t *>(p);
        pl.ipc_opened[r] = true;
      }
    }
    // (the block's counters were zeroed by ff_peer_export, i.e. before any peer could know its handle: a peer that has
    //  attached already may arrive at this rank's barrier counter at once)
    pl.rank = rank; pl.world = world; pl.epoch = 0; pl.fresh = false;
    pl.ready = true;
    return FF_OK;
  });
}

int ff_peer_detach(ff_ctx *c) {
  return guarded([&]() -> int {
    if (!c) { set_error("null context"); return FF_EINVAL; }
    FF_CUDA(cudaSetDevice(c->device));
    cudaStreamSynchronize(c->stream);
    peer_unmap(c);
    return FF_OK;
  });
}

const int32_t *ff_peer_totals_device(ff_ctx *c) {
  if (!c || !c->peer.block.p) return nullptr;
  return reinterpret_cast<const int32_t *>(c->peer.block.as<uint8_t>() + kPeerHeadBytes + c->peer.hit_cap * 8);
}

int ff_discover_sharded_device(ff_ctx *c, const uint64_t *d_guides_all, int64_t n_all, int max_mm, int max_ot, uint32_t metrics,
                               ff_device_result *out) {
  return guarded([&]() -> int {
    if (!c || !out) { set_error("null argument"); return FF_EINVAL; }
    FF_CUDA(cudaSetDevice(c->device));
    DeviceResult r;
    FF_TRY(discover_sharded(c, d_guides_all, n_all, max_mm, max_ot, 0, &r));
    const int64_t first = c->peer.world > 0 ? n_all * c->peer.rank / c->peer.world : 0;
    FF_TRY(score_slot(c, d_guides_all + first, r, metrics, 0));
    out->d_bulge = nullptr;
    out->n_guides = r.n_guides; out->n_hits = r.n_hits; out->n_candidate_hits = r.n_candidate_hits; out->n_compares = r.n_compares;
    out->d_row_ptr = r.d_row_ptr; out->d_targets = r.d_targets; out->d_mismatches = r.d_mismatches;
    out->d_total_count = r.d_total_count; out->d_overflowed = r.d_overflowed;
    out->d_cfd_max = (metrics & FF_METRIC_CFD) ? c->out[0].cfd_max.as<double>() : nullptr;
    out->d_cfd_specificity = (metrics & FF_METRIC_CFD) ? c->out[0].cfd_spec.as<double>() : nullptr;
    out->d_hsu2013 = (metrics & FF_METRIC_HSU2013) ? c->out[0].hsu.as<double>() : nullptr;
    return FF_OK;
  });
}

int ff_discover_sharded(ff_ctx *c, const uint64_t *guides_all, int64_t n_all, int max_mm, int max_ot, ff_hits **out) {
  return guarded([&]() -> int {
    if (!c || !out || (n_all > 0 && !guides_all)) { set_error("null argument"); return FF_EINVAL; }
    *out = nullptr;
    if (n_all < 0) { set_error("bad discover argument"); return FF_EINVAL; }
    FF_CUDA(cudaSetDevice(c->device));
    FF_TRY(c->scratch_guides.reserve((n_all > 0 ? n_all : 1) * 8));
    if (n_all > 0) FF_CUDA(cudaMemcpyAsync(c->scratch_guides.p, guides_all, n_all * 8, cudaMemcpyHostToDevice, c->stream));
    DeviceResult r;
    FF_TRY(discover_sharded(c, c->scratch_guides.as<uint64_t>(), n_all, max_mm, max_ot, 0, &r));
    HitsOwner *o = owner_get();
    if (!o) { set_error("out of host memory"); return FF_ENOMEM; }
    int rc = FF_OK;
    const int64_t G = r.n_guides, H = r.n_hits;
    const bool compact = c->opt.compact_hits != 0;
    if ((rc = o->row_ptr.reserve((G + 1) * 8)) || (rc = o->total.reserve((G + 1) * 4)) || (rc = o->ovf.reserve(G + 1)) ||
        (rc = o->mm.reserve(H + 1)) || (rc = compact ? o->tidx.reserve((H + 1) * 4) : o->targets.reserve((H + 1) * 8))) {
      owner_put(o);
      return rc;
    }
    cudaStream_t st = c->stream;  // (the stream is idle: discover_sharded waits for its status words)
    cudaError_t e = cudaMemcpyAsync(o->row_ptr.p, r.d_row_ptr, (G + 1) * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && H > 0)
      e = compact ? cudaMemcpyAsync(o->tidx.p, r.d_tidx, H * 4, cudaMemcpyDeviceToHost, st) : cudaMemcpyAsync(o->targets.p, r.d_targets, H * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && H > 0) e = cudaMemcpyAsync(o->mm.p, r.d_mismatches, H, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && G > 0) e = cudaMemcpyAsync(o->total.p, r.d_total_count, G * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && G > 0) e = cudaMemcpyAsync(o->ovf.p, r.d_overflowed, G, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { owner_put(o); return cuda_fail(e, "D2H of discover results", __FILE__, __LINE__); }
    ff_hits &h = o->pub;
    h.n_guides = G; h.n_hits = H;
    h.row_ptr = o->row_ptr.as<int64_t>(); h.targets = compact ? nullptr : o->targets.as<uint64_t>(); h.mismatches = o->mm.as<uint8_t>();
    h.target_index = compact ? o->tidx.as<uint32_t>() : nullptr;
    h.pos_ptr = nullptr; h.positions = nullptr;
    h.total_count = o->total.as<int32_t>(); h.overflowed = o->ovf.as<uint8_t>();
    h.n_compares = r.n_compares; h.n_candidate_hits = r.n_candidate_hits;
    h.bulge = nullptr;
    h.opaque = o;
    *out = &h;
    return FF_OK;
  });
}

int ff_last_timings(const ff_ctx *c, ff_timings *out) {
  return guarded([&]() -> int {
    if (!c || !out) { set_error("null argument"); return FF_EINVAL; }
    *out = c->last;
    return FF_OK;
  });
}

}  // extern "C"
