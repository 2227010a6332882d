// Several GPUs behind one host process (include/flashfry_b200.h, ff_multi_*): what a single-process host such as the
// FlashFry JVM (modules/OffTargetDiscovery.scala:117 "multithreaded (not supported currently)") needs to reach every
// GPU of a box.  The path shards by guide: every device holds a replica of the database, rank r takes
// guides [r G / n, (r + 1) G / n), there is no data-path collective; one ncclAllGather of the per-guide occurrence
// totals at the end gives every device (and the host) the global vector.  One persistent host thread per device.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>

#include "ff_common.cuh"
#include "ff_kernels.cuh"

namespace ff {

// NCCL is resolved at run time (dlopen): the library has no link-time dependency on it, a host that never asks for
// several GPUs never loads it, and inside a process that already carries an NCCL (torch) that copy is the one used.
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (handle) return true;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      handle = dlopen(name, RTLD_NOW | RTLD_NOLOAD);
      if (!handle) handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) return false;
#define FF_SYM(field, sym) field = reinterpret_cast<decltype(field)>(dlsym(handle, sym)); if (!field) return false
    FF_SYM(CommInitAll, "ncclCommInitAll");
    FF_SYM(CommDestroy, "ncclCommDestroy");
    FF_SYM(AllGather, "ncclAllGather");
    FF_SYM(Broadcast, "ncclBroadcast");
    FF_SYM(GroupStart, "ncclGroupStart");
    FF_SYM(GroupEnd, "ncclGroupEnd");
    FF_SYM(GetErrorString, "ncclGetErrorString");
#undef FF_SYM
    return true;
  }
};
static NcclApi g_nccl;
static std::mutex g_nccl_mu;

// a persistent worker thread bound to one device
struct Worker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int()> task;
  bool has_task = false, done = false, quit = false;
  int rc = FF_OK;
  char err[512] = "";
  void loop() {
    for (;;) {
      std::function<int()> t;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return has_task || quit; });
        if (quit) return;
        t = task;
      }
      int r;
      try { r = t(); } catch (...) { set_error("internal error in a device worker"); r = FF_EIO; }
      {
        std::lock_guard<std::mutex> lk(mu);
        rc = r;
        if (r != FF_OK) snprintf(err, sizeof err, "%s", ff_last_error());  // errors are thread-local: carry the text over
        has_task = false; done = true;
      }
      cv.notify_all();
    }
  }
  void submit(std::function<int()> t) {
    { std::lock_guard<std::mutex> lk(mu); task = std::move(t); has_task = true; done = false; }
    cv.notify_all();
  }
  int wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return done; });
    return rc;
  }
};

}  // namespace ff

struct ff_multi {
  int n = 0;
  std::vector<int> devices;
  std::vector<ff_ctx *> ctx;
  std::vector<ncclComm_t> comm;
  std::vector<ff::Worker *> workers;
  std::vector<ff::DevBuf> send, all;  // per device: its shard's totals (padded), the gathered vector
  ff::HostBuf gathered;               // rank 0's copy of the gathered vector
  // database-sharded mode (ff_shard.inl): option shard_mode = 1; exchange blocks are made and mapped on first use
  int shard_mode = 0;
  long long peer_hit_cap = 0;
  int64_t peer_g_cap = 0;
  bool peers_ready = false, warmed = false;
};

using namespace ff;

// run fn(rank) on every device's worker; first failure wins
static int on_all(ff_multi *m, const std::function<int(int)> &fn) {
  for (int r = 0; r < m->n; ++r) m->workers[r]->submit([=]() { return fn(r); });
  int rc = FF_OK;
  for (int r = 0; r < m->n; ++r) {
    const int x = m->workers[r]->wait();
    if (x != FF_OK && rc == FF_OK) { rc = x; set_error("rank %d (device %d): %s", r, m->devices[r], m->workers[r]->err); }
    else if (x != FF_OK) { std::string both = std::string(ff_last_error()) + " | rank " + std::to_string(r) + ": " + m->workers[r]->err; set_error("%s", both.c_str()); }
  }
  return rc;
}

extern "C" {

void ff_shard_range(int64_t n_guides, int n_shards, int shard, int64_t *first, int64_t *count) {
  if (n_shards <= 0 || shard < 0 || shard >= n_shards || n_guides < 0) { if (first) *first = 0; if (count) *count = 0; return; }
  const int64_t lo = n_guides * shard / n_shards, hi = n_guides * (shard + 1) / n_shards;
  if (first) *first = lo;
  if (count) *count = hi - lo;
}

void ff_multi_destroy(ff_multi *m) {
  if (!m) return;
  for (int r = 0; r < (int)m->workers.size(); ++r) {
    Worker *w = m->workers[r];
    { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; }
    w->cv.notify_all();
    if (w->th.joinable()) w->th.join();
    delete w;
  }
  for (int r = 0; r < (int)m->ctx.size(); ++r) {
    if (m->ctx[r]) {
      cudaSetDevice(m->devices[r]);
      if (r < (int)m->send.size()) { m->send[r].release(); m->all[r].release(); }
    }
  }
  for (ncclComm_t c : m->comm) if (c && g_nccl.CommDestroy) g_nccl.CommDestroy(c);
  for (ff_ctx *c : m->ctx) ff_destroy(c);
  m->gathered.release();
  delete m;
}

int ff_multi_create(ff_multi **out, const int *device_ids, int n_devices) {
  if (!out || !device_ids || n_devices <= 0 || n_devices > 64) { set_error("bad device list"); return FF_EINVAL; }
  *out = nullptr;
  try {
    ff_multi *m = new ff_multi();
    m->n = n_devices;
    m->devices.assign(device_ids, device_ids + n_devices);
    m->ctx.assign(n_devices, nullptr);
    m->send.resize(n_devices); m->all.resize(n_devices);
    for (int r = 0; r < n_devices; ++r) {
      const int rc = ff_create(&m->ctx[r], device_ids[r]);
      if (rc != FF_OK) { ff_multi_destroy(m); return rc; }
    }
    bool distinct = true;  // (the same device twice: two ranks on one GPU, for tests of the sharded mode; NCCL refuses that)
    for (int a = 0; a < n_devices; ++a)
      for (int b = a + 1; b < n_devices; ++b) distinct = distinct && device_ids[a] != device_ids[b];
    if (n_devices > 1 && distinct) {
      std::lock_guard<std::mutex> lk(g_nccl_mu);
      if (!g_nccl.load()) { set_error("NCCL (libnccl.so.2) could not be loaded: %s", dlerror() ? dlerror() : "symbol missing"); ff_multi_destroy(m); return FF_EUNSUPPORTED; }
      m->comm.assign(n_devices, nullptr);
      const ncclResult_t nr = g_nccl.CommInitAll(m->comm.data(), n_devices, device_ids);
      if (nr != ncclSuccess) { set_error("ncclCommInitAll: %s", g_nccl.GetErrorString(nr)); m->comm.clear(); ff_multi_destroy(m); return FF_ECUDA; }
    }
    for (int r = 0; r < n_devices; ++r) {
      Worker *w = new Worker();
      m->workers.push_back(w);
      w->th = std::thread([w] { w->loop(); });
    }
    *out = m;
    return FF_OK;
  } catch (const std::bad_alloc &) { set_error("out of host memory"); return FF_ENOMEM; }
  catch (...) { set_error("internal error"); return FF_EIO; }
}

int ff_multi_size(const ff_multi *m) { return m ? m->n : 0; }
ff_ctx *ff_multi_ctx(ff_multi *m, int rank) { return (m && rank >= 0 && rank < m->n) ? m->ctx[rank] : nullptr; }

int ff_multi_set_option(ff_multi *m, const char *key, long long value) {
  if (!m || !key) { set_error("null argument"); return FF_EINVAL; }
  if (strcmp(key, "shard_mode") == 0) {  // 0 = shard the guides (ncclAllGather of totals), 1 = shard the index work (NVLink peer memory)
    if (value < 0 || value > 1) { set_error("option shard_mode: 0 or 1"); return FF_EINVAL; }
    m->shard_mode = (int)value;
    return FF_OK;
  }
  if (strcmp(key, "peer_hit_cap") == 0) {
    if (value < 0 || value > (1ll << 31)) { set_error("option peer_hit_cap out of range"); return FF_EINVAL; }
    m->peer_hit_cap = value; m->peers_ready = false;
    return FF_OK;
  }
  for (ff_ctx *c : m->ctx) FF_TRY(ff_set_option(c, key, value));
  return FF_OK;
}

// database-sharded mode: make the exchange blocks and map them into every context (same process: peer access)
static int multi_peers(ff_multi *m, int64_t n_guides) {
  if (m->peers_ready && n_guides <= m->peer_g_cap) return FF_OK;
  const int n = m->n;
  const int64_t g_cap = std::max<int64_t>(n_guides, 1 << 20);
  std::vector<void *> blocks(n, nullptr);
  for (int r = 0; r < n; ++r) {
    unsigned char handle[FF_PEER_HANDLE_BYTES];
    FF_TRY(ff_peer_export(m->ctx[r], (uint64_t)m->peer_hit_cap, g_cap, handle, &blocks[r]));
  }
  FF_TRY(on_all(m, [=](int r) { return ff_peer_attach(m->ctx[r], r, n, nullptr, blocks.data()); }));
  m->peer_g_cap = g_cap; m->peers_ready = true;
  return FF_OK;
}

int ff_multi_synth_database(ff_multi *m, int enzyme_index, uint64_t n_targets, uint64_t seed) {
  if (!m) { set_error("null argument"); return FF_EINVAL; }
  return on_all(m, [=](int r) { return ff_synth_database(m->ctx[r], enzyme_index, n_targets, seed); });
}

// The files are read and inflated ONCE (device 0's context); the decoded target and position arrays reach the other
// devices over NVLink with ncclBroadcast, and every device builds its own seed index.
int ff_multi_load_database(ff_multi *m, const char *db_path, const char *header_path) {
  if (!m || !db_path) { set_error("null argument"); return FF_EINVAL; }
  if (m->n > 1 && m->comm.empty())  // (two ranks on one device: no NCCL; every context reads the files itself)
    return on_all(m, [=](int r) { return ff_load_database(m->ctx[r], db_path, header_path); });
  FF_TRY(ff_load_database(m->ctx[0], db_path, header_path));
  if (m->n == 1) return FF_OK;
  const Database &src = m->ctx[0]->db;
  const uint64_t n_t = src.n_targets, n_p = src.d_positions ? src.n_positions : 0;
  return on_all(m, [=, &src](int r) -> int {
    ff_ctx *c = m->ctx[r];
    FF_CUDA(cudaSetDevice(c->device));
    if (r != 0) {
      c->db.release();
      c->host_targets_n = 0;
      Database &db = c->db;
      db.pack = src.pack; db.bin_width = src.bin_width; db.n_targets = n_t; db.n_positions = n_p; db.contigs = src.contigs;
      FF_CUDA(cudaMalloc(&db.d_targets, (n_t + 1) * 8));
      if (n_p) FF_CUDA(cudaMalloc(&db.d_positions, (n_p + 1) * 8));
    }
    Database &db = c->db;
    ncclResult_t nr = g_nccl.GroupStart();
    if (nr == ncclSuccess) nr = g_nccl.Broadcast(db.d_targets, db.d_targets, n_t, ncclUint64, 0, m->comm[r], c->stream);
    if (nr == ncclSuccess && n_p) nr = g_nccl.Broadcast(db.d_positions, db.d_positions, n_p, ncclUint64, 0, m->comm[r], c->stream);
    if (nr == ncclSuccess) nr = g_nccl.GroupEnd();
    if (nr != ncclSuccess) { set_error("ncclBroadcast: %s", g_nccl.GetErrorString(nr)); return FF_ECUDA; }
    FF_CUDA(cudaStreamSynchronize(c->stream));
    if (r != 0) {
      const int rc = db_build_index(c);
      if (rc != FF_OK) { c->db.release(); return rc; }
    }
    return FF_OK;
  });
}

int ff_multi_discover(ff_multi *m, const uint64_t *guides, int64_t n_guides, int max_mismatch, int max_off_targets, int want_positions,
                      ff_hits **out, int32_t *total_count_all) {
  if (!m || !out || n_guides < 0 || (n_guides > 0 && !guides)) { set_error("bad argument"); return FF_EINVAL; }
  for (int r = 0; r < m->n; ++r) out[r] = nullptr;
  const int n = m->n;
  if (m->shard_mode == 1 && n > 1 && !want_positions) {
    // every rank scans 1/n of the index for all guides; candidates, barriers and the all-gather of the totals go through
    // peer memory (ff_shard.inl).  A guide set the sharded ordering does not take falls through to the guide-sharded call.
    FF_TRY(multi_peers(m, n_guides));
    // workspaces first, on every rank (ranks that share a device must not allocate while another one waits in a barrier)
    const bool shared_device = m->comm.empty();  // (ranks on ONE device: tests.  Then one rank at a time, kernels pre-loaded)
    auto reserve = [=](int r) -> int {
      FF_CUDA(cudaSetDevice(m->ctx[r]->device));
      FF_TRY(m->ctx[r]->scratch_guides.reserve((n_guides > 0 ? n_guides : 1) * 8));
      return discover_sharded_reserve(m->ctx[r], n_guides, max_mismatch, shared_device && !m->warmed);
    };
    if (shared_device) {
      for (int r = 0; r < n; ++r) {
        m->workers[r]->submit([=]() { return reserve(r); });
        const int x = m->workers[r]->wait();
        if (x != FF_OK) { set_error("rank %d: %s", r, m->workers[r]->err); return x; }
      }
      m->warmed = true;
    } else {
      FF_TRY(on_all(m, reserve));
    }
    const int rc = on_all(m, [=](int r) { return ff_discover_sharded(m->ctx[r], guides, n_guides, max_mismatch, max_off_targets, &out[r]); });
    if (rc == FF_OK) {
      if (total_count_all && n_guides > 0) {
        FF_CUDA(cudaSetDevice(m->ctx[0]->device));
        FF_CUDA(cudaMemcpy(total_count_all, ff_peer_totals_device(m->ctx[0]), (size_t)n_guides * 4, cudaMemcpyDeviceToHost));
      }
      return FF_OK;
    }
    for (int r = 0; r < n; ++r) { if (out[r]) ff_hits_free(out[r]); out[r] = nullptr; }
    m->peers_ready = false;  // (after a failed call the ranks' barrier counters may disagree: fresh blocks next time)
    if (rc != FF_EUNSUPPORTED) return rc;
  }
  if (n > 1 && m->comm.empty()) { set_error("guide-sharded discover needs NCCL, i.e. distinct devices"); return FF_EUNSUPPORTED; }
  const int64_t per = (n_guides + n - 1) / n > 0 ? (n_guides + n - 1) / n : 1;  // padded shard length of the all-gather
  int rc = on_all(m, [=](int r) -> int {
    ff_ctx *c = m->ctx[r];
    int64_t first = 0, count = 0;
    ff_shard_range(n_guides, n, r, &first, &count);
    FF_TRY(ff_discover(c, guides + first, count, max_mismatch, max_off_targets, want_positions, &out[r]));
    if (n == 1) return FF_OK;
    // one all-gather of the per-guide totals: every device ends with the global vector (shards padded to `per`)
    FF_CUDA(cudaSetDevice(c->device));
    FF_TRY(m->send[r].reserve((size_t)per * 4));
    FF_TRY(m->all[r].reserve((size_t)per * 4 * n));
    FF_CUDA(cudaMemsetAsync(m->send[r].p, 0, (size_t)per * 4, c->stream));
    if (count > 0) FF_CUDA(cudaMemcpyAsync(m->send[r].p, out[r]->total_count, (size_t)count * 4, cudaMemcpyHostToDevice, c->stream));
    const ncclResult_t nr = g_nccl.AllGather(m->send[r].p, m->all[r].p, (size_t)per, ncclInt32, m->comm[r], c->stream);
    if (nr != ncclSuccess) { set_error("ncclAllGather: %s", g_nccl.GetErrorString(nr)); return FF_ECUDA; }
    if (r == 0 && total_count_all) {
      FF_TRY(m->gathered.reserve((size_t)per * 4 * n));
      FF_CUDA(cudaMemcpyAsync(m->gathered.p, m->all[r].p, (size_t)per * 4 * n, cudaMemcpyDeviceToHost, c->stream));
    }
    FF_CUDA(cudaStreamSynchronize(c->stream));
    return FF_OK;
  });
  if (rc != FF_OK) {
    for (int r = 0; r < m->n; ++r) { if (out[r]) ff_hits_free(out[r]); out[r] = nullptr; }
    return rc;
  }
  if (total_count_all) {
    for (int r = 0; r < n; ++r) {
      int64_t first = 0, count = 0;
      ff_shard_range(n_guides, n, r, &first, &count);
      if (n == 1) memcpy(total_count_all + first, out[r]->total_count, (size_t)count * 4);
      else memcpy(total_count_all + first, m->gathered.as<int32_t>() + (size_t)r * per, (size_t)count * 4);
    }
  }
  return FF_OK;
}

const int32_t *ff_multi_device_totals(ff_multi *m, int rank) {
  return (m && m->n > 1 && rank >= 0 && rank < m->n) ? m->all[rank].as<int32_t>() : nullptr;
}

}  // extern "C"
