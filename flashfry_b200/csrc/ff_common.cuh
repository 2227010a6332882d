// Shared declarations of libflashfry_b200: context layout, device buffers, error plumbing.
// Product code: nothing under oracle/ is ever included or linked here.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/flashfry_b200.h"

namespace ff {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define FF_CUDA(expr)                                                          \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) return ff::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define FF_TRY(expr)          \
  do {                        \
    int _rc = (expr);         \
    if (_rc != FF_OK) return _rc; \
  } while (0)

// A grow-only device allocation (steady-state calls never hit cudaMalloc).
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);  // contents are NOT preserved when it grows
  void release();
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Pinned host allocation, pooled by the context and lent to ff_hits.
struct HostBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);
  void release();
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// standards/StandardScanParameters.scala:28-48 -- the fields the device path needs
struct Pack {
  int enzyme_index = 0, scan_len = 0, pam_len = 0, five_prime = 0;
  uint64_t cmp_mask = 0;
};
int pack_from_index(int enzyme_index, Pack *out);  // StandardScanParameters.scala:61-70

constexpr int kPrefixBases = 7;             // device-side first-level prefix (the reference's default bin width)
constexpr int kNumBins = 1 << (2 * kPrefixBases);
constexpr int kMaxSubBases = 7;

struct Database {
  bool resident = false;
  Pack pack;
  int bin_width = 7;         // of the file header; the device index always uses kPrefixBases
  uint64_t n_targets = 0, n_positions = 0;
  int sub_bases = 0;         // s: depth of the sub-bin index below the 7-mer
  uint64_t *d_targets = nullptr;   // [n_targets] target longs, database order
  uint32_t *d_tlow = nullptr;      // [n_targets + pad] low words (everything below the 7-mer prefix)
  uint32_t *d_sub_off = nullptr;   // [4^(7+s) + 1] first target of every (7+s)-mer
  uint64_t *d_pos_off = nullptr;   // [n_targets + 1] exclusive scan of counts (only with positions)
  uint64_t *d_positions = nullptr; // [n_positions]
  uint16_t *d_mask7 = nullptr;     // 4^7 XOR masks sorted by Hamming distance (in bases)
  uint16_t *d_submask = nullptr;   // 4^s XOR masks sorted by Hamming distance
  uint32_t *d_submask32 = nullptr; // the same, packed as mask | distance << 16
  int m7off[kPrefixBases + 2] = {0};   // m7off[d] .. m7off[d+1] = masks at distance exactly d
  int nsub[kMaxSubBases + 2] = {0};    // nsub[r] = # sub masks at distance <= r
  uint64_t device_bytes = 0;
  std::vector<std::string> contigs;
  void release();
};

struct Hits;  // host-side result owner (ff_api.cu)

}  // namespace ff

struct ff_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  ff::Database db;

  // ---- per-call workspaces (grow-only) ----
  ff::DevBuf guides, gkeys, gkeys_sorted, gentry, gentry_sorted, goff, cub_tmp;
  ff::DevBuf hit_keys, hit_keys_sorted, counters;
  ff::DevBuf seg_start, n_keep, row_ptr, total_count, overflowed, out_targets, out_mm, out_tidx;
  ff::DevBuf pos_cnt, pos_ptr, out_positions;
  ff::DevBuf cfd_per_ot, hsu_per_ot, cfd_max, cfd_spec, hsu;
  ff::DevBuf scratch_guides;  // H2D staging target for ff_discover
  size_t hit_cap = 0;

  cudaEvent_t ev[8] = {nullptr};
  ff_timings last = {};
  std::vector<ff::HostBuf *> host_pool;
};
